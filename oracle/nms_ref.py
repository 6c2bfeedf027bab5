"""Oracle (test infrastructure): non_max_suppression and the greedy NMS it calls.

Reference: basics/utils/general.py:425-512 (``non_max_suppression``), :269-276
(``xywh2xyxy``), :392-414 (``box_iou``).  The greedy suppression itself is
``torchvision.ops.nms`` (general.py:496), a third-party dependency whose source is not
under /root/reference (requirements.txt:11 ``torchvision>=0.8.1``; 0.26.0 installed).
Its published CPU algorithm is restated in ``greedy_nms`` below:

    order = stable argsort of scores, descending
    for i in order (not yet suppressed): keep i; for every later j not suppressed:
        inter = max(0, min(x2)-max(x1)) * max(0, min(y2)-max(y1))
        ovr   = inter / (area_i + area_j - inter)          (all float32)
        suppress j  iff  ovr > iou_threshold               (float promoted to double)

All arithmetic below is numpy float32 so that index sets are bit-exact with the
reference run on the same float32 predictions.  Tie rule for the ``n > max_nms``
cut (general.py:490-491, an unstable torch argsort in the reference) is pinned here
as STABLE (equal scores keep candidate order); the reference leaves it unspecified.
"""
import numpy as np

F32 = np.float32


def xywh2xyxy(b):
    """general.py:269-276; float32 in, float32 out."""
    b = np.asarray(b, dtype=F32)
    half_w = b[:, 2] / F32(2)
    half_h = b[:, 3] / F32(2)
    return np.stack([b[:, 0] - half_w, b[:, 1] - half_h, b[:, 0] + half_w, b[:, 1] + half_h], axis=1)


def _iou_row(box, boxes, area, areas):
    lt = np.maximum(box[:2], boxes[:, :2])
    rb = np.minimum(box[2:], boxes[:, 2:])
    wh = np.maximum(rb - lt, F32(0))
    inter = wh[:, 0] * wh[:, 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / (area + areas - inter)


def greedy_nms(boxes, scores, iou_thres, max_keep=None):
    """torchvision.ops.nms CPU semantics (see module docstring).  Returns kept indices
    (into ``boxes``) in descending-score order.  ``max_keep`` stops early once that many
    boxes are kept -- the prefix is identical to the full run because greedy NMS is
    sequential in score order."""
    boxes = np.ascontiguousarray(boxes, dtype=F32)
    scores = np.asarray(scores, dtype=F32)
    n = boxes.shape[0]
    order = np.argsort(-scores, kind="stable")
    sb = boxes[order]
    areas = (sb[:, 2] - sb[:, 0]) * (sb[:, 3] - sb[:, 1])
    dead = np.zeros(n, dtype=bool)
    keep = []
    thr = float(iou_thres)
    for p in range(n):
        if dead[p]:
            continue
        keep.append(order[p])
        if max_keep is not None and len(keep) >= max_keep:
            break
        if p + 1 < n:
            ovr = _iou_row(sb[p], sb[p + 1:], areas[p], areas[p + 1:])
            dead[p + 1:] |= ovr.astype(np.float64) > thr
    return np.asarray(keep, dtype=np.int64)


def box_iou(b1, b2):
    """general.py:392-414, float32: inter / (area1[:,None] + area2 - inter)."""
    b1 = np.asarray(b1, dtype=F32)
    b2 = np.asarray(b2, dtype=F32)
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    wh = np.maximum(np.minimum(b1[:, None, 2:], b2[None, :, 2:]) - np.maximum(b1[:, None, :2], b2[None, :, :2]), F32(0))
    inter = wh[..., 0] * wh[..., 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / (a1[:, None] + a2[None, :] - inter)


def candidates(pred_img, conf_thres, multi_label, classes=None):
    """Rows of the n x 6 detection matrix (xyxy, conf, cls) of one image, in the
    reference's candidate order.  general.py:449-480."""
    conf_t = F32(conf_thres)
    x = np.asarray(pred_img, dtype=F32)
    x = x[x[:, 4] > conf_t]
    if x.shape[0] == 0:
        return np.zeros((0, 6), dtype=F32)
    cls_conf = x[:, 5:] * x[:, 4:5]
    box = xywh2xyxy(x[:, :4])
    if multi_label:
        i, j = np.nonzero(cls_conf > conf_t)  # row-major, like torch.nonzero
        det = np.concatenate([box[i], cls_conf[i, j][:, None], j[:, None].astype(F32)], axis=1)
    else:
        j = np.argmax(cls_conf, axis=1)  # first maximum, like torch.max on CPU
        conf = cls_conf[np.arange(x.shape[0]), j]
        det = np.concatenate([box, conf[:, None], j[:, None].astype(F32)], axis=1)[conf > conf_t]
    if classes is not None:
        det = det[np.isin(det[:, 5], np.asarray(classes, dtype=F32))]
    return det.astype(F32)


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, max_det=300, max_nms=30000, max_wh=4096, merge=True,
                        redundant=True, return_indices=False, early_stop=False, labels=()):
    """general.py:425-512 on a float32 numpy array [B, R, 5+nc].

    Returns a list of [n_i, 6] float32 arrays; with ``return_indices`` also, per image,
    the kept row numbers into that image's candidate matrix after the max_nms cut
    (the quantity that must be bit-exact).  The 10 s wall-clock watchdog
    (general.py:439,508-510) is deliberately not restated.
    ``labels`` (general.py:451-458, autolabelling): per image an array [n, 5] of (class, cx, cy, w, h); every label joins
    the image's candidates AFTER the objectness filter as a row with objectness 1 and a one-hot class score of 1.
    """
    prediction = np.asarray(prediction, dtype=F32)
    nc = prediction.shape[2] - 5
    multi_label = bool(multi_label) and nc > 1
    outs, idxs = [], []
    for xi, img in enumerate(prediction):
        if labels and len(labels[xi]):
            # rows with objectness 1 pass the objectness filter whenever conf_thres < 1 (and die at the class-confidence filter
            # otherwise, as in the reference), so appending them before the filter gives the reference's candidate list and order
            l = np.asarray(labels[xi], dtype=F32)
            v = np.zeros((l.shape[0], nc + 5), dtype=F32)
            v[:, :4] = l[:, 1:5]
            v[:, 4] = 1.0
            v[np.arange(l.shape[0]), l[:, 0].astype(np.int64) + 5] = 1.0
            img = np.concatenate([img, v], 0)
        x = candidates(img, conf_thres, multi_label, classes)
        n = x.shape[0]
        if n == 0:
            outs.append(np.zeros((0, 6), dtype=F32))
            idxs.append(np.zeros((0,), dtype=np.int64))
            continue
        if n > max_nms:
            x = x[np.argsort(-x[:, 4], kind="stable")[:max_nms]]
        c = x[:, 5:6] * F32(0 if agnostic else max_wh)
        boxes = x[:, :4] + c
        scores = x[:, 4]
        keep = greedy_nms(boxes, scores, iou_thres, max_keep=max_det if early_stop else None)
        keep = keep[:max_det]
        x = x.copy()
        if merge and 1 < n < 3000:
            hit = box_iou(boxes[keep], boxes) > F32(iou_thres)
            wts = hit.astype(F32) * scores[None, :]
            merged = (wts.astype(np.float64) @ x[:, :4].astype(np.float64)) / wts.astype(np.float64).sum(1, keepdims=True)
            x[keep, :4] = merged.astype(F32)
            if redundant:
                keep = keep[hit.sum(1) > 1]
        outs.append(x[keep])
        idxs.append(keep)
    return (outs, idxs) if return_indices else outs
