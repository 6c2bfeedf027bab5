"""Post-processing entry points with the reference's signatures (basics/utils/general.py)."""
import math

import torch

from ... import ops


def make_divisible(x, divisor):
    return math.ceil(x / divisor) * divisor


def xywh2xyxy(x):
    """[n,4] (cx, cy, w, h) -> (x1, y1, x2, y2); reference general.py:269."""
    y = x.clone()
    half = x[:, 2:4] / 2
    y[:, 0:2] = x[:, 0:2] - half
    y[:, 2:4] = x[:, 0:2] + half
    return y


def box_iou(box1, box2):
    """Pairwise IoU of [N,4] and [M,4] xyxy boxes; reference general.py:392."""
    a1 = (box1[:, 2] - box1[:, 0]) * (box1[:, 3] - box1[:, 1])
    a2 = (box2[:, 2] - box2[:, 0]) * (box2[:, 3] - box2[:, 1])
    wh = (torch.min(box1[:, None, 2:], box2[:, 2:]) - torch.max(box1[:, None, :2], box2[:, :2])).clamp(0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (a1[:, None] + a2 - inter)


def non_max_suppression_padded(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                               multi_label=False, out=None, counts=None, want_keep_idx=False):
    """Device-resident form: (det [B,300,6], counts [B] int32, keep_idx | None), no host sync.
    ``out`` / ``counts`` may be views into a communication buffer."""
    return ops.nms(prediction.float(), conf_thres, iou_thres, classes, agnostic, multi_label, out=out, counts=counts,
                   want_keep_idx=want_keep_idx)


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        labels=()):
    """Reference signature (general.py:425): list of [n_i, 6] tensors (xyxy, conf, cls), one per image.

    The whole batch is processed by one sequence of kernels; the only host sync is the read of
    the per-image counts needed to build the Python list.  ``labels`` (autolabelling, general.py:451-458): per image a tensor
    [n, 5] of (class, cx, cy, w, h); each label joins its image's candidates as a row with objectness 1 and a one-hot class
    score -- here as extra prediction rows behind the image's own (images with fewer labels are padded with objectness-0 rows,
    which the scan drops).  The reference's 10 s watchdog does not exist here.
    """
    prediction = prediction.float()
    if labels and any(len(l) for l in labels):
        B, _, no = prediction.shape
        n_max = max(len(l) for l in labels)
        extra = torch.zeros((B, n_max, no), dtype=torch.float32, device=prediction.device)
        for i, l in enumerate(labels):
            if len(l):
                l = torch.as_tensor(l, dtype=torch.float32, device=prediction.device)
                extra[i, :len(l), :4] = l[:, 1:5]
                extra[i, :len(l), 4] = 1.0
                extra[i, torch.arange(len(l), device=prediction.device), l[:, 0].long() + 5] = 1.0
        prediction = torch.cat((prediction, extra), dim=1)
    det, counts, _ = ops.nms(prediction, conf_thres, iou_thres, classes, agnostic, multi_label)
    return [det[i, :n] for i, n in enumerate(counts.tolist())]
