"""Case tables shared by the golden generator (tests/golden/make_golden.py keeps its own
copy -- it must not import repo test code next to the reference) and the tests."""

# name -> (dim, (H, W), heads, ws, shift, linear_mlp, B)
SWIN_CASES = {
    "hd16_shift2_convmlp": (96, (16, 16), 6, 8, 2, False, 2),
    "hd16_shift0_linmlp": (96, (16, 16), 6, 8, 0, True, 2),
    "hd32_shift2_rect": (64, (24, 16), 2, 8, 2, False, 1),
    "pad_shift2": (48, (12, 12), 3, 8, 2, True, 2),
    "pad_shift0_rect": (48, (10, 12), 3, 8, 0, False, 1),
    "global_clamped": (128, (8, 8), 2, 32, 0, True, 2),
    "ws7_shift3": (48, (14, 14), 3, 7, 3, True, 1),
    "ws4_shift1_hd8": (32, (8, 12), 4, 4, 1, False, 2),
    "global_n256_hd64": (128, (16, 16), 2, 32, 0, True, 1),
}

# name -> (variant, C, heads, (h, w), ws, B)
CATTN_CASES = {
    "vit_ws1": ("vit", 48, 12, (16, 16), 1, 2),
    "vit_ws1_shift1": ("vit_shift1", 48, 12, (128, 128), 1, 1),
    "v2_ws2": ("v2", 24, 12, (16, 16), 2, 2),
    "v2_ws3_pad": ("v2", 48, 12, (16, 16), 3, 1),
    "v2_ws7_pad": ("v2", 48, 6, (16, 20), 7, 1),
    "v2_ws8": ("v2", 96, 4, (16, 16), 8, 1),
    "v2_ws4_h1": ("v2", 24, 1, (8, 8), 4, 2),
}

# name -> (B, R, img, active, seed, kwargs)
NMS_CASES = {
    "single_label": (3, 4096, 256, 0.2, 0, dict(conf_thres=0.25, iou_thres=0.45)),
    "multi_label": (2, 2048, 256, 0.1, 1, dict(conf_thres=0.001, iou_thres=0.6, multi_label=True)),
    "agnostic": (2, 2048, 128, 0.2, 2, dict(conf_thres=0.25, iou_thres=0.45, agnostic=True)),
    "class_filter": (2, 4096, 256, 0.2, 3, dict(conf_thres=0.1, iou_thres=0.45, classes=[1, 5])),
    "dense_no_merge": (1, 16384, 192, 0.9, 4, dict(conf_thres=0.05, iou_thres=0.3)),
    "cap_30000": (1, 49152, 1024, 0.2, 5, dict(conf_thres=0.001, iou_thres=0.6, multi_label=True)),
    "empty": (2, 512, 256, 0.0, 6, dict(conf_thres=0.25, iou_thres=0.45)),
    # autolabelling: apriori label boxes join the candidates with confidence 1 (reference general.py:451-458); image 1 has none
    "apriori_labels": (3, 2048, 256, 0.2, 7, dict(conf_thres=0.25, iou_thres=0.45, labels_seed=9)),
}


def nms_kwargs(kw, B, img):
    """NMS_CASES kwargs with ``labels_seed`` expanded to the reference's ``labels`` argument: per image a float32 array
    [n, 5] of (class, cx, cy, w, h) in pixels, deterministic in the seed (image 1 gets no labels)."""
    import numpy as np
    kw = dict(kw)
    seed = kw.pop("labels_seed", None)
    if seed is not None:
        r = np.random.RandomState(seed)
        labels = []
        for i in range(B):
            n = 0 if i == 1 else int(r.randint(3, 9))
            cls = r.randint(0, 8, size=(n, 1)).astype(np.float32)
            cxy = r.uniform(0.1 * img, 0.9 * img, size=(n, 2)).astype(np.float32)
            wh = r.uniform(0.03 * img, 0.25 * img, size=(n, 2)).astype(np.float32)
            labels.append(np.concatenate([cls, cxy, wh], 1))
        kw["labels"] = labels
    return kw

DETECT_ANCHORS = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119]]
DETECT_STRIDES = (4.0, 8.0)
DETECT_FEATS = ((2, 16, 6, 10), (2, 24, 3, 5))


def swin_state_shapes(dim, ws, linear_mlp, heads, res=None):
    """state_dict float entries of a reference SwinTransformerBlock (names + shapes),
    as listed in SURVEY.md section 3.2; used to rebuild the deterministic weights."""
    ws_eff = min(ws, min(res)) if res is not None else ws
    shapes = {
        "norm1.weight": (dim,), "norm1.bias": (dim,),
        "attn.relative_position_bias_table": ((2 * ws_eff - 1) ** 2, heads),
        "attn.qkv.weight": (3 * dim, dim), "attn.qkv.bias": (3 * dim,),
        "attn.proj.weight": (dim, dim), "attn.proj.bias": (dim,),
        "norm2.weight": (dim,), "norm2.bias": (dim,),
    }
    if linear_mlp:
        shapes.update({"mlp.fc1.weight": (4 * dim, dim), "mlp.fc1.bias": (4 * dim,),
                       "mlp.fc2.weight": (dim, 4 * dim), "mlp.fc2.bias": (dim,)})
    else:
        shapes.update({"mlp.fc1.weight": (dim, dim), "mlp.fc1.bias": (dim,),
                       "mlp.conv1.weight": (dim, dim, 2, 2), "mlp.conv1.bias": (dim,),
                       "mlp.fc2.weight": (dim, dim), "mlp.fc2.bias": (dim,)})
    return shapes


# Round 2 fixtures (tests/golden/swin_blocks_big.npz, variants.npz).  name -> (dim, (H, W), heads, ws, shift, linear_mlp, B)
SWIN_BIG_CASES = {
    "s1_dim192_shift0_lin": (192, (16, 16), 12, 8, 0, True, 2),
    "s1_dim192_shift2_conv": (192, (16, 16), 12, 8, 2, False, 2),
    "s2_dim384_shift2_conv": (384, (16, 16), 12, 8, 2, False, 1),
    "s2_dim384_shift0_lin": (384, (16, 8), 12, 8, 0, True, 2),
    "s3_dim768_global": (768, (32, 32), 12, 32, 0, True, 1),
}
# name -> (dim, ws, heads, B_, masked)
V2ATTN_CASES = {"v2attn_ws8": (96, 8, 6, 4, False), "v2attn_ws4_mask": (48, 4, 3, 8, True), "v2attn_ws7": (64, 7, 2, 3, False)}
# name -> (dim, heads, S, B, use_rel_pos)
SAM_CASES = {"sam_rel_16": (64, 4, 16, 2, True), "sam_norel_8": (48, 3, 8, 2, False), "sam_rel_32": (128, 2, 32, 1, True)}


def v2attn_state_shapes(dim, ws, heads):
    return {"logit_scale": (heads, 1, 1), "cpb_mlp.0.weight": (512, 2), "cpb_mlp.0.bias": (512,), "cpb_mlp.2.weight": (heads, 512),
            "qkv.weight": (3 * dim, dim), "q_bias": (dim,), "v_bias": (dim,), "proj.weight": (dim, dim), "proj.bias": (dim,)}


def sam_state_shapes(dim, heads, S, rel):
    shapes = {"qkv.weight": (3 * dim, dim), "qkv.bias": (3 * dim,), "proj.weight": (dim, dim), "proj.bias": (dim,)}
    if rel:
        shapes.update({"rel_pos_h": (2 * S - 1, dim // heads), "rel_pos_w": (2 * S - 1, dim // heads)})
    return shapes


MF_SHAPES = {"mask_map_r.weight": (1, 3, 1, 1), "mask_map_r.bias": (1,), "mask_map_i.weight": (1, 1, 1, 1), "mask_map_i.bias": (1,),
             "bottleneck1.weight": (16, 1, 3, 3), "bottleneck2.weight": (48, 3, 3, 3), "se.fc.0.weight": (4, 64), "se.fc.2.weight": (64, 4),
             "se_r.fc.0.weight": (1, 3), "se_r.fc.2.weight": (3, 1), "se_i.fc.0.weight": (1, 1), "se_i.fc.2.weight": (1, 1)}
