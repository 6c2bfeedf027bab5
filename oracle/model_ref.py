"""Oracle (test infrastructure): whole RGB+IR detector forward as a pure function of a
state_dict.  CPU, torch functional ops.  Restates

    Model.forward / forward_once   basics/models/model.py:151-296
    ImageEncoderViT.forward        basics/models/backbone_vit.py:190-272
    PatchMerging.forward           basics/models/backbone_vit.py:839-860
    Conv / Bottleneck / C3         basics/models/common.py:38-127
    Detect.forward (eval)          basics/models/model.py:48-65

for the only runnable config, models/model.yaml.  Unlike the reference (which hard-codes
a 128x128 token grid, backbone_vit.py:119,136,153) the token grid follows the input
size -- this is the "reference classes re-instantiated at a scaled input_resolution"
adapter of SURVEY.md section 8c; at 512x512 it is pinned against the unmodified
reference by tests/golden/model_512.npz.
"""
import math

import torch
import torch.nn.functional as F

from . import attention_ref as A
from .detect_ref import detect_decode

STAGE_SHIFTS = (0, 2, 0, 2, 0, 2, 0, 2)  # backbone_vit.py:114
BN_EPS = 1e-3  # utils/torch_utils.py:151 (initialize_weights)

# models/model.yaml:65-74 after parse_model (model.py:350-435) with depth 0.33 / width 0.5
HEAD_ROWS = (
    (2, "Conv", (1, 1)),
    (-1, "Upsample", 2),
    ((-1, 1), "Concat", 1),
    (-1, "C3", 1),
    (-1, "Conv", (1, 1)),
    (-1, "Upsample", 2),
    ((-1, 0), "Concat", 1),
    (-1, "C3", 1),
    ((10,), "Detect", None),
)


def _conv_bn_silu(x, p, pre, k, dtype):
    """common.Conv: conv(no bias, autopad) -> BatchNorm(eval) -> SiLU.  common.py:38-50."""
    w = p[pre + "conv.weight"].to(dtype)
    y = F.conv2d(x, w, None, 1, k // 2)
    y = F.batch_norm(y, p[pre + "bn.running_mean"].to(dtype), p[pre + "bn.running_var"].to(dtype),
                     p[pre + "bn.weight"].to(dtype), p[pre + "bn.bias"].to(dtype), False, 0.0, BN_EPS)
    return F.silu(y)


def _c3(x, p, pre, n, dtype):
    """common.C3 with shortcut=False: cv3(cat(m(cv1 x), cv2 x)).  common.py:114-127, :55-66."""
    a = _conv_bn_silu(x, p, pre + "cv1.", 1, dtype)
    for i in range(n):
        a = _conv_bn_silu(_conv_bn_silu(a, p, f"{pre}m.{i}.cv1.", 1, dtype), p, f"{pre}m.{i}.cv2.", 3, dtype)
    b = _conv_bn_silu(x, p, pre + "cv2.", 1, dtype)
    return _conv_bn_silu(torch.cat((a, b), 1), p, pre + "cv3.", 1, dtype)


def patch_merging(x, p, pre, H, W, dtype):
    """backbone_vit.py:839-860: 2x2 gather in the order (0,0),(1,0),(0,1),(1,1), Linear, LN."""
    B, L, C = x.shape
    g = x.reshape(B, H, W, C)
    g = torch.cat((g[:, 0::2, 0::2], g[:, 1::2, 0::2], g[:, 0::2, 1::2], g[:, 1::2, 1::2]), -1)
    g = F.linear(g.reshape(B, -1, 4 * C), p[pre + "reduction.weight"].to(dtype))
    return F.layer_norm(g, (2 * C,), p[pre + "norm.weight"].to(dtype), p[pre + "norm.bias"].to(dtype), 1e-5)


def backbone_forward(x4, p, heads=12, dtype=torch.float32, pre="image_encoder."):
    """ImageEncoderViT.forward.  x4 [B,4,H,W] -> [y0, y1, y2] NCHW."""
    f = lambda n: p[pre + n].to(dtype)
    x4 = x4.to(dtype)
    streams = []
    for c, name in enumerate("rgbi"):
        pad = 1 if name == "r" else 0  # channel_embed_r keeps PatchEmbed's default padding (1,1): backbone_vit.py:69-74,751
        e = F.conv2d(x4[:, c:c + 1], f(f"channel_embed_{name}.proj.weight"), f(f"channel_embed_{name}.proj.bias"), 4, pad)
        streams.append(e.permute(0, 2, 3, 1))
    ln_w = [f(f"chan_block.norm{i}.weight") for i in range(1, 5)]
    ln_b = [f(f"chan_block.norm{i}.bias") for i in range(1, 5)]
    fused = A.cattention_block(streams, ln_w, ln_b, heads, ws=1, shift=0, dtype=dtype)
    x = torch.cat(fused, -1).permute(0, 3, 1, 2)
    x = F.conv2d(x, f("patch_embed.proj.weight"), f("patch_embed.proj.bias")).permute(0, 2, 3, 1)
    pos = p.get(pre + "pos_embed")
    if pos is not None and x.shape[1] == pos.shape[1]:  # silently skipped otherwise: backbone_vit.py:215-217
        x = x + pos.to(dtype)
    B, h, w, C = x.shape
    x = x.reshape(B, h * w, C)
    kept = []
    for i in range(6):
        x = A.swin_block(x, p, f"{pre}stage1.{i}.", h, w, heads, 8, STAGE_SHIFTS[i], STAGE_SHIFTS[i] == 0, dtype)
        if i >= 4:
            kept.append(x.reshape(B, h, w, C))
    y0 = torch.cat(kept, -1)
    x = patch_merging(x, p, pre + "pmerging1.", h, w, dtype)
    h, w = h // 2, w // 2
    for i in range(4):
        x = A.swin_block(x, p, f"{pre}stage2.{i}.", h, w, heads, 8, STAGE_SHIFTS[i], STAGE_SHIFTS[i] == 0, dtype)
    y1 = x.reshape(B, h, w, -1)
    x = patch_merging(x, p, pre + "pmerging2.", h, w, dtype)
    h, w = h // 2, w // 2
    x = A.swin_block(x, p, f"{pre}stage3.0.", h, w, heads, 32, 0, True, dtype)
    y2 = x.reshape(B, h, w, -1)
    nchw = lambda t: t.permute(0, 3, 1, 2)
    return [F.conv2d(nchw(y0), f("neck1.weight")), F.conv2d(nchw(y1), f("neck2.weight")),
            F.conv2d(nchw(y2), f("neck3.weight"))]


def head_forward(feats, p, stride=4.0, dtype=torch.float32, pre="detect."):
    """The head loop of forward_once (model.py:268-281) + Detect (eval)."""
    y = list(feats)
    x = feats[-1]
    for i, (frm, kind, arg) in enumerate(HEAD_ROWS):
        if frm != -1:
            x = y[frm] if isinstance(frm, int) else [x if j == -1 else y[j] for j in frm]
        if kind == "Conv":
            x = _conv_bn_silu(x, p, f"{pre}{i}.", arg[0], dtype)
        elif kind == "Upsample":
            x = F.interpolate(x, scale_factor=arg, mode="nearest")
        elif kind == "Concat":
            x = torch.cat(x, arg)
        elif kind == "C3":
            x = _c3(x, p, f"{pre}{i}.", arg, dtype)
        elif kind == "Detect":
            raw = F.conv2d(x[0], p[f"{pre}{i}.m.0.weight"].to(dtype), p[f"{pre}{i}.m.0.bias"].to(dtype))
            anchors_px = p[f"{pre}{i}.anchor_grid"].reshape(-1, 2)
            z, xp = detect_decode(raw, anchors_px, stride, dtype)
            return z, [xp]
        y.append(x)
    raise RuntimeError("head has no Detect row")


def model_forward(rgb, ir, p, dtype=torch.float32):
    """Model.forward(x, ir, 'RGB+IR') in eval mode: (pred [B,R,13], [raw [B,3,ny,nx,13]])."""
    x4 = torch.cat((rgb, ir[:, 0:1]), 1)  # model.py:192
    feats = backbone_forward(x4, p, dtype=dtype)
    return head_forward(feats, p, dtype=dtype)
