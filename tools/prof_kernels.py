#!/usr/bin/env python
"""Runs each sodt kernel a few times at the BASELINE config-2 geometry (batch 32, 1024x1024 input, bf16)
so that `ncu --set full -k regex:...` can capture it without paying for the whole model.
Usage: python tools/prof_kernels.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures as fx  # noqa: E402  (synthetic inputs only)
from sodt_b200 import ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
B = 32


def qkv(h, C):
    return torch.randn(B, h, h, 3 * C, device=dev, generator=g).to(torch.bfloat16)


cases = [("stage1", qkv(256, 192), 12, 8), ("stage2", qkv(128, 384), 12, 8), ("stage3", qkv(64, 768), 12, 32)]
for name, t, heads, ws in cases:
    table = 0.02 * torch.randn((2 * ws - 1) ** 2, heads, device=dev, generator=g)
    for shift in ((0, 2) if ws == 8 else (0,)):
        for _ in range(reps):
            ops.window_attention(t, table, heads, ws, shift)
    torch.cuda.synchronize()
    del t
# GEMM-shaped layers (stage-1 geometry): qkv, proj + residual, fc1 + GELU, fc2 + residual, conv-MLP taps, PatchMerging
M = B * 256 * 256
xt = torch.randn(M, 192, device=dev, generator=g).to(torch.bfloat16)
w = lambda n, k: (torch.randn(n, k, device=dev, generator=g) / k ** 0.5).to(torch.bfloat16)
bias = lambda n: 0.1 * torch.randn(n, device=dev, generator=g)
gam, bet = 1.0 + 0.1 * torch.randn(192, device=dev, generator=g), 0.1 * torch.randn(192, device=dev, generator=g)
wq, bq, w1, b1 = w(576, 192), bias(576), w(768, 192), bias(768)
for _ in range(reps):
    x1, part = ops.linear(xt, w(192, 192), bias(192), residual=xt, want_stats=True)      # proj + residual, emits row statistics
    mr = ops.finalize_stats(part, 192, 1e-5)
    ops.linear(x1, wq, bq, ln=(mr, gam, bet))                                                 # norm1 + qkv
    hid = ops.linear(x1, w1, b1, act="gelu", ln=(mr, gam, bet))                               # norm2 + fc1 + GELU
    ops.linear(hid, w(192, 768), bias(192), residual=xt, want_stats=True)                # fc2 + residual
    ops.mlp_ln(x1, (mr, gam, bet, 1e-5), w1, b1, w(192, 768), bias(192), want_stats=True)       # the same MLP half as one kernel
    tb = 0.02 * torch.randn(225, 12, device=dev, generator=g)
    for sh in (0, 2):         # norm1 + qkv + window attention as one kernel
        ops.attn_block(x1.view(B, 256, 256, 192), (mr, gam, bet, 1e-5), wq, bq, tb, 12, 8, sh)
    ops.conv2d_nhwc(xt.view(B, 256, 256, 192), w(192, 4 * 192), bias(192), (2, 2), (0, 0), "gelu")
    ops.patch_merge_linear(xt.view(B, 256, 256, 192), w(384, 768))
    ops.add_layernorm(xt, None, torch.ones(192, device=dev), torch.zeros(192, device=dev), 1e-5)
    ops.row_stats(xt, 1e-5)
    del hid, x1, part, mr
del xt
# stage-3 GEMMs (K >= 768: CTA pairs, cta_group::2)
x3 = torch.randn(B * 64 * 64, 768, device=dev, generator=g).to(torch.bfloat16)
g3, b3 = 1.0 + 0.1 * torch.randn(768, device=dev, generator=g), 0.1 * torch.randn(768, device=dev, generator=g)
mr3 = ops.row_stats(x3, 1e-5)
for _ in range(reps):
    h3 = ops.linear(x3, w(3072, 768), bias(3072), act="gelu", ln=(mr3, g3, b3))
    ops.linear(h3, w(768, 3072), bias(768), residual=x3, want_stats=True)
    del h3
del x3
# head: 1x1 conv over cat(up2x(low), skip) through zero-stride TMA dimensions
low = torch.randn(B, 128, 128, 128, device=dev, generator=g).to(torch.bfloat16)
skip = torch.randn(B, 256, 256, 256, device=dev, generator=g).to(torch.bfloat16)
for _ in range(reps):
    ops.upcat_conv1x1(ops.UpCat(low.permute(0, 3, 1, 2), skip.permute(0, 3, 1, 2)), w(128, 384), bias(128), "silu")
del low, skip
# conv-shaped GEMMs on halo tiles: head 3x3 (64 -> 64 at 256^2) and the stage-2 conv-MLP taps (2x2, 384 -> 384 at 128^2)
xc = torch.randn(B, 256, 256, 64, device=dev, generator=g).to(torch.bfloat16)
xc2 = torch.randn(B, 128, 128, 384, device=dev, generator=g).to(torch.bfloat16)
for _ in range(reps):
    ops.conv2d_nhwc(xc, w(64, 9 * 64), bias(64), (3, 3), (1, 1), "silu")
    ops.conv2d_nhwc(xc2, w(384, 4 * 384), bias(384), (2, 2), (0, 0), "gelu")
del xc, xc2
img = torch.rand(B, 4, 1024, 1024, device=dev, generator=g).to(torch.bfloat16)
cw, cb = 0.3 * torch.randn(4, 48, 16, device=dev, generator=g), 0.1 * torch.randn(4, 48, device=dev, generator=g)
for _ in range(reps):
    ops.frontend(img, cw, cb, torch.ones(4, 48, device=dev), torch.zeros(4, 48, device=dev), pad_r=1, eps=1e-5)
del img
rgb = torch.randint(0, 256, (B, 3, 1024, 1024), generator=g, dtype=torch.uint8, device=dev)
ir = torch.randint(0, 256, (B, 1, 1024, 1024), generator=g, dtype=torch.uint8, device=dev)
pos = (0.5 * torch.randn(1, 256, 256, 192, device=dev, generator=g)).to(torch.bfloat16)
for _ in range(reps):       # uint8 images -> patch-embedded tokens in one kernel
    ops.frontend_embed_u8(rgb, ir, cw, cb, torch.ones(4, 48, device=dev), torch.zeros(4, 48, device=dev), w(192, 192), bias(192), pos,
                          pad_r=1, eps=1e-6, want_stats=True)
del rgb, ir, pos
torch.cuda.synchronize()
streams = [torch.randn(B, 48, 256, 256, device=dev, generator=g).to(torch.bfloat16).permute(0, 2, 3, 1) for _ in range(4)]
ln_w, ln_b = torch.ones(4, 48, device=dev), torch.zeros(4, 48, device=dev)
for _ in range(reps):
    ops.cattn_block(*streams, ln_w, ln_b, 12)
del streams
raw = torch.randn(B, 39, 256, 256, device=dev, generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
anchors = torch.tensor([[10., 13.], [16., 30.], [33., 23.]], device=dev)
for _ in range(reps):
    ops.detect_decode(raw, anchors, 4.0)
pred = torch.from_numpy(fx.synthetic_predictions(B, 196608, 8, 1024, 0.02, 0)).to(dev)
for _ in range(reps):
    ops.nms(pred, 0.25, 0.45)
    ops.nms(pred, 0.001, 0.6, multi_label=True)
torch.cuda.synchronize()
print("ok")
