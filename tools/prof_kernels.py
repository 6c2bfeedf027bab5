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
