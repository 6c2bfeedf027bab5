#!/usr/bin/env python
"""Stage-1 / stage-2 geometry window attention only (for ncu).  Usage: prof_win8.py [reps] [which: 1|2|3|all]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sodt_b200 import ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
which = sys.argv[2] if len(sys.argv) > 2 else "all"
g = torch.Generator(device="cuda").manual_seed(0)
B = 32
cases = {"1": (256, 192, 12, 8, 2), "1s0": (256, 192, 12, 8, 0), "2": (128, 384, 12, 8, 2), "2s0": (128, 384, 12, 8, 0), "3": (64, 768, 12, 32, 0)}
for key, (h, C, heads, ws, shift) in cases.items():
    if which not in ("all", key):
        continue
    qkv = torch.randn(B, h, h, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
    table = 0.02 * torch.randn((2 * ws - 1) ** 2, heads, device="cuda", generator=g)
    for _ in range(2):
        ops.window_attention(qkv, table, heads, ws, shift)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.window_attention(qkv, table, heads, ws, shift)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    nbytes = B * h * h * 4 * C * 2
    flops = 4 * (ws * ws) * C * B * h * h
    print(f"case {key}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s  {flops / ms / 1e9:.0f} TFLOP/s")
    del qkv
