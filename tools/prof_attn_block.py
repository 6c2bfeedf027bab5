#!/usr/bin/env python
"""The fused attention half (norm1 + qkv + window attention, sodt_attn_block_fwd) against the two kernels it replaces at the
benchmark geometry (B = 32, 256 x 256 tokens, C = 192, 12 heads): bit equality, then timing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sodt_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
B, H, C, heads, ws = 32, 256, 192, 12, 8


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


M = B * H * H
a = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
wp = (torch.randn(C, C, device=dev, generator=g) / C ** 0.5).to(torch.bfloat16)
x, st = ops.linear(a, wp, None, residual=a, want_stats=True)
x = x.view(B, H, H, C)
wq = (torch.randn(3 * C, C, device=dev, generator=g) / C ** 0.5).to(torch.bfloat16)
bq = (0.2 * torch.randn(3 * C, device=dev, generator=g)).to(torch.bfloat16)
gam = (1.0 + 0.2 * torch.randn(C, device=dev, generator=g)).to(torch.bfloat16)
bet = (0.1 * torch.randn(C, device=dev, generator=g)).to(torch.bfloat16)
table = 0.5 * torch.randn((2 * ws - 1) ** 2, heads, device=dev, generator=g)
ln = (st, gam, bet, 1e-5)
for shift in (0, 2):
    qkv = ops.linear(x, wq, bq, ln=ln)
    ref = ops.window_attention(qkv, table, heads, ws, shift)
    out = ops.attn_block(x, ln, wq, bq, table, heads, ws, shift)
    torch.cuda.synchronize()
    print(f"shift {shift}: bit-identical {torch.equal(out, ref)}  max abs diff {(out.float() - ref.float()).abs().max().item():.3e}")
    t_q = timed(lambda: ops.linear(x, wq, bq, ln=ln))
    t_a = timed(lambda: ops.window_attention(qkv, table, heads, ws, shift))
    t_f = timed(lambda: ops.attn_block(x, ln, wq, bq, table, heads, ws, shift))
    print(f"   qkv GEMM {t_q:.3f} ms + window attention {t_a:.3f} ms = {t_q + t_a:.3f} ms;  fused {t_f:.3f} ms")
