#!/usr/bin/env python
"""Fused MLP kernel (ops.mlp_ln) vs the fc1 + fc2 GEMM pair at the stage-1 shape: numerics on a small and the full M, then
timing.  `python tools/prof_mlp.py [rows]`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sodt_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
C, HID = 192, 768
gam, bet = 1.0 + 0.1 * torch.randn(C, device=dev, generator=g), 0.1 * torch.randn(C, device=dev, generator=g)
w1 = (torch.randn(HID, C, device=dev, generator=g) / 14).to(torch.bfloat16)
b1 = 0.1 * torch.randn(HID, device=dev, generator=g)
w2 = (torch.randn(C, HID, device=dev, generator=g) / 28).to(torch.bfloat16)
b2 = 0.1 * torch.randn(C, device=dev, generator=g)


def pair(x, st):
    h = ops.linear(x, w1, b1, act="gelu", ln=(st, gam, bet, 1e-5))
    return ops.linear(h, w2, b2, residual=x, want_stats=True)


def fused(x, st, f16=True):
    return ops.mlp_ln(x, (st, gam, bet, 1e-5), w1, b1, w2, b2, want_stats=True, hidden_fp16=f16)


def check(M):
    x = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
    st = ops.row_stats(x, 1e-5)
    o, so = fused(x, st)
    ob, _ = fused(x, st, False)
    torch.cuda.synchronize()
    xd = x.double()
    y = torch.nn.functional.layer_norm(xd, (C,), gam.double(), bet.double(), 1e-5)
    ref = xd + torch.nn.functional.gelu(y @ w1.double().t() + b1.double()) @ w2.double().t() + b2.double()
    err = ((o.double() - ref).norm() / ref.norm()).item()
    op, _ = pair(x, st)
    errp = ((op.double() - ref).norm() / ref.norm()).item()
    s = so.sum(0)
    serr = ((s[:, 0] - ref.sum(1).float()).abs().max() / ref.sum(1).abs().max()).item()
    errb = ((ob.double() - ref).norm() / ref.norm()).item()
    print(f"M={M}: fused rel err fp16-hidden {err:.3e} bf16-hidden {errb:.3e} (GEMM pair {errp:.3e}); stats sum err {serr:.2e}; max abs {float((o.double()-ref).abs().max()):.3e}")


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


for M in (128, 200, 4096 + 64, 148 * 128 * 3 + 5):
    check(M)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32 * 256 * 256
x = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
st = ops.row_stats(x, 1e-5)
o, _ = fused(x, st, False)
op, _ = pair(x, st)
print("full M: fused (bf16 hidden) vs pair rel diff", ((o.float() - op.float()).norm() / op.float().norm()).item())
print(f"M={M}: GEMM pair {timed(lambda: pair(x, st)):.3f} ms, fused bf16-hidden {timed(lambda: fused(x, st, False)):.3f} ms, "
      f"fp16-hidden {timed(lambda: fused(x, st, True)):.3f} ms")
