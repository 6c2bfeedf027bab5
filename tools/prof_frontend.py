#!/usr/bin/env python
"""One-kernel front end (ops.frontend_embed_u8) vs frontend_u8 + embedding GEMM at the benchmark geometry (B = 32, 1024^2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sodt_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
B, H, W, E, D = (int(sys.argv[1]) if len(sys.argv) > 1 else 32), 1024, 1024, 48, 192
rgb = torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8, device=dev)
ir = torch.randint(0, 256, (B, 1, H, W), generator=g, dtype=torch.uint8, device=dev)
bf = lambda t: t.to(torch.bfloat16).float()
cw, cb = bf(torch.randn(4, E, 16, device=dev, generator=g) / 4), bf(0.1 * torch.randn(4, E, device=dev, generator=g))
lw, lb = bf(1 + 0.1 * torch.randn(4, E, device=dev, generator=g)), bf(0.1 * torch.randn(4, E, device=dev, generator=g))
pw = (torch.randn(D, 4 * E, device=dev, generator=g) / 14).to(torch.bfloat16)
pb = bf(0.1 * torch.randn(D, device=dev, generator=g))
pos = (0.5 * torch.randn(1, H // 4, W // 4, D, device=dev, generator=g)).to(torch.bfloat16)


def two():
    cat = ops.frontend_u8(rgb, ir, cw, cb, lw, lb, torch.bfloat16, pad_r=1, eps=1e-6)
    return ops.linear(cat, pw, pb, residual=pos, want_stats=True)


def one():
    return ops.frontend_embed_u8(rgb, ir, cw, cb, lw, lb, pw, pb, pos, pad_r=1, eps=1e-6, want_stats=True)


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


o1, s1 = one()
o2, s2 = two()
torch.cuda.synchronize()
print("rel diff one vs two kernels:", ((o1.float() - o2.float()).norm() / o2.float().norm()).item(),
      "stats:", ((s1.sum(0) - s2.sum(0)).abs().max() / s2.sum(0).abs().max()).item())
print(f"two kernels {timed(two):.3f} ms, one kernel {timed(one):.3f} ms")
