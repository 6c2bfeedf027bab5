#!/usr/bin/env python
"""The K >= 384 GEMM-shaped layers of stages 2 / 3 (and the conv taps) at the benchmark geometry: numerics against a float64
product on sampled rows, then timing.  Run once as is (CTA pairs) and once with SODT_NO_CTA2=1 (single-CTA tiles)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sodt_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
B = 32


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


w = lambda n, k: (torch.randn(n, k, device=dev, generator=g) / k ** 0.5).to(torch.bfloat16)
bias = lambda n: 0.1 * torch.randn(n, device=dev, generator=g)
print("mode:", "single-CTA tiles" if os.environ.get("SODT_NO_CTA2") == "1" else "CTA pairs")
for M, N, K, kind in ((524288, 1152, 384, "ln"), (524288, 1536, 384, "ln+gelu"), (524288, 384, 1536, "res+stats"), (524288, 384, 384, "res+stats"),
                      (131072, 2304, 768, "ln"), (131072, 3072, 768, "ln+gelu"), (131072, 768, 3072, "res+stats"), (524288 + 77, 256, 384, "plain")):
    x = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    W, b = w(N, K), bias(N)
    gam, bet = 1.0 + 0.1 * torch.randn(K, device=dev, generator=g), 0.1 * torch.randn(K, device=dev, generator=g)
    idx = torch.randint(0, M, (512,), device=dev, generator=g)
    idx[:4] = torch.tensor([0, 127, 128, M - 1], device=dev)
    xs = x[idx].double()
    if kind.startswith("ln"):
        st = ops.row_stats(x, 1e-5)
        act = "gelu" if "gelu" in kind else None
        fn = lambda: ops.linear(x, W, b, act=act, ln=(st, gam, bet, 1e-5))
        ref = torch.nn.functional.layer_norm(xs, (K,), gam.double(), bet.double(), 1e-5) @ W.double().t() + b.double()
        if act:
            ref = torch.nn.functional.gelu(ref)
        out = fn()
    elif kind == "res+stats":
        r = torch.randn(M, N, device=dev, generator=g).to(torch.bfloat16)
        fn = lambda: ops.linear(x, W, b, residual=r, want_stats=True)
        ref = xs @ W.double().t() + b.double() + r[idx].double()
        out, part = fn()
        s = part.sum(0)[idx, 0].double()
        print("   stats err", float((s - ref.sum(1)).abs().max() / ref.sum(1).abs().max()))
    else:
        fn = lambda: ops.linear(x, W, None)
        ref = xs @ W.double().t()
        out = fn()
    err = ((out[idx].double() - ref).norm() / ref.norm()).item()
    flops = 2.0 * M * N * K
    t = timed(fn)
    print(f"M={M} N={N} K={K} {kind}: rel err {err:.2e}  {t:.3f} ms  {flops / t / 1e9:.0f} TFLOP/s")
    del x, out
for H, C in ((256, 192), (128, 384)):
    x = torch.randn(B, H, H, C, device=dev, generator=g).to(torch.bfloat16)
    wt, b = w(C, 4 * C), bias(C)
    fn = lambda: ops.conv2d_nhwc(x, wt, b, (2, 2), (0, 0), "gelu")
    out = fn()
    bi, yi, xi = 3, H - 1, H - 2
    taps = []
    for ky in range(2):
        for kx in range(2):
            yy, xx = yi + ky, xi + kx
            taps.append(x[bi, yy, xx].double() if yy < H and xx < H else torch.zeros(C, device=dev, dtype=torch.float64))
    ref = torch.nn.functional.gelu(torch.cat(taps) @ wt.double().t() + b.double())
    err = ((out[bi, yi, xi].double() - ref).norm() / ref.norm()).item()
    t = timed(fn)
    print(f"conv2x2 H={H} C={C}: rel err {err:.2e}  {t:.3f} ms  {2.0 * B * H * H * C * 4 * C / t / 1e9:.0f} TFLOP/s")
