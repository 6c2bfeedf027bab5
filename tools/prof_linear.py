#!/usr/bin/env python
"""Times the GEMM kernel's variants at the stage-1 shapes (M = 2.1 M tokens): plain / bias / LayerNorm fold / row statistics.
`python tools/prof_linear.py one N` launches a single bias-free GEMM of width N (for an ncu source-level capture)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sodt_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
M = 32 * 256 * 256
x = torch.randn(M, 192, device=dev, generator=g).to(torch.bfloat16)
gam, bet = 1.0 + 0.1 * torch.randn(192, device=dev, generator=g), 0.1 * torch.randn(192, device=dev, generator=g)
mr = ops.row_stats(x, 1e-5)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


if len(sys.argv) > 2 and sys.argv[1] == "one":
    N = int(sys.argv[2])
    w = (torch.randn(N, 192, device=dev, generator=g) / 14).to(torch.bfloat16)
    ops.linear(x, w, None)
    torch.cuda.synchronize()
    print("ok")
    sys.exit(0)
for N, act in ((576, None), (768, "gelu"), (192, None)):
    w = (torch.randn(N, 192, device=dev, generator=g) / 14).to(torch.bfloat16)
    b = 0.1 * torch.randn(N, device=dev, generator=g)
    res = [("nobias", timed(lambda: ops.linear(x, w, None, act=act))), ("bias", timed(lambda: ops.linear(x, w, b, act=act))),
           ("ln", timed(lambda: ops.linear(x, w, b, act=act, ln=(mr, gam, bet)))),
           ("bias+stats", timed(lambda: ops.linear(x, w, b, act=act, want_stats=True)))]
    print(f"N={N} act={act}: " + "  ".join(f"{k} {v:.3f}" for k, v in res))
