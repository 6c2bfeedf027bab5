#!/usr/bin/env python
"""The conv-shaped GEMMs of the benchmark step (conv-MLP 2x2 taps of stages 1 / 2, the head's 3x3 convs) at the benchmark
geometry: numerics against torch's float64 conv on one image, then timing.  Run once as is (halo tiles: the kw taps of a
kernel row share one TMA box) and once with SODT_CONV_HALO=0 (one box per tap)."""
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


print("mode:", "one box per tap" if os.environ.get("SODT_CONV_HALO") == "0" else "halo tiles")
for H, Cin, Cout, k, pad, act in ((256, 192, 192, (2, 2), (0, 0), "gelu"), (128, 384, 384, (2, 2), (0, 0), "gelu"),
                                  (256, 64, 64, (3, 3), (1, 1), "silu"), (128, 128, 128, (3, 3), (1, 1), "silu")):
    x = torch.randn(B, H, H, Cin, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, *k, device=dev, generator=g) / (Cin * k[0] * k[1]) ** 0.5).to(torch.bfloat16)
    b = 0.1 * torch.randn(Cout, device=dev, generator=g)
    wt = ops.conv_weight_taps(w)
    fn = lambda: ops.conv2d_nhwc(x, wt, b, k, pad, act)
    out = fn()
    errs = []
    for bi in (0, B - 1):
        xin = torch.nn.functional.pad(x[bi:bi + 1].double().permute(0, 3, 1, 2), (pad[1], k[1] - 1 - pad[1], pad[0], k[0] - 1 - pad[0]))
        ref = torch.nn.functional.conv2d(xin, w.double(), b.double())
        ref = (torch.nn.functional.silu(ref) if act == "silu" else torch.nn.functional.gelu(ref)).permute(0, 2, 3, 1)
        errs.append(((out[bi:bi + 1].double() - ref).norm() / ref.norm()).item())
    t = timed(fn)
    fl = 2.0 * B * H * H * Cin * Cout * k[0] * k[1]
    print(f"conv {k[0]}x{k[1]} H={H} {Cin}->{Cout}: rel err {max(errs):.2e}  {t:.3f} ms  {fl / t / 1e9:.0f} TFLOP/s")
    del x, out
