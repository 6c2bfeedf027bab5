#!/usr/bin/env python
"""clock64 timeline of CTA 0 of the GEMM kernel (library built with -DLIN_TRACE: tools/build_variant.sh LIN_TRACE; run with
SODT_B200_LIB=<that library>).  Usage: trace_linear.py qkv | proj | fc1s2"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sodt_b200 import _capi, ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
what = sys.argv[1] if len(sys.argv) > 1 else "qkv"
M, N, K, kind = {"qkv": (2097152, 576, 192, "ln"), "proj": (2097152, 192, 192, "res"), "fc1s2": (524288, 1536, 384, "ln"),
                 "qkvs2": (524288, 1152, 384, "ln")}[what]
x = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
W = (torch.randn(N, K, device=dev, generator=g) / K ** 0.5).to(torch.bfloat16)
b = 0.1 * torch.randn(N, device=dev, generator=g)
gam, bet = torch.ones(K, device=dev), torch.zeros(K, device=dev)
if kind == "ln":
    st = ops.row_stats(x, 1e-5)
    fn = lambda: ops.linear(x, W, b, ln=(st, gam, bet, 1e-5))
else:
    r = torch.randn(M, N, device=dev, generator=g).to(torch.bfloat16)
    fn = lambda: ops.linear(x, W, b, residual=r, want_stats=True)
for _ in range(3):
    fn()
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); fn(); e.record(); torch.cuda.synchronize()
print(f"{what}: M={M} N={N} K={K}: {a.elapsed_time(e):.3f} ms")
buf = (ctypes.c_longlong * (6 * 64 * 16))()
h = ctypes.CDLL(_capi.LIB_PATH)
h.sodt_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
print("rc", h.sodt_debug_trace(buf, ctypes.sizeof(buf)))
tr = torch.tensor(list(buf)).view(6, 64, 16)
nkb = K // 64
t0 = int(tr[0, 8, 0])
print("tile | MMA: start acc_empty-ok kb-full... commit | producer slot-free per kb | per epilogue group: wait-start acc_full [box: wait_read-start wait_read-done store]...")
for i in range(8, 28):
    m = " ".join(f"{int(tr[0, i, k]) - t0:7d}" for k in [0, 1] + list(range(2, 2 + nkb)) + [15])
    pr = " ".join(f"{int(tr[1, i, k]) - t0:7d}" for k in range(nkb))
    eg = []
    for gidx in range(4):
        vals = [int(tr[2 + gidx, i, k]) for k in range(14)]
        eg.append(" ".join(f"{v - t0:7d}" if v else "      -" for v in vals[:2 + 3 * 4]))
    print(f"{i:3d} | {m} | {pr}")
    for gidx in range(4):
        print(f"      g{gidx}: {eg[gidx]}")
