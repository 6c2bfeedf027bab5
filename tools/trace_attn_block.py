#!/usr/bin/env python
"""clock64 timeline of CTA 0 of the fused attention kernel (library built with -DATTN_TRACE: tools/build_variant.sh attntrace
-DATTN_TRACE; run with SODT_B200_LIB=<that library>).  Usage: trace_attn_block.py [shift]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sodt_b200 import _capi, ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
B, H, C, heads, ws = 32, 256, 192, 12, 8
shift = int(sys.argv[1]) if len(sys.argv) > 1 else 0
M = B * H * H
x = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
st = ops.row_stats(x, 1e-5)
x = x.view(B, H, H, C)
wq = (torch.randn(3 * C, C, device=dev, generator=g) / C ** 0.5).to(torch.bfloat16)
bq = (0.2 * torch.randn(3 * C, device=dev, generator=g)).to(torch.bfloat16)
gam, bet = torch.ones(C, device=dev, dtype=torch.bfloat16), torch.zeros(C, device=dev, dtype=torch.bfloat16)
table = 0.5 * torch.randn((2 * ws - 1) ** 2, heads, device=dev, generator=g)
ln = (st, gam, bet, 1e-5)
for _ in range(3):
    ops.attn_block(x, ln, wq, bq, table, heads, ws, shift)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); ops.attn_block(x, ln, wq, bq, table, heads, ws, shift); e.record(); torch.cuda.synchronize()
print(f"shift {shift}: {a.elapsed_time(e):.3f} ms")
buf = (ctypes.c_longlong * (6 * 64 * 16))()
h = ctypes.CDLL(_capi.LIB_PATH)
h.sodt_attn_trace.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
print("rc", h.sodt_attn_trace(buf, ctypes.sizeof(buf)))
cb = (ctypes.c_longlong * (256 * 2))()
h.sodt_attn_cta_cycles.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
h.sodt_attn_cta_cycles(cb, ctypes.sizeof(cb))
cyc = torch.tensor(list(cb)).view(256, 2)[:148]
cs = cyc[:, 0].float()
print(f"per-CTA cycles: min {cs.min().item():.0f} median {cs.median().item():.0f} max {cs.max().item():.0f}; tiles per CTA {cyc[:, 1].min().item()}-{cyc[:, 1].max().item()}")
print("slowest CTAs:", [(int(i), int(cs[i].item())) for i in cs.argsort(descending=True)[:8]])
tr = torch.tensor(list(buf)).view(6, 64, 16)
t0 = int(tr[0, 6, 0])
f = lambda v: f"{int(v) - t0:7d}" if int(v) else "      -"
print("GEMM per unit u: start d_free-ok w0 w1 w2 commit | (tile start: x-wait x-ok)")
for u in range(6, 24):
    print(f"  u{u:3d} " + " ".join(f(tr[0, u, k]) for k in range(6)) + " | " + " ".join(f(tr[0, u, k]) for k in (6, 7)))
print("attention MMA per window k: ready(QK issue) QK-committed p_full-ok PV-committed   [group 0 | group 1]")
for k in range(6, 24):
    print(f"  k{k:3d} " + " ".join(f(tr[1, k, e_]) for e_ in range(4)) + "  |  " + " ".join(f(tr[2, k, e_]) for e_ in range(4)))
print("softmax group per own unit j: start d_full-ok convert-done | w0: s_full p_full-arrived pv_done store | w1: same   [group 0 / group 1]")
for j in range(3, 12):
    for gi in range(2):
        print(f"  g{gi} j{j:3d} " + " ".join(f(tr[3 + gi, j, e_]) for e_ in range(3)) + " | " + " ".join(f(tr[3 + gi, j, e_]) for e_ in range(3, 7)) +
              " | " + " ".join(f(tr[3 + gi, j, e_]) for e_ in range(7, 11)))
print("producer per tile: x_empty-ok x-issued")
for t in range(2, 8):
    print(f"  t{t:3d} " + " ".join(f(tr[5, t, e_]) for e_ in range(2)))
