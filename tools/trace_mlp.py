#!/usr/bin/env python
"""Prints the clock64 timeline CTA 0 of the fused MLP kernel recorded (library built with -DMLP_TRACE, SODT_B200_LIB=...)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sodt_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
C, HID, M = 192, 768, 32 * 256 * 256
gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
w1 = (torch.randn(HID, C, device=dev, generator=g) / 14).to(torch.bfloat16)
b1 = 0.1 * torch.randn(HID, device=dev, generator=g)
w2 = (torch.randn(C, HID, device=dev, generator=g) / 28).to(torch.bfloat16)
b2 = 0.1 * torch.randn(C, device=dev, generator=g)
x = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
st = ops.row_stats(x, 1e-5)
for _ in range(3):
    o, so = ops.mlp_ln(x, (st, gam, bet, 1e-5), w1, b1, w2, b2, want_stats=True)
torch.cuda.synchronize()
tr = so.view(-1)[: 3 * 64 * 8 * 2].view(torch.int64).view(3, 64, 8).cpu()
t0 = int(tr[0, 0, 0])
NCH = HID // 128
print("chunk | MMA: wait-start h_ready o_empty g2-issued g1-issued | epi w0: h_full ld-done math-done st-done o-done | epi w15: same   (cycles since start)")
for q in range(12, 36):
    m = [int(v) - t0 for v in tr[0, q, :5]]
    e0 = [int(tr[1, q, k]) - t0 for k in (1, 6, 7, 3, 5)]
    e1 = [int(tr[2, q, k]) - t0 for k in (1, 6, 7, 3, 5)]
    print(f"{q:3d} t{q // NCH} c{q % NCH} | {m[0]:7d} {m[1]:7d} {m[2]:7d} {m[3]:7d} {m[4]:7d} | {e0[0]:7d} {e0[1]:7d} {e0[2]:7d} {e0[3]:7d} {e0[4]:7d} | {e1[0]:7d} {e1[1]:7d} {e1[2]:7d} {e1[3]:7d} {e1[4]:7d}")
