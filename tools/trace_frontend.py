#!/usr/bin/env python
"""Prints the clock64 timeline CTA 0 of the one-kernel front end recorded (library built with -DFE_TRACE, SODT_B200_LIB=...)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sodt_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
B, H, W, E, D = 32, 1024, 1024, 48, 192
rgb = torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8, device=dev)
ir = torch.randint(0, 256, (B, 1, H, W), generator=g, dtype=torch.uint8, device=dev)
cw, cb = torch.randn(4, E, 16, device=dev, generator=g) / 4, 0.1 * torch.randn(4, E, device=dev, generator=g)
lw, lb = 1 + 0.1 * torch.randn(4, E, device=dev, generator=g), 0.1 * torch.randn(4, E, device=dev, generator=g)
pw = (torch.randn(D, 4 * E, device=dev, generator=g) / 14).to(torch.bfloat16)
pb = 0.1 * torch.randn(D, device=dev, generator=g)
pos = (0.5 * torch.randn(1, H // 4, W // 4, D, device=dev, generator=g)).to(torch.bfloat16)
for _ in range(3):
    o, st = ops.frontend_embed_u8(rgb, ir, cw, cb, lw, lb, pw, pb, pos, pad_r=1, eps=1e-6, want_stats=True)
torch.cuda.synchronize()
tr = st.view(-1)[: 5 * 32 * 8 * 2].view(torch.int64).view(5, 32, 8).cpu()
t0 = int(tr[1, 0, 0])
print("tile | MMA: a_full conv_empty cat_full acc_empty | loader: loaded arrive | epi1 g0: conv_full stats0 cat_empty stats1 . done | epi1 g1 | epi2: start acc_full pos0 pos1 pos2 box01-done box2-done")
for i in range(8, 24):
    f = lambda r, ks: " ".join(f"{int(tr[r, i, k]) - t0:7d}" for k in ks)
    print(f"{i:3d} | {f(0, (0, 1, 2, 3))} | {f(1, (0, 1))} | {f(2, (0, 1, 2, 3, 5))} | {f(3, (0, 1, 2, 3, 5))} | {f(4, (0, 1, 2, 3, 4, 5, 6))}")
