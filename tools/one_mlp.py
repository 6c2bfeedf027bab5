#!/usr/bin/env python
"""Launches the fused MLP kernel a few times at the stage-1 shape (target of an ncu capture)."""
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
    ops.mlp_ln(x, (st, gam, bet, 1e-5), w1, b1, w2, b2, want_stats=True)
torch.cuda.synchronize()
print("ok")
