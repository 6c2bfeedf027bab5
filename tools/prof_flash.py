#!/usr/bin/env python
"""Times the flash window-attention kernel (stage 3: 32x32 windows, head_dim 64) at the benchmark geometry."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sodt_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
qkv = torch.randn(B, 64, 64, 3 * 768, device=dev, generator=g).to(torch.bfloat16)
table = 0.02 * torch.randn(63 * 63, 12, device=dev, generator=g)
for _ in range(3):
    o = ops.window_attention(qkv, table, 12, 32, 0)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.window_attention(qkv, table, 12, 32, 0)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
flops = 4.0 * B * 4 * 12 * 1024 * 1024 * 64
print(f"flash B={B}: {ms:.3f} ms, {flops / ms / 1e9:.0f} TFLOP/s")
