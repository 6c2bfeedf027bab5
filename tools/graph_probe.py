#!/usr/bin/env python
"""Measures host enqueue time vs device time of one detector step, and the same step replayed from a CUDA graph."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sodt_b200.runtime import Detector  # noqa: E402

B, S = 32, 1024
det = Detector(device="cuda", dtype=torch.bfloat16, seed=0)
g = torch.Generator().manual_seed(0)
rgb = torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=g).cuda()
ir = torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=g).cuda()
for _ in range(3):
    det.detect_device(rgb, ir)
torch.cuda.synchronize()
n = 10
t0 = time.perf_counter()
for _ in range(n):
    det.detect_device(rgb, ir)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"eager: host enqueue {1e3 * (t1 - t0) / n:.1f} ms/step, total {1e3 * (t2 - t0) / n:.1f} ms/step")

buf = det.buffer(B)
ref = buf.flat.clone()
gr = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    det.detect_device(rgb, ir)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(gr):
    det.detect_device(rgb, ir, buf)
torch.cuda.synchronize()
buf.flat.zero_()
gr.replay()
torch.cuda.synchronize()
print("graph result equals eager:", torch.equal(buf.flat, ref))
t0 = time.perf_counter()
for _ in range(n):
    gr.replay()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"graph: host enqueue {1e3 * (t1 - t0) / n:.1f} ms/step, total {1e3 * (t2 - t0) / n:.1f} ms/step")

# ---- host link: pinned upload bandwidth, alone and under load
host = torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=g).pin_memory()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    d = host.to("cuda", non_blocking=True)
torch.cuda.synchronize()
a.record()
for _ in range(5):
    d = host.to("cuda", non_blocking=True)
b.record()
torch.cuda.synchronize()
print(f"H2D pinned alone: {5 * host.numel() / a.elapsed_time(b) / 1e6:.1f} GB/s")
cs = torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(cs):
    a.record(cs)
    for _ in range(5):
        d = host.to("cuda", non_blocking=True)
    b.record(cs)
for _ in range(5):
    gr.replay()
torch.cuda.synchronize()
print(f"H2D pinned while the model runs: {5 * host.numel() / a.elapsed_time(b) / 1e6:.1f} GB/s; 5 steps + copies wall {1e3 * (time.perf_counter() - t0):.1f} ms")

# ---- the public pipelined API
hosts = [(torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=g).pin_memory(),
          torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=g).pin_memory()) for _ in range(2)]
for n in (4, 10, 20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in det.detect_stream(hosts[i % 2] for i in range(n)):
        pass
    torch.cuda.synchronize()
    print(f"detect_stream {n} steps: {1e3 * (time.perf_counter() - t0) / n:.1f} ms/step")
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    det.detect_device(hosts[i % 2][0].to("cuda", non_blocking=True), hosts[i % 2][1].to("cuda", non_blocking=True))
torch.cuda.synchronize()
print(f"serial upload + detect_device: {1e3 * (time.perf_counter() - t0) / 10:.1f} ms/step")
