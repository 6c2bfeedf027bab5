#!/usr/bin/env python
"""HBM bandwidth by access mix (write-only, read-only, copy) - context for the roofline of write-heavy GEMM epilogues."""
import torch

n = 1 << 30   # 4 GiB fp32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


t = timed(lambda: a.fill_(1.0))
print(f"write-only (fill): {4 * n / t / 1e6:.0f} GB/s")
t = timed(lambda: a.zero_())
print(f"write-only (memset): {4 * n / t / 1e6:.0f} GB/s")
t = timed(lambda: a.sum())
print(f"read-only (sum): {4 * n / t / 1e6:.0f} GB/s")
t = timed(lambda: b.copy_(a))
print(f"copy (1 read : 1 write): {8 * n / t / 1e6:.0f} GB/s")
c = torch.empty(4 * n // 4, dtype=torch.float32, device="cuda")
t = timed(lambda: torch.add(a, b, out=c))
print(f"add (2 reads : 1 write): {12 * n / t / 1e6:.0f} GB/s")
