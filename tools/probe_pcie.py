#!/usr/bin/env python
"""Pinned host <-> device copy bandwidth of the box (context for bench.py's e2e number)."""
import torch
dev = torch.device("cuda:0")
n = 1 << 26                                    # 268 MB of fp32
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device=dev)
h2 = torch.empty(n, dtype=torch.float32).pin_memory()
d2 = torch.randn(n, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


gb = n * 4 / 1e9
print(f"H2D {gb / t(lambda: d.copy_(h, non_blocking=True)) * 1e3:.1f} GB/s")
print(f"D2H {gb / t(lambda: h2.copy_(d2, non_blocking=True)) * 1e3:.1f} GB/s")


def both():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)


print(f"H2D + D2H concurrently: {2 * gb / t(both) * 1e3:.1f} GB/s total")
