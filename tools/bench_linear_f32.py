#!/usr/bin/env python
"""fp32-accurate conditioner GEMMs of the parity mode: 3xTF32 on tcgen05 (csrc/linear_tf32.cu) against the
CUDA-core fp32 kernel (csrc/linear_f32.cu) and torch's cuBLAS SGEMM, cfg-2 and class-default shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200 import _ops
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
M = 1 << 20
for K, N, act in [(32, 128, 1), (128, 128, 1), (128, 736, 0), (32, 800, 1), (800, 800, 1), (800, 736, 0)]:
    rows = M if K * N < 200000 else M // 4
    x = torch.randn(rows, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    res = {}
    for name, fn in [("tf32x3", lambda: _ops.linear_f32(x, w, b, act)),
                     ("cuda-core", lambda: _ops.linear_f32(x, w, b, act, allow_tc=False)),
                     ("cuBLAS sgemm", lambda: torch.tanh(torch.addmm(b, x, w.t())) if act else torch.addmm(b, x, w.t()))]:
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 5
    fl = 2.0 * rows * K * N
    io = 4.0 * rows * (K + N)
    print(f"[{rows} x {K}] -> {N} act={act}: " + "  ".join(f"{k} {v:7.3f} ms ({fl / v / 1e9:6.1f} TFLOP/s)" for k, v in res.items())
          + f"   HBM floor {io / 6.5447e9:6.3f} ms", flush=True)
