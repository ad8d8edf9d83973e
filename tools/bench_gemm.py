#!/usr/bin/env python
"""Micro-benchmark of nfk_linear_bf16 (tcgen05 GEMM + bias/tanh epilogue) on conditioner shapes."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200 import _bf16
dev = torch.device("cuda:0")
peak = 1615.4
pp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = float(json.load(open(pp))["bf16_tflops"])
M = 1 << 20
for K, N, act, f32 in ((32, 128, 1, False), (128, 128, 1, False), (128, 736, 0, True), (32, 800, 1, False),
                       (800, 800, 1, False), (800, 736, 0, True)):
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    for _ in range(3):
        _bf16.linear_bf16(x, w, b, act, f32)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _bf16.linear_bf16(x, w, b, act, f32)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2 * M * K * N
    byt = M * K * 2 + N * K * 2 + M * N * (4 if f32 else 2)
    print(f"M=2^20 K={K:4d} N={N:4d} {'f32' if f32 else 'bf16'} out: {ms:7.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s ({fl / ms / 1e9 / peak:5.1%} of measured bf16 peak)"
          f"  {byt / ms / 1e6:7.1f} GB/s algorithmic", flush=True)
