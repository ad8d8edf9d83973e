#!/usr/bin/env python
"""Micro-benchmark of the wide conditioner GEMM (csrc/gemm_ws.cu) on the H=800 shapes of
BASELINE configs 2 and 3, next to torch.matmul (cuBLAS) on the same bf16 operands."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200 import _wide as W
from normalizingflow_b200._lib import i32_array

dev = torch.device("cuda:0")
peak = 1615.4
pp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = float(json.load(open(pp))["bf16_tflops"])
M = int(os.environ.get("ROWS", 1 << 20))
from normalizingflow_b200 import _lib
_lib.lib.nfk_set_gemm_ws_pair_mode(int(os.environ.get("PAIR", -1)))
print("pair mode", os.environ.get("PAIR", -1), flush=True)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for K, N, act, f32 in ((32, 800, 1, False), (800, 800, 1, False), (800, 736, 0, True), (76, 800, 1, False),
                       (800, 1748, 0, True), (800, 874, 0, True), (128, 128, 1, False)):
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    kb = W.blocks(K)
    a_img = W.pack_input(x, K, 1, [0], kb)
    tiles = W.plan_tiles(W.blocks(N))
    w_img, bp = W.weight_image(w, b, kb, tiles)
    lay = dict(w=w_img, b=bp, KB=kb, kmma_last=(K - 64 * (kb - 1) + 15) // 16, tiles=tiles, tiles_c=i32_array(tiles),
               n_out=N)
    ms = timeit(lambda: W.gemm(a_img, lay, M, act, f32))
    xb, wb = x.to(torch.bfloat16), w.to(torch.bfloat16)
    ms_cublas = timeit(lambda: torch.matmul(xb, wb.t()))
    del x
    fl = 2 * M * K * N
    byt = M * K * 2 + N * K * 2 + M * N * (4 if f32 else 2)
    print("co-resident CTA pairs:", _lib.lib.nfk_gemm_ws_last_clusters(), end="  ")
    print(f"M={M} K={K:4d} N={N:4d} tiles={tiles} {'f32' if f32 else 'bf16'} out: {ms:7.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s "
          f"({fl / ms / 1e9 / peak:5.1%} of measured bf16 peak)  {byt / ms / 1e6:7.1f} GB/s algorithmic | "
          f"cuBLAS bf16 matmul (no bias/act, bf16 out) {ms_cublas:7.3f} ms", flush=True)
    del a_img, xb, wb
    torch.cuda.empty_cache()
