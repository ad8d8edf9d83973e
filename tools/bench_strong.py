#!/usr/bin/env python
"""One GPU's share of the strong-scaling step (2^20 rows over 8 GPUs = 131,072 rows): fwd + inv of the cfg-2 flow,
eager and as one CUDA graph, with and without programmatic dependent launch.  python tools/bench_strong.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200 import _lib
from normalizingflow_b200.flows import NSF_CL
from normalizingflow_b200.graphs import GraphedCallable
from normalizingflow_b200.models import GaussianPrior, NormalizingFlowModel

dev = torch.device("cuda:0")
torch.manual_seed(0)
fl = [NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=128, mask=[i % 2], arith="fast") for i in range(8)]
for f in fl:
    f.psi.precision = "bf16"
m = NormalizingFlowModel(GaussianPrior(64, device=dev), fl, device=dev).to(dev)
for rows in (131072, 262144, 1 << 20):
    x = torch.randn(rows, 64, device=dev)
    z = torch.randn(rows, 64, device=dev)

    def step(xx, zz):
        with torch.no_grad():
            a = m.forward(xx)
            b = m.inverse(zz)
        return a + b

    def timeit(fn, iters=20):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    for pdl in (1, 0):
        _lib.lib.nfk_set_fused2_pdl(pdl)
        t_e = timeit(lambda: step(x, z))
        g = GraphedCallable(step, x, z)
        t_g = timeit(lambda: g.graph.replay())
        print(f"rows {rows:8d} pdl {pdl}: eager {t_e:.4f} ms  graph {t_g:.4f} ms  -> {rows / t_g / 1e3:.1f} M samples/s per GPU (graph)", flush=True)
_lib.lib.nfk_set_fused2_pdl(1)
