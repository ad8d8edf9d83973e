#!/usr/bin/env python
"""Planar / Radial stacks (BASELINE config 4: 32 layers, d = 128) at 2^20 rows: ms per stack evaluation, samples/s and
the fraction of the HBM roofline (per-layer algorithmic bytes 1,032 B/row/layer; stack-fused floor 1,028 B/row)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200 import _ops
from tools.benchlib import peaks

dev = torch.device("cuda:0")
N, d, L = 1 << 20, 128, 32
torch.manual_seed(0)
x = torch.randn(N, d, device=dev)
w = (torch.rand(L, d, device=dev) * 2 - 1) / d ** 0.5
u = (torch.rand(L, d, device=dev) * 2 - 1) / d ** 0.5
b = (torch.rand(L, device=dev) * 2 - 1) / d ** 0.5
hbm = peaks()[0]


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for mma in (True, False):
    _ops.PLANAR_MMA = mma
    ms = timeit(lambda: _ops.planar_stack(x, w, u, b))
    print(json.dumps({"kernel": "planar_stack_mma (Gram form, tensor cores)" if mma else "planar_stack (per-layer, registers)",
                      "ms": ms, "samples_per_s": N / ms * 1e3,
                      "frac_of_hbm_stack_fused_floor": (2 * d * 4 + 4) * N / (ms * 1e-3) / 1e9 / hbm,
                      "GBps_vs_per_layer_bytes": L * (2 * d * 4 + 8) * N / (ms * 1e-3) / 1e9}))
_ops.PLANAR_MMA = True

# ---- radial: per-layer launches vs the fused runs
from normalizingflow_b200 import flows, models
for per_sample in (True, False):
    layers = [flows.Radial(d, per_sample=per_sample) for _ in range(L)]
    m = models.NormalizingFlowModel(models.GaussianPrior(d, device=dev), layers, device=dev).to(dev)
    for fuse in (True, False):
        m.fuse_planar = fuse
        with torch.no_grad():
            ms = timeit(lambda: m.forward(x), iters=5)
        passes = (1 if (fuse and per_sample) else (L + 0.5 if fuse else (L if per_sample else 1.5 * L)))
        print(json.dumps({"kernel": f"32 x Radial(128, per_sample={per_sample}) " + ("fused run" if fuse else "per-layer launches"),
                          "ms": ms, "ms_per_layer": ms / L, "samples_per_s": N / ms * 1e3,
                          "frac_of_hbm_own_traffic": passes * 2 * d * 4 * N / (ms * 1e-3) / 1e9 / hbm,
                          "GBps_vs_per_layer_bytes": L * (2 * d * 4 + 8) * N / (ms * 1e-3) / 1e9}))
