#!/usr/bin/env python
"""Micro-benchmark of the RQS coupling kernel alone (SURVEY.md §8(d) transform-kernel input:
params ~ N(0,1) [N,32,23], x ~ N(0,1) [N,64]).  Prints GB/s of algorithmic bytes and the
fraction of the measured HBM peak for each arithmetic mode and direction.

    python tools/bench_kernel.py [--rows 1048576] [--tune R,threads,stages,ctas ...]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200 import _lib, _ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 20)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--tune", nargs="*", default=["0,0,0,0"])
    ap.add_argument("--modes", nargs="*", default=["hybrid", "exact", "fast"])
    ap.add_argument("--size", type=int, default=32)
    ap.add_argument("--dim", type=int, default=2)
    ap.add_argument("--mask", default="1")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    N = a.rows
    mask = [int(c) for c in a.mask.split(",")]
    d = a.size * a.dim
    F_t = a.size * (a.dim - len(mask))
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(N, d, device=dev, generator=g)
    params = torch.randn(N, F_t, 23, device=dev, generator=g)
    peak = 6544.7
    pp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peak = float(json.load(open(pp))["hbm_gbs"])
    row_bytes = F_t * 23 * 4 + 2 * d * 4 + 8
    print(f'size={a.size} dim={a.dim} mask={mask} F_t={F_t} row_bytes={row_bytes}')
    ld = torch.zeros(N, device=dev)
    for tune in a.tune:
        R, th, st, ct = (-1 if v == "g" else int(v) for v in tune.split(","))
        _lib.lib.nfk_set_tuning(R, th, st, ct)
        for mode in a.modes:
            for inv in (False, True):
                for _ in range(3):
                    _ops.rqs_coupling(x, params, a.size, a.dim, mask, 8, 3.0, inv, mode, logdet=ld)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.iters):
                    _ops.rqs_coupling(x, params, a.size, a.dim, mask, 8, 3.0, inv, mode, logdet=ld)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.iters
                gbs = row_bytes * N / (ms * 1e-3) / 1e9
                print(f"tune={tune} mode={mode:6s} inv={int(inv)} {ms:8.3f} ms  {gbs:8.1f} GB/s  {gbs / peak:6.1%} of measured HBM peak"
                      f"  ({N / ms / 1e3:.1f} M rows/s)", flush=True)


if __name__ == "__main__":
    main()
