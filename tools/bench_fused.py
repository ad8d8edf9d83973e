#!/usr/bin/env python
"""Micro-benchmark of the fused NSF layer kernel alone: one layer, N rows, both directions."""
import argparse, os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200 import _fused, _lib
from normalizingflow_b200.flows import NSF_CL

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 20)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--modes", nargs="*", default=["hybrid", "fast", "exact"])
    ap.add_argument("--split", action="store_true", help="split-operand (fp32-class) kernel: conditioner precision fp32x3")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    layer = NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=a.hidden, mask=[1])
    layer.psi.precision = "fp32x3" if a.split else "bf16"
    layer = layer.to(dev)
    x = torch.randn(a.rows, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    ld = torch.zeros(a.rows, device=dev)
    for mode in a.modes:
        layer.arith = mode
        for inv in (False, True):
            for _ in range(3):
                _fused.run(layer, x, inv, ld)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                _fused.run(layer, x, inv, ld)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            print(f"fused mode={mode:6s} inv={int(inv)} {ms:8.3f} ms  {a.rows / ms / 1e3:8.1f} M rows/s  "
                  f"algorithmic {3464 * a.rows / ms / 1e6:8.1f} GB/s", flush=True)

if __name__ == "__main__":
    main()
