#!/usr/bin/env python
"""Phase timeline (clock64) of one tile of the fused NSF layer kernel: CTA 0, third tile."""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200 import _fused, _lib
from normalizingflow_b200.flows import NSF_CL
dev = torch.device("cuda:0")
torch.manual_seed(0)
layer = NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=128, mask=[1]); layer.psi.precision = "bf16"; layer = layer.to(dev)
layer.arith = sys.argv[1] if len(sys.argv) > 1 else "hybrid"
x = torch.randn(1 << 20, 64, device=dev)
buf = torch.zeros(64, dtype=torch.int64, device=dev)
_fused.run(layer, x, False); torch.cuda.synchronize()
_lib.lib.nfk_set_fused_trace(ctypes.c_void_p(buf.data_ptr()))
_fused.run(layer, x, False); torch.cuda.synchronize()
_lib.lib.nfk_set_fused_trace(ctypes.c_void_p(0))
b = buf.cpu().tolist()
ctl = [v for v in b[:32] if v]; epi = [v for v in b[32:] if v]
t0 = min(ctl + epi)
names_c = ["pre-wait A1", "A1 ready", "MMA1 issued", "pre-wait A2", "A2 ready", "MMA2 issued", "A3 ready"] + [f"MMA3[{c}] issued" for c in range(8)]
names_e = ["pre-wait x", "x landed", "A1 built", "MMA1 done", "epi1 done", "MMA2 done", "epi2 done"]
for c in range(8): names_e += [f"D3[{c}] ready", f"spline[{c}] done"]
ev = [(v - t0, "ctl  " + n) for v, n in zip(ctl, names_c)] + [(v - t0, "epi0 " + n) for v, n in zip(epi, names_e)]
for t, n in sorted(ev): print(f"{t:8d}  {n}")
