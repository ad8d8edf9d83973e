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
buf = torch.zeros(17 * 32, dtype=torch.int64, device=dev)
_fused.run(layer, x, False); torch.cuda.synchronize()
_lib.lib.nfk_set_fused_trace(ctypes.c_void_p(buf.data_ptr()))
_fused.run(layer, x, False); torch.cuda.synchronize()
_lib.lib.nfk_set_fused_trace(ctypes.c_void_p(0))
b = buf.cpu().tolist()
ctl = [v for v in b[:32] if v]
epi = [[v for v in b[32 + 32 * w:64 + 32 * w] if v] for w in range(16)]
t0 = min(ctl + [v for e in epi for v in e])
names_c = ["pre-wait A1", "A1 ready", "MMA1 issued", "pre-wait A2", "A2 ready", "MMA2 issued", "A3 ready"] + [f"MMA3[{c}] issued" for c in range(8)]
names_e = ["pre-wait x", "x landed", "A1 built", "MMA1 done", "epi1 done", "MMA2 done", "epi2 done"]
for c in range(8): names_e += [f"D3[{c}] ready", f"spline[{c}] done"]
ev = [(v - t0, "ctl  " + n) for v, n in zip(ctl, names_c)] + [(v - t0, "epi0 " + n) for v, n in zip(epi[0], names_e)]
for t, n in sorted(ev): print(f"{t:8d}  {n}")
# per-warp view: time spent waiting for D3[c] and evaluating chunk c (warp w sits on SM sub-partition w % 4;
# the control warp shares sub-partition 0)
print("\nwarp smsp slice | tile start | per chunk: wait / math (clk)")
for w in range(16):
    e = [v - t0 for v in epi[w]]
    if len(e) < 23: continue
    cells = []
    prev = e[6]
    for c in range(8):
        ready, done = e[7 + 2 * c], e[8 + 2 * c]
        cells.append(f"{ready - prev:5d}/{done - ready:5d}")
        prev = done
    print(f"{w:4d} {w % 4:4d} {w // 4:5d} | {e[0]:7d} | " + " ".join(cells) + f" | end {e[22]:7d}")
