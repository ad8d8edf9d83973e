"""Phase timeline (clock64) of CTA 0's second tile of the one-launch layer backward.  Needs the trace build:
   tools/ubench/build_variant.sh bwdtrace nsf_fused_bwd -DFB_TRACE
   NFK_LIB=$PWD/tools/ubench/libnfk_bwdtrace.bin python tools/ubench/trace_bwd.py"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from normalizingflow_b200 import _fused, _lib, flows  # noqa: E402

torch.manual_seed(0)
lay = flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=128, mask=[0]).cuda()
lay.psi.precision = "bf16"
N = 65536
x = torch.randn(N, 64, device="cuda")
g = torch.randn(N, 64, device="cuda")
with torch.no_grad():
    for _ in range(3):
        _fused.layer_backward(lay, x, g, None, 1.0, False)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 256)()
rc = _lib.lib.nfk_fused_bwd_trace_read(buf)
assert rc == 0, rc
b = list(buf)
names = {0: "mma  GEMM1 issued", 1: "mma  GEMM2 issued", 14: "mma  dH1 GEMM issued", 15: "mma  dXc GEMM issued"}
for c in range(8):
    names[2 + c] = f"mma  GEMM3[{c}] issued"
for p in range(4):
    names[10 + p] = f"mma  dH2[{p}] issued"
hid = ["tile start", "A1 built", "GEMM1 done", "h1 written", "GEMM2 done", "h2 written", "dH2 ready", "dZ2 written",
       "dH1 ready", "dZ1 written", "dXc ready", "cond grads stored"]
for i, n in enumerate(hid):
    names[32 + i] = "hid  " + n
for s in range(4):
    for c in range(8):
        names[64 + s * 32 + c * 3] = f"adj{s} chunk {c} wait"
        names[64 + s * 32 + c * 3 + 1] = f"adj{s} chunk {c} D3 ready"
        names[64 + s * 32 + c * 3 + 2] = f"adj{s} chunk {c} G written"
ev = sorted((b[k], n) for k, n in names.items() if b[k])
t0 = ev[0][0]
only0 = "--all" not in sys.argv
for t, n in ev:
    if only0 and n.startswith("adj") and not n.startswith("adj0"):
        continue
    print(f"{t - t0:8d}  {n}")
print("\nslice | per chunk: wait for D3 / adjoint + G (clk)")
for s in range(4):
    cells = []
    for c in range(8):
        w, r, d = (b[64 + s * 32 + c * 3 + i] for i in range(3))
        cells.append(f"{r - w:6d}/{d - r:6d}")
    print(f"{s:5d} | " + " ".join(cells))
