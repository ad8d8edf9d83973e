#!/usr/bin/env python
"""Two forward-KL training steps of the cfg-2 flow (bf16 conditioner) for an ncu launch list:
python tools/profile_train.py [H] [rows]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200.flows import NSF_CL
from normalizingflow_b200.models import GaussianPrior, NormalizingFlowModel

from normalizingflow_b200 import _lib
_lib.lib.nfk_set_gemm_ws_pair_mode(int(os.environ.get("PAIR", -1)))
H = int(sys.argv[1]) if len(sys.argv) > 1 else 800
N = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
dev = torch.device("cuda:0")
torch.manual_seed(0)
fl = [NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=H, mask=[i % 2]) for i in range(8)]
for f in fl:
    f.psi.precision = "bf16"
m = NormalizingFlowModel(GaussianPrior(64, device=dev), fl, device=dev).to(dev)
opt = torch.optim.Adam(m.parameters(), lr=1e-4)
x = torch.randn(N, 64, device=dev)


def step():
    z, plp, ld = m(x)
    loss = -torch.mean(plp + ld)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step()
e1.record()
torch.cuda.synchronize()
print(f"H={H} rows={N}: train step {e0.elapsed_time(e1):.2f} ms")
