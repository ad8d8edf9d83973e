#!/usr/bin/env python
"""One log-prob + grad evaluation of BASELINE config 5 (65536 chains, d=64, 8 x NSF_CL, bf16
conditioner) without CUDA graphs, for an ncu launch list: python tools/profile_hmc.py [H] [evals]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200.flows import NSF_CL
from normalizingflow_b200.hmc import FlowSimulation
from normalizingflow_b200.models import GaussianPrior, NormalizingFlowModel

from normalizingflow_b200 import _lib
_lib.lib.nfk_set_gemm_ws_pair_mode(int(os.environ.get("PAIR", -1)))
H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
evals = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
flows = [NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=H, mask=[i % 2]) for i in range(8)]
for f in flows:
    f.psi.precision = "bf16"
m = NormalizingFlowModel(GaussianPrior(64, device=dev), flows, device=dev).to(dev)
q = torch.randn(65536, 64, device=dev)
sim = FlowSimulation(m, n_chains=65536, init_pos=q)
sim.use_graph = False
for _ in range(evals):
    U, F = sim.potential_and_force(sim.get_position())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
U, F = sim.potential_and_force(sim.get_position())
e1.record()
torch.cuda.synchronize()
print(f"H={H}: one log-prob+grad evaluation of 65536 chains: {e0.elapsed_time(e1):.3f} ms")
