"""Time per log-prob + gradient evaluation inside a leapfrog trajectory (65,536 chains, path_len 10), graph replay and
eager, with and without the leapfrog fold / tile-flag chains."""
import sys
import torch
sys.path.insert(0, ".")
from normalizingflow_b200 import _fused, flows, models
from normalizingflow_b200.hmc import FlowSimulation
torch.manual_seed(0)
dev = torch.device("cuda")
fl = [flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=128, mask=[i % 2]) for i in range(8)]
m = models.NormalizingFlowModel(models.GaussianPrior(64, device=dev), fl, device=dev).to(dev)
for f in fl:
    f.psi.precision = "bf16"
C = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
q0 = 0.7 * torch.randn(C, 64, device=dev)
p0 = torch.randn(C, 64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for graph in (True,):
    for chain, fold, ce in ((True, True, True), (True, True, False), (True, False, False), (False, False, False)):
        if True:
            _fused.TILE_CHAIN = chain
            _fused.CHAIN_EVALS = ce
            sim = FlowSimulation(m, n_chains=C, init_pos=q0)
            sim.use_graph, sim.fused_leapfrog = graph, fold
            sim.set_velocity(p0)
            for _ in range(2):
                sim.integration_step(path_len=10, dt=0.01)
            torch.cuda.synchronize()
            n0 = sim.grad_evals
            e0.record()
            for _ in range(5):
                sim.integration_step(path_len=10, dt=0.01)
            e1.record()
            torch.cuda.synchronize()
            print(f"C={C} graph={graph} tile_chain={chain} leapfrog_fold={fold} chain_evals={ce}: {e0.elapsed_time(e1) / (sim.grad_evals - n0):.4f} ms per evaluation "
                  f"({sim.grad_evals - n0} evaluations)", flush=True)
_fused.TILE_CHAIN = True
