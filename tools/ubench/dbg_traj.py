import sys, torch
sys.path.insert(0, ".")
from normalizingflow_b200 import _fused, flows, models
from normalizingflow_b200.hmc import FlowSimulation
torch.manual_seed(0)
dev = torch.device("cuda")
fl = [flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=16, mask=[i % 2]) for i in range(8)]
m = models.NormalizingFlowModel(models.GaussianPrior(64, device=dev), fl, device=dev).to(dev)
for f in fl: f.psi.precision = "bf16"
C = int(sys.argv[1]) if len(sys.argv) > 1 else 5 * 128
gen = torch.Generator(device="cuda").manual_seed(4)
q0 = torch.randn(C, 64, device="cuda", generator=gen) * 0.5
p0 = torch.randn(C, 64, device="cuda", generator=gen)
def run(graph, fold, chain_evals, n=2, path_len=6):
    _fused.CHAIN_EVALS = chain_evals
    sim = FlowSimulation(m, n_chains=C, init_pos=q0)
    sim.use_graph, sim.fused_leapfrog = graph, fold
    sim.set_velocity(p0)
    outs = []
    for _ in range(n):
        q, U = sim.integration_step(path_len=path_len, dt=0.01)
        outs.append((q.clone(), U.clone(), sim.velocity.clone()))
    torch.cuda.synchronize()
    return outs
for graph in (False, True):
    for pl in (1, 6):
        r = run(graph, True, True, path_len=pl)
        rf = run(False, True, False, path_len=pl)
        d = [max(float((a - b).abs().max()) for a, b in zip(x, y)) for x, y in zip(r, rf)]
        print(f"C={C} graph={graph} fold=True chain_evals=True vs fold without chaining, path_len={pl}: max diff per call {d}", flush=True)
