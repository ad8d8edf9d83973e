"""Where an HMC epoch (65,536 chains, path_len 10) spends GPU time outside the trajectory graph."""
import sys
import torch
sys.path.insert(0, ".")
from normalizingflow_b200 import flows, models
from normalizingflow_b200.hmc import HMC, FlowSimulation
torch.manual_seed(0)
dev = torch.device("cuda")
fl = [flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=128, mask=[i % 2]) for i in range(8)]
m = models.NormalizingFlowModel(models.GaussianPrior(64, device=dev), fl, device=dev).to(dev)
for f in fl:
    f.psi.precision = "bf16"
C = 65536
sim = FlowSimulation(m, n_chains=C, nparticles=32, dim=2, generator=torch.Generator(device=dev).manual_seed(5))
h = HMC(sim, path_len=10, dt=0.05, dim=2, beta=1.0)
h.hmc(epochs=3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
h.hmc(epochs=8)
e1.record()
torch.cuda.synchronize()
print(f"epoch: {e0.elapsed_time(e1) / 8:.3f} ms")
sim.set_velocity(torch.randn(C, 64, device=dev))
e0.record()
for _ in range(8):
    sim.integration_step(10, 0.05)
e1.record()
torch.cuda.synchronize()
print(f"integration_step alone: {e0.elapsed_time(e1) / 8:.3f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    h.hmc(epochs=4)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
