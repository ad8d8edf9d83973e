"""Race hunt: first launch after a weight repack vs a repeat of the same launch, bitwise, many trials."""
import sys, torch
sys.path.insert(0, "/root/repo")
from normalizingflow_b200 import _lib
from normalizingflow_b200.flows import NSF_CL
from normalizingflow_b200.models import GaussianPrior, NormalizingFlowModel
dev = torch.device("cuda:0")
torch.manual_seed(0)
L = 4
for pdl in (1, 0):
    _lib.lib.nfk_set_fused2_pdl(pdl)
    for prec, arith in (("fp32x3", "hybrid"), ("bf16", "fast")):
        for N, H in ((1000, 32), (4096, 128), (131072, 128)):
            flows = [NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=H, mask=[i % 2], arith=arith) for i in range(L)]
            for f in flows:
                f.psi.precision = prec
            model = NormalizingFlowModel(GaussianPrior(64, device=dev), flows, device=dev).to(dev)
            hx = torch.randn(N, 64)
            bad = 0
            trials = 60 if N > 10000 else 150
            for t in range(trials):
                with torch.no_grad():
                    for p in model.parameters():
                        p.add_(1e-3 * torch.randn_like(p))          # version bump -> images repacked on next use
                    x = hx.to(dev)
                    a = model.forward(x)
                    torch.cuda.synchronize()
                    b = model.forward(x)
                    torch.cuda.synchronize()
                if not (torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])):
                    bad += 1
            print(f"pdl {pdl} {prec:6s} N {N:6d} H {H:3d}: first-launch != repeat in {bad}/{trials} trials", flush=True)
