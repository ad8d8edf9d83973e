import sys, torch
sys.path.insert(0, "/root/repo")
from normalizingflow_b200 import _lib
from normalizingflow_b200.flows import NSF_CL
from normalizingflow_b200.models import GaussianPrior, NormalizingFlowModel
from oracle import nf_oracle as O
dev = torch.device("cuda:0")
for pdl in (1, 0):
    _lib.lib.nfk_set_fused2_pdl(pdl)
    for H in (32,):
        torch.manual_seed(0)
        L, N = 4, 1000
        flows = [NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=H, mask=[i % 2]) for i in range(L)]
        model = NormalizingFlowModel(GaussianPrior(64, device=dev), flows, device=dev).to(dev)
        x = torch.randn(N, 64, generator=torch.Generator().manual_seed(1))
        z = torch.randn(N, 64, generator=torch.Generator().manual_seed(2))
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        specs = [dict(type="NSF_CL", size=32, dim=2, K=8, B=3.0, mask=[i % 2]) for i in range(L)]
        rz, rplp, rld = O.flow_forward(specs, sd, x)
        rx, rldi = O.flow_inverse(specs, sd, z)
        def rel(a, b):
            a, b = a.double().cpu(), b.double()
            return float(((a - b).abs() / b.abs().clamp_min(1.0)).max())
        for rep in range(2):
            for prec, arith in (("bf16", "fast"), ("fp32x3", "hybrid"), ("fp32", "hybrid")):
                for f in flows:
                    f.psi.precision, f.arith = prec, arith
                with torch.no_grad():
                    zz, plp, ld = model.forward(x.to(dev))
                    xx, ldi = model.inverse(z.to(dev))
                torch.cuda.synchronize()
                print("pdl", pdl, "rep", rep, prec, arith, "z %.2e ld %.2e x %.2e ldi %.2e" % (rel(zz, rz), rel(ld, rld), rel(xx, rx), rel(ldi, rldi)))
