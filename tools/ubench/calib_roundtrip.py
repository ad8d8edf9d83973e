import torch, sys
sys.path.insert(0, '.')
from normalizingflow_b200 import _ops as ops
from normalizingflow_b200.flows import NSF_CL
dev = torch.device("cuda"); N = 1 << 20
def q(t, p):
    t = t.abs().flatten().float(); return float(t.kthvalue(max(1, int(p * t.numel()))).values)
g = torch.Generator(device=dev).manual_seed(11)
x = torch.randn(N, 64, device=dev, generator=g)
params = torch.randn(N, 32, 23, device=dev, generator=g)
for arith in ("hybrid", "fast", "exact"):
    z, ld, bf = ops.rqs_coupling(x, params, 32, 2, [0], 8, 3.0, False, arith, want_bins=True)
    x2, ld2, bi = ops.rqs_coupling(z, params, 32, 2, [0], 8, 3.0, True, arith, want_bins=True)
    err = (x2 - x).abs() / x.abs().clamp_min(1.0)
    s = (ld + ld2).abs() / ld.abs().clamp_min(1.0)
    print(arith, "x:", [f"{q(err, p):.2e}" for p in (0.5, 0.99, 0.9999)], f"max {float(err.max()):.2e}",
          "ld:", [f"{q(s, p):.2e}" for p in (0.5, 0.99, 0.9999)], f"max {float(s.max()):.2e}",
          "bin mism", int(((bf != bi) & (bf >= 0)).sum()))
del params
for hidden in (128, 800):
    torch.manual_seed(5)
    layer = NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=hidden, mask=[0]); layer.psi.precision = "bf16"; layer = layer.to(dev)
    with torch.no_grad():
        z, ld = layer.forward(x); x2, ld2 = layer.inverse(z)
    err = (x2 - x).abs() / x.abs().clamp_min(1.0)
    s = (ld + ld2).abs() / ld.abs().clamp_min(1.0)
    print(hidden, "x:", [f"{q(err, p):.2e}" for p in (0.5, 0.99, 0.9999)], f"max {float(err.max()):.2e}",
          "ld:", [f"{q(s, p):.2e}" for p in (0.5, 0.99, 0.9999)], f"max {float(s.max()):.2e}")
