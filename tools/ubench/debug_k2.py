import sys, torch
sys.path.insert(0, '.')
from normalizingflow_b200 import _ops as ops
from oracle import nf_oracle as O
size, dim, mask, K, B, N, i = 32, 2, [0], 2, 4.0, 4099, 2
F_t = size * (dim - len(mask))
g = torch.Generator().manual_seed(1000 + i)
x = torch.randn(N, size * dim, generator=g) * (0.6 * B)
x.view(-1)[::97] = B; x.view(-1)[5::131] = -B; x.view(-1)[7::113] = 1.5 * B
params = torch.randn(N, F_t, 3 * K - 1, generator=g) * 1.5
ro, rl, rb = O.nsf_cl_transform(x, params, size, dim, mask, K, B, False)
r64, _, _ = O.nsf_cl_transform(x.double(), params.double(), size, dim, mask, K, B, False)
co, cl, cb = O.nsf_cl_transform(x.cuda(), params.cuda(), size, dim, mask, K, B, False)
for arith in ("hybrid", "fast", "exact"):
    out, ld, bins = ops.rqs_coupling(x.cuda(), params.cuda(), size, dim, mask, K, B, False, arith, want_bins=True)
    e = ((out.cpu().double() - r64).abs() / r64.abs().clamp_min(1)).view(N, size, dim)[:, :, 1]
    ec = ((ro.double() - r64).abs() / r64.abs().clamp_min(1)).view(N, size, dim)[:, :, 1]
    idx = e.flatten().topk(3).indices
    print(arith, "max err vs fp64", float(e.max()), "reference fp32 cpu vs fp64", float(ec.max()))
    for j in idx.tolist():
        n, f = divmod(j, size)
        print("   row", n, "feat", f, "x", float(x.view(N, size, dim)[n, f, 1]), "ours", float(out.view(N, size, dim)[n, f, 1]),
              "ref32", float(ro.view(N, size, dim)[n, f, 1]), "ref64", float(r64.view(N, size, dim)[n, f, 1]), "bin", int(bins[n, f]), int(rb[n, f]),
              "params", [round(float(v), 3) for v in params[n, f]])
