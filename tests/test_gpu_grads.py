"""GPU parity of the hand-written backward kernels against autograd gradients of the unmodified
reference (tests/golden/grads.npz): d(sum(out*gz) + sum(ld*gl)) w.r.t. inputs, conditioner
outputs and every weight."""
import numpy as np
import pytest
import torch

from tests.helpers import T, golden, rel_err, sub_sd

pytestmark = pytest.mark.gpu
GTOL = 2e-4      # |a-b| / max(1,|b|): gradients chain ~10 fp32 ops with cancellations


def test_rqs_bwd_kernel_and_layer_autograd():
    from normalizingflow_b200 import _ops, flows
    g = golden("grads.npz")
    for mi in (0, 1):
        size, dim = int(g[f"nsf{mi}.size"]), int(g[f"nsf{mi}.dim"])
        mask = [int(m) for m in g[f"nsf{mi}.mask"]]
        layer = flows.NSF_CL(size, dim=dim, K=8, B=3.0, hidden_dim=12, mask=mask)
        layer.load_state_dict(sub_sd(g, f"nsf{mi}.sd."))
        layer = layer.cuda()
        for inv in (False, True):
            s = f"nsf{mi}." + ("inv." if inv else "fwd.")
            x, gz, gl = T(g[s + "x"]).cuda(), T(g[s + "gz"]).cuda(), T(g[s + "gl"]).cuda()
            params = T(g[s + "params"]).cuda()
            # (1) the kernel alone, on the reference's own conditioner output
            gx_direct, gp = _ops.rqs_coupling_bwd(x, params, gz, gl, size, dim, mask, 8, 3.0, inv)
            assert rel_err(gp, g[s + "gparams"]) <= GTOL, (mi, inv, rel_err(gp, g[s + "gparams"]))
            # (2) the whole layer through torch.autograd (conditioner fp32 kernels + their backward)
            xr = x.clone().requires_grad_()
            layer.zero_grad()
            out, ld = layer.inverse(xr) if inv else layer.forward(xr)
            ((out * gz).sum() + (ld * gl).sum()).backward()
            assert rel_err(out.detach(), g[s + "out"]) <= 2e-5 and rel_err(ld.detach(), g[s + "ld"]) <= 5e-5
            assert rel_err(xr.grad, g[s + "gx"]) <= GTOL, (mi, inv, rel_err(xr.grad, g[s + "gx"]))
            for k, v in layer.named_parameters():
                assert rel_err(v.grad, g[s + "gw." + k]) <= GTOL, (mi, inv, k, rel_err(v.grad, g[s + "gw." + k]))


def test_rqs_bwd_matches_finite_differences_of_forward_kernel():
    """independent of the reference: central differences of the EXACT forward kernel in x"""
    from normalizingflow_b200 import _ops
    gen = torch.Generator().manual_seed(9)
    N = 64
    x = (torch.rand(N, 64, generator=gen) * 4 - 2).cuda()
    params = torch.randn(N, 32, 23, generator=gen).cuda()
    for inv in (False, True):
        gz = torch.ones(N, 64, device="cuda")
        gx, _ = _ops.rqs_coupling_bwd(x, params, gz, None, 32, 2, [1], 8, 3.0, inv)
        h = 1e-3
        op, _, bp = _ops.rqs_coupling(x + h, params, 32, 2, [1], 8, 3.0, inv, "exact", want_bins=True)
        om, _, bm = _ops.rqs_coupling(x - h, params, 32, 2, [1], 8, 3.0, inv, "exact", want_bins=True)
        fd = ((op - om) / (2 * h)).reshape(N, 32, 2)[:, :, 1]          # d y / d x of the transformed column
        an = gx.reshape(N, 32, 2)[:, :, 0]                              # mask [1]: transformed input is column 0
        same_bin = (bp == bm)
        assert same_bin.float().mean() > 0.9
        err = ((fd - an).abs() / an.abs().clamp_min(1.0))[same_bin]
        assert float(err.max()) <= 5e-3, float(err.max())


def test_realnvp_grads():
    from normalizingflow_b200 import flows
    g = golden("grads.npz")
    layer = flows.RealNVP(6, hidden_dim=10)
    layer.load_state_dict(sub_sd(g, "rnvp.sd."))
    layer = layer.cuda()
    for inv in (False, True):
        s = "rnvp." + ("inv." if inv else "fwd.")
        x = T(g[s + "x"]).cuda().requires_grad_()
        layer.zero_grad()
        out, ld = layer.inverse(x) if inv else layer.forward(x)
        ((out * T(g[s + "gz"]).cuda()).sum() + (ld * T(g[s + "gl"]).cuda()).sum()).backward()
        assert rel_err(x.grad, g[s + "gx"]) <= GTOL
        for k, v in layer.named_parameters():
            assert rel_err(v.grad, g[s + "gw." + k]) <= GTOL, (inv, k)


def test_planar_grads_single_and_fused_stack():
    from normalizingflow_b200 import flows, models
    g = golden("grads.npz")
    pl = flows.Planar(16)
    pl.load_state_dict(sub_sd(g, "planar.sd."))
    pl = pl.cuda()
    x = T(g["planar.x"]).cuda().requires_grad_()
    out, ld = pl.forward(x)
    ((out * T(g["planar.gz"]).cuda()).sum() + (ld * T(g["planar.gl"]).cuda()).sum()).backward()
    assert rel_err(out.detach(), g["planar.out"]) <= 1e-5 and rel_err(ld.detach(), g["planar.ld"]) <= 1e-5
    assert rel_err(x.grad, g["planar.gx"]) <= GTOL
    for k, v in pl.named_parameters():
        assert rel_err(v.grad, g["planar.gw." + k]) <= GTOL, k
    # fused 5-layer stack == the same 5 layers applied one by one (autograd through both)
    torch.manual_seed(4)
    layers = [flows.Planar(16).cuda() for _ in range(5)]
    xs = torch.randn(300, 16, device="cuda")
    res = []
    for fuse in (True, False):
        m = models.NormalizingFlowModel(models.GaussianPrior(16, device="cuda"), layers, device="cuda", fuse_planar=fuse)
        m.zero_grad()
        xr = xs.clone().requires_grad_()
        z, plp, ldt = m.forward(xr)
        (-(plp + ldt).mean()).backward()
        res.append((z.detach(), ldt.detach(), xr.grad, [p.grad.clone() for p in m.parameters()]))
    assert rel_err(res[0][0], res[1][0]) <= 1e-5 and rel_err(res[0][1], res[1][1]) <= 1e-5
    assert rel_err(res[0][2], res[1][2]) <= GTOL
    for a, b in zip(res[0][3], res[1][3]):
        assert rel_err(a, b) <= GTOL


def test_radial_grads_global_and_per_sample():
    from normalizingflow_b200 import flows
    from oracle import nf_oracle as O
    g = golden("grads.npz")
    rd = flows.Radial(16)
    rd.load_state_dict(sub_sd(g, "radial.sd."))
    rd = rd.cuda()
    x = T(g["radial.x"]).cuda().requires_grad_()
    out, ld = rd.forward(x)
    ((out * T(g["radial.gz"]).cuda()).sum() + (ld * T(g["radial.gl"]).cuda()).sum()).backward()
    assert ld.shape == (1,)
    assert rel_err(out.detach(), g["radial.out"]) <= 1e-5 and rel_err(ld.detach(), g["radial.ld"]) <= 1e-5
    assert rel_err(x.grad, g["radial.gx"]) <= GTOL
    for k, v in rd.named_parameters():
        assert rel_err(v.grad, g["radial.gw." + k]) <= GTOL, (k, rel_err(v.grad, g["radial.gw." + k]))
    # corrected per-row-norm mode vs autograd through the oracle
    rd2 = flows.Radial(16, per_sample=True)
    rd2.load_state_dict(sub_sd(g, "radial.sd."))
    rd2 = rd2.cuda()
    gz, gl = T(g["radial.gz"]), T(g["radial.gl"])
    xs = T(g["radial.x"])
    x2 = xs.cuda().requires_grad_()
    out2, ld2 = rd2.forward(x2)
    ((out2 * gz.cuda()).sum() + (ld2 * gl.cuda()).sum()).backward()
    ps = {k: T(g["radial.sd." + k]).clone().requires_grad_() for k in ("x0", "log_alpha", "beta")}
    xr = xs.clone().requires_grad_()
    ro, rl = O.radial(xr, ps["x0"], ps["log_alpha"], ps["beta"], per_sample=True)
    ((ro * gz).sum() + (rl * gl).sum()).backward()
    assert ld2.shape == (20,)
    assert rel_err(out2.detach(), ro.detach()) <= 1e-5 and rel_err(ld2.detach(), rl.detach()) <= 1e-5
    assert rel_err(x2.grad, xr.grad) <= GTOL
    for k, v in rd2.named_parameters():
        assert rel_err(v.grad, ps[k].grad) <= GTOL, k
