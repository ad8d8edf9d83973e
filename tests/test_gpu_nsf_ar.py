"""GPU parity of the autoregressive spline flow NSF_AR (nf/flows.py:152-209) and of
a reference-written checkpoint (N3) against fixtures produced by the unmodified reference."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import RTOL_FP32, T, golden, rel_err, sub_sd

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _layer(g, tag, arith="hybrid"):
    from normalizingflow_b200 import flows
    dim, K, B, H = int(g[tag + ".dim"]), int(g[tag + ".K"]), float(g[tag + ".B"]), int(g[tag + ".H"])
    lay = flows.NSF_AR(dim, K=K, B=B, hidden_dim=H, arith=arith)
    res = lay.load_state_dict(sub_sd(g, tag + ".sd."))
    assert not res.missing_keys and not res.unexpected_keys
    return lay.cuda(), dim, K, B


@pytest.mark.parametrize("tag", ["d6k8", "d4k32", "d12k8"])
def test_nsf_ar_forward_inverse_match_reference(tag):
    g = golden("nsf_ar.npz")
    lay, dim, K, B = _layer(g, tag)
    with torch.no_grad():
        z, ld = lay.forward(T(g[tag + ".x"]).cuda())
        x, ldi = lay.inverse(T(g[tag + ".zin"]).cuda())
    # fp32 conditioner + spline per dimension; log_det sums `dim` per-element terms
    assert rel_err(z, g[tag + ".z"]) <= 2 * RTOL_FP32, rel_err(z, g[tag + ".z"])
    assert rel_err(ld, g[tag + ".ld"]) <= 5 * RTOL_FP32, rel_err(ld, g[tag + ".ld"])
    # the inverse feeds each output into the next conditioner: errors compound over `dim` steps
    assert rel_err(x, g[tag + ".x_inv"]) <= 1e-4, rel_err(x, g[tag + ".x_inv"])
    assert rel_err(ldi, g[tag + ".ld_inv"]) <= 2e-4, rel_err(ldi, g[tag + ".ld_inv"])


@pytest.mark.parametrize("tag", ["d6k8", "d4k32"])
def test_nsf_ar_bins_match_reference(tag):
    """bin indices of the forward pass, given the reference-identical fp32 conditioner chain"""
    from normalizingflow_b200 import _ops
    g = golden("nsf_ar.npz")
    lay, dim, K, B = _layer(g, tag)
    x = T(g[tag + ".x"]).cuda()
    with torch.no_grad():
        N, P = x.shape[0], 3 * K - 1
        ang = torch.tensor(np.pi, dtype=torch.float32, device="cuda") * x / B
        c, s = torch.cos(ang), torch.sin(ang)
        cols = [lay.init_param.expand(N, P)] + [lay.layers[i - 1](torch.cat((c[:, :i], s[:, :i]), -1)) for i in range(1, dim)]
        _, _, bins = _ops.rqs_elementwise(x, torch.stack(cols, 1), K, B, False, "hybrid", want_bins=True)
    ref = g[tag + ".bins"]
    mism = int((bins.cpu().numpy().astype(np.int64) != ref).sum())
    # the conditioner runs on another device than the reference's (fp32 GEMM summation order), so an
    # input within an ulp of a knot may move: allow a handful, none expected on these fixtures
    assert mism <= 2, mism


@pytest.mark.parametrize("tag", ["d6k8", "d12k8"])
def test_nsf_ar_gradients_match_reference_autograd(tag):
    g = golden("nsf_ar.npz")
    lay, dim, K, B = _layer(g, tag)
    x = T(g[tag + ".x"]).cuda().requires_grad_()
    r, s = T(g[tag + ".r"]).cuda(), T(g[tag + ".s"]).cuda()
    z, ld = lay.forward(x)
    loss = (z * r).sum() + (ld * s).sum()
    names = [n for n, _ in lay.named_parameters()]
    grads = torch.autograd.grad(loss, [x] + list(lay.parameters()))
    gx_ref = T(g[tag + ".gx"])
    assert float((grads[0].cpu() - gx_ref).abs().max()) <= 2e-3 * float(gx_ref.abs().max())
    for n, gv in zip(names, grads[1:]):
        ref = T(g[tag + ".g." + n])
        assert float((gv.cpu() - ref).abs().max()) <= 2e-3 * max(1e-3, float(ref.abs().max())), n


def test_reference_checkpoint_reproduces_reference_outputs():
    from normalizingflow_b200 import checkpoint, flows, models
    dev = torch.device("cuda")
    fl = [flows.NSF_CL(4, dim=2, K=8, B=3.0, hidden_dim=8, mask=[i % 2]) for i in range(2)] + [flows.RealNVP(8, hidden_dim=8)]
    m = models.NormalizingFlowModel(models.GaussianPrior(8, device=dev), fl, device=dev)
    _, _, res = checkpoint.load_checkpoint(os.path.join(ROOT, "tests", "golden", "ref_checkpoint.pth"), m, device=dev)
    assert not res.missing_keys and not res.unexpected_keys
    g = golden("ref_checkpoint_out.npz")
    with torch.no_grad():
        z, plp, ld = m.forward(T(g["x"]).cuda())
    assert rel_err(z, g["z"]) <= 3 * RTOL_FP32
    assert rel_err(plp, g["plp"]) <= 5e-5 and rel_err(ld, g["ld"]) <= 5e-5


@pytest.mark.parametrize("dim,K,H", [(12, 8, 24), (40, 32, 200), (96, 8, 130)])
def test_nsf_ar_grouped_conditioners_match_per_layer_path(dim, K, H):
    """bf16 conditioners: the grouped path (one trig-feature pack launch + three grouped tcgen05 GEMM
    launches for all dim-1 conditioners) vs running the conditioners one by one, and vs the fp32
    layer (1e-2 class)."""
    from normalizingflow_b200 import _wide, flows
    torch.manual_seed(dim)
    lay = flows.NSF_AR(dim, K=K, B=3.0, hidden_dim=H).cuda()
    x = (1.3 * torch.randn(700, dim, generator=torch.Generator().manual_seed(2))).cuda()
    with torch.no_grad():
        z32, ld32 = lay.forward(x)                          # fp32 conditioners
        for f in lay.layers:
            f.precision = "bf16"
        assert _wide.nsf_ar_grouped_ok(lay)
        zg, ldg = lay.forward(x)                            # grouped
        pg = _wide.nsf_ar_params(lay, x, max_rows=256)      # row-chunked: same numbers
        P = 3 * K - 1
        ang = torch.tensor(np.pi, dtype=torch.float32, device="cuda") * x / 3.0
        c, s = torch.cos(ang), torch.sin(ang)
        cols = [lay.init_param.expand(700, P)] + [lay.layers[i - 1](torch.cat((c[:, :i], s[:, :i]), -1)) for i in range(1, dim)]
        pl = torch.stack(cols, 1)                           # one conditioner at a time (bf16)
    assert rel_err(pg, pl.double().cpu()) <= 2e-2, rel_err(pg, pl.double().cpu())
    assert rel_err(_wide.nsf_ar_params(lay, x), pg.double().cpu()) <= 1e-6
    assert rel_err(zg, z32.double().cpu()) <= 2e-2 and rel_err(ldg, ld32.double().cpu()) <= 5e-2
