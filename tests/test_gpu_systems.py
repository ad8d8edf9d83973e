"""GPU parity of the priors / targets next to the flow (applications/src/systems.py)
against fixtures produced by the unmodified reference, plus gradient checks against torch autograd
through the oracle."""
import pytest
import torch

from tests.helpers import T, golden, rel_err

pytestmark = pytest.mark.gpu


def test_einstein_crystal_logprob_sample_and_grad():
    from normalizingflow_b200 import systems
    from oracle import nf_oracle as O
    g = golden("systems.npz")
    for tag in ("ec_free", "ec_box"):
        box = float(g[tag + ".box"]) or None
        alpha = float(g[tag + ".alpha"])
        ec = systems.EinsteinCrystal(T(g[tag + ".centers"]), dim=3, boxlength=box, alpha=alpha)
        x = T(g[tag + ".x"]).cuda()
        assert rel_err(ec.log_prob(x), g[tag + ".lp"]) <= 1e-5
        assert rel_err(ec.potential(x), -g[tag + ".lp"]) <= 1e-5
        xg = x.clone().requires_grad_()
        (gx,) = torch.autograd.grad(ec.log_prob(xg).sum(), xg)
        xr = T(g[tag + ".x"]).clone().requires_grad_()
        (gr,) = torch.autograd.grad(O.einstein_logprob(xr, T(g[tag + ".centers"]), 3, alpha, box).sum(), xr)
        assert rel_err(gx, gr) <= 1e-5
        s = ec.sample(4000)
        assert s.shape == (4000, 36)
        dev = (s.reshape(4000, 12, 3) - ec.centers)
        if box is not None:
            dev = dev - (dev.abs() > 0.5 * box) * torch.sign(dev) * box
        assert abs(float(dev.var()) * alpha - 1.0) < 0.05       # noise variance 1/alpha


@pytest.mark.parametrize("tag", ["lj_nocut", "lj_cut_shift", "lj_cut"])
def test_lj_potential_and_force(tag):
    from normalizingflow_b200 import systems
    from oracle import nf_oracle as O
    g = golden("systems.npz")
    box = float(g["lj.box"])
    cutoff = float(g[tag + ".cutoff"]) or None
    shift = bool(int(g[tag + ".shift"]))
    lj = systems.LJ(boxlength=box, epsilon=1.3, sigma=0.95, cutoff=cutoff, shift=shift)
    pos = T(g["lj.pos"])
    U = lj.potential(pos.cuda())
    assert U.shape == (pos.shape[0],)
    assert rel_err(U, g[tag + ".U"]) <= 2e-5, rel_err(U, g[tag + ".U"])
    # dU/dpos vs autograd through the oracle (pairs sit away from the cutoff radius on this fixture)
    pg = pos.cuda().clone().requires_grad_()
    (gp,) = torch.autograd.grad(lj.potential(pg).sum(), pg)
    pr = pos.clone().requires_grad_()
    (gr,) = torch.autograd.grad(O.lj_potential(pr, box, 1.3, 0.95, cutoff, shift).sum(), pr)
    assert float((gp.cpu() - gr).abs().max()) <= 1e-3 * float(gr.abs().max())


def test_gaussian_mixture_logprob_and_sample():
    from normalizingflow_b200 import systems
    g = golden("systems.npz")
    gm = systems.GaussianMixture(T(g["gm.centers"]), T(g["gm.vars"]), npoints=3, dim=2)
    lp = gm.log_prob(T(g["gm.x"]).cuda())
    assert rel_err(lp, g["gm.lp"]) <= 1e-5
    s = gm.sample(2000, flatten=False)
    assert s.shape == (2000, 3, 2) and torch.isfinite(gm.log_prob(s.reshape(2000, -1))).all()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_bar_on_device_matches_reference(tag):
    """nfk_bar (whole fixed point in one launch, fp64 accumulation) vs applications/src/bar.py."""
    from normalizingflow_b200 import estimators
    g = golden("bar.npz")
    wF, wR = T(g[tag + ".wF"]).cuda(), T(g[tag + ".wR"]).cuda()
    dF, iters = estimators.BAR(wF, wR, return_iterations=True)
    assert abs(dF - float(g[tag + ".dF64"])) <= 1e-9 * max(1.0, abs(dF)) and 2 <= iters <= 1000
    # fp32 work values (what the reference's test.py feeds it): fp64 accumulation agrees with the
    # reference's fp32 numpy run to the iteration tolerance
    dF32 = estimators.BAR(wF.float(), wR.float())
    assert abs(dF32 - float(g[tag + ".dF32"])) <= 2e-4
    assert abs(dF32 - float(g[tag + ".dF64"])) <= 1e-5
    # one iteration from DeltaF = 1 is  1 - BARzero(1)
    one = estimators.BAR(wF, wR, DeltaF=1.0, maximum_iterations=1)
    assert abs(one - (1.0 - float(g[tag + ".fzero_at_1"]))) <= 1e-9


def test_log_mean_exp():
    from normalizingflow_b200 import estimators
    a = torch.randn(777, 33, generator=torch.Generator().manual_seed(3)) * 30
    ref = torch.logsumexp(a.double(), 0) - torch.log(torch.tensor(777.0, dtype=torch.float64))
    assert rel_err(estimators.log_mean_exp(a.cuda(), 0), ref) <= 1e-6
    assert rel_err(estimators.log_mean_exp(a.cuda(), 1), torch.logsumexp(a.double(), 1) - torch.log(torch.tensor(33.0, dtype=torch.float64))) <= 1e-6
