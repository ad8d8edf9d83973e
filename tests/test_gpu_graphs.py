"""CUDA-graph capture of whole steps (normalizingflow_b200/graphs.py): replaying a captured sampling
call / training step gives what the eager calls give."""
import copy

import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _model(precision):
    from normalizingflow_b200 import flows, models
    dev = torch.device("cuda")
    torch.manual_seed(3)
    fl = [flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=48, mask=[i % 2]) for i in range(3)] + [flows.RealNVP(64, hidden_dim=24)]
    for f in fl[:3]:
        f.psi.precision = precision
    return models.NormalizingFlowModel(models.GaussianPrior(64, device=dev), fl, device=dev).to(dev)


def test_graphed_sampling_call_equals_eager():
    from normalizingflow_b200.graphs import GraphedCallable
    m = _model("bf16")
    z1 = torch.randn(512, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    z2 = torch.randn(512, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    g = GraphedCallable(lambda z: m.inverse(z), z1)
    with torch.no_grad():
        for z in (z1, z2, z1):
            xg, ldg = g(z)
            xe, lde = m.inverse(z)
            assert torch.equal(xg, xe) and torch.equal(ldg, lde)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_train_step_tracks_eager_training(precision):
    """Three forward-KL steps (applications/src/train.py:22-29) replayed from one captured graph vs the
    same steps run eagerly from the same initial weights."""
    from normalizingflow_b200.graphs import GraphedTrainStep
    me = _model(precision)
    mg = copy.deepcopy(me)
    x = torch.randn(640, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))

    def loss_of(m):
        def f(xx):
            z, plp, ld = m(xx)
            return -torch.mean(plp + ld)
        return f
    oe = torch.optim.Adam(me.parameters(), lr=1e-3, capturable=True)
    og = torch.optim.Adam(mg.parameters(), lr=1e-3, capturable=True)
    # GraphedTrainStep runs `warmup` real optimisation steps while warming up: give the eager twin the same
    step = GraphedTrainStep(loss_of(mg), og, x, warmup=2)
    losses_e = []
    for _ in range(2 + 3):
        oe.zero_grad(set_to_none=True)
        l = loss_of(me)(x)
        l.backward()
        oe.step()
        losses_e.append(float(l.detach()))
    losses_g = [float(step(x)) for _ in range(3)]
    tol = 2e-3 if precision == "bf16" else 2e-4
    for a, b in zip(losses_g, losses_e[2:]):
        assert abs(a - b) <= tol * max(1.0, abs(b)), (losses_g, losses_e)
    assert losses_g[-1] < losses_g[0]
    for pe, pg in zip(me.parameters(), mg.parameters()):
        assert rel_err(pg, pe.detach().double().cpu()) <= 5e-3


def test_eager_evaluation_after_graphed_steps_sees_the_new_weights():
    """A replayed GraphedTrainStep rewrites the parameters without bumping their version counters; the
    weight-image caches (fused / wide) must still be refreshed (parameter epoch, _lib.invalidate_caches):
    eager evaluate() after graphed steps equals a fresh deep copy of the trained model."""
    from normalizingflow_b200.graphs import GraphedTrainStep
    m = _model("bf16")
    x = torch.randn(640, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(6))
    with torch.no_grad():
        before = m.evaluate(x).clone()          # fills the inference-side weight-image caches

    def loss(xx):
        z, plp, ld = m(xx)
        return -torch.mean(plp + ld)
    opt = torch.optim.Adam(m.parameters(), lr=5e-3, capturable=True)
    step = GraphedTrainStep(loss, opt, x, warmup=1)
    for _ in range(3):
        step(x)
    with torch.no_grad():
        after = m.evaluate(x)
        fresh = copy.deepcopy(m)
        for mod in fresh.modules():             # a copy with no caches at all
            for k in [k for k in mod.__dict__ if k.endswith("_cache") or k == "_ar_tables"]:
                mod.__dict__.pop(k)
        ref = fresh.evaluate(x)
    assert torch.equal(after, ref)
    assert float((after - before).abs().max()) > 1e-3
