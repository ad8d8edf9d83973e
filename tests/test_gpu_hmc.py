"""Flow-preconditioned HMC on the GPU (BASELINE config 5): log-prob + grad through the flow's
custom kernels, leapfrog kernels, batched Metropolis."""
import pytest
import torch

from tests.helpers import T, golden, rel_err, sub_sd

pytestmark = pytest.mark.gpu


def _model(H=16, precision="fp32"):
    from normalizingflow_b200 import flows, models
    g = golden("models.npz")
    dev = torch.device("cuda")
    fl = [flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=H, mask=[i % 2]) for i in range(8)]
    m = models.NormalizingFlowModel(models.GaussianPrior(64, device=dev), fl, device=dev)
    m.load_state_dict(sub_sd(g, "nsf.sd."))
    for f in fl:
        f.psi.precision = precision
    return m.to(dev)


def test_force_matches_oracle_autograd():
    """grad_x log p(x) through 8 NSF layers (custom forward + backward kernels) vs torch autograd
    through the oracle's restatement of the reference on the CPU."""
    from normalizingflow_b200.hmc import FlowSimulation
    from oracle import nf_oracle as O
    m = _model()
    g = golden("models.npz")
    x = T(g["nsf.x"])
    sim = FlowSimulation(m, n_chains=x.shape[0], init_pos=x)
    U, F = sim.potential_and_force(sim.get_position())
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    specs = [dict(type="NSF_CL", size=32, dim=2, K=8, B=3.0, mask=[i % 2]) for i in range(8)]
    xr = x.clone().requires_grad_()
    z, plp, ld = O.flow_forward(specs, sd, xr)
    (gref,) = torch.autograd.grad((plp + ld).sum(), xr)
    assert rel_err(-U, (plp + ld).detach()) <= 5e-5
    assert rel_err(F, gref) <= 5e-4, rel_err(F, gref)
    assert all(p.requires_grad for p in m.parameters())      # restored after the dgrad-only pass


def test_leapfrog_is_reversible_and_conserves_energy():
    from normalizingflow_b200.hmc import FlowSimulation
    m = _model()
    C = 256
    gen = torch.Generator(device="cuda").manual_seed(0)
    q0 = torch.randn(C, 64, device="cuda", generator=gen) * 0.5
    p0 = torch.randn(C, 64, device="cuda", generator=gen)
    sim = FlowSimulation(m, n_chains=C, init_pos=q0)
    U0 = sim.get_potential()
    sim.set_velocity(p0)
    q1, U1 = sim.integration_step(path_len=10, dt=0.01)
    H0 = U0 + 0.5 * (p0 * p0).sum(1)
    H1 = U1 + 0.5 * (sim.velocity * sim.velocity).sum(1)
    assert float((H1 - H0).abs().median()) < 0.05            # symplectic: small energy error
    # reverse the momenta and integrate back
    sim.set_velocity(-sim.velocity)
    q2, _ = sim.integration_step(path_len=10, dt=0.01)
    assert float((q2 - q0).abs().max()) < 5e-3
    assert sim.grad_evals == 22


def test_batched_hmc_through_reference_driver_interface():
    from normalizingflow_b200.hmc import HMC, FlowSimulation
    m = _model()
    C = 512
    sim = FlowSimulation(m, n_chains=C, nparticles=32, dim=2, generator=torch.Generator(device="cuda").manual_seed(1))
    h = HMC(sim, path_len=5, dt=0.05, dim=2, beta=1.0)
    pos, pot, logp, acc = h.hmc(epochs=4)
    assert pos.shape == (4, C, 64) and pot.shape == (4, C) and logp.shape == (C,)
    assert 0.2 < acc <= 1.0
    assert torch.isfinite(pos).all() and torch.isfinite(pot).all()
    # the reference's acceptance rule ignores the kinetic energy (Q12), so chains can only keep or
    # lower their potential on average: the batch mean must not increase
    assert float(pot[-1].mean()) <= float(pot[0].mean()) + 1e-3
