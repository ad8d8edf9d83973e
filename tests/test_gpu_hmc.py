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


def test_tensor_core_force_matches_autograd_and_oracle():
    """bf16-conditioner model: the hand-written forward+backward path (spline adjoint as a GEMM
    epilogue + dgrad GEMMs, csrc/gemm_ws.cu) vs autograd through the same model, and vs torch
    autograd through the fp32 oracle (1e-2 class of the bf16 conditioner)."""
    from normalizingflow_b200 import _wide
    from normalizingflow_b200.hmc import FlowSimulation
    from oracle import nf_oracle as O
    m = _model(precision="bf16")
    g = golden("models.npz")
    x = T(g["nsf.x"])
    assert _wide.flow_logp_and_grad(m, x.cuda()) is not None
    sim = FlowSimulation(m, n_chains=x.shape[0], init_pos=x)
    sim.fused_grad = False          # this test is about the wide (multi-launch, bf16-operand) path; the one-launch path at
    #                                 hidden width <= 128 has fp16 forward operands and its own test below
    U, F = sim.potential_and_force(sim.get_position())
    sim.tensor_core_grad = False
    for f in m.flows:
        f.fused = False                                   # reference: unfused autograd Functions
    U2, F2 = sim.potential_and_force(sim.get_position())
    scale = float(F2.abs().max())
    assert rel_err(U, U2.double().cpu()) <= 2e-3
    assert float((F - F2).abs().max()) <= 2e-2 * scale
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    specs = [dict(type="NSF_CL", size=32, dim=2, K=8, B=3.0, mask=[i % 2]) for i in range(8)]
    xr = x.clone().requires_grad_()
    z, plp, ld = O.flow_forward(specs, sd, xr)
    (gref,) = torch.autograd.grad((plp + ld).sum(), xr)
    assert rel_err(-U, (plp + ld).detach()) <= 1e-2
    # d logdet / dx is discontinuous across knots (the spline is C1): the bf16 conditioner moves the
    # knots by ~1e-3, so a few elements next to a knot see another bin's second derivative -- gate
    # the bulk, not the maximum
    err = (F.cpu() - gref).abs() / (1.0 + gref.abs())
    assert float(err.median()) <= 5e-3, float(err.median())
    assert float((err > 3e-2).float().mean()) <= 8e-2, float((err > 3e-2).float().mean())


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("size,dim,mask,H", [(32, 2, [1], 200), (38, 3, [0, 2], 160), (38, 3, [1], 96)])
def test_layer_backward_matches_autograd(inverse, size, dim, mask, H):
    """One layer: dL/dx from the tensor-core backward vs autograd through the unfused kernels of
    the SAME bf16-conditioner layer (same hidden activations, fp32 dgrad), for
    L = sum(out * r) + sum(logdet * s) with random r, s."""
    from normalizingflow_b200 import _wide, flows
    torch.manual_seed(7)
    lay = flows.NSF_CL(size, dim=dim, K=8, B=3.0, hidden_dim=H, mask=mask).cuda()
    with torch.no_grad():
        lay.psi.network[4].weight.mul_(3.0)
    lay.psi.precision = "bf16"
    N, d = 900, size * dim
    gen = torch.Generator().manual_seed(11)
    x = (1.2 * torch.randn(N, d, generator=gen)).cuda()
    r = torch.randn(N, d, generator=gen).cuda()
    s = torch.randn(N, generator=gen).cuda()
    xg = x.clone().requires_grad_()
    lay.fused = False
    out, ld = (lay.inverse if inverse else lay.forward)(xg)          # autograd Functions (unfused)
    names = [n for n, _ in lay.named_parameters()]
    gall = torch.autograd.grad((out * r).sum() + (ld * s).sum(), [xg] + list(lay.parameters()))
    gref = gall[0]
    lay.fused = True
    assert _wide.grad_eligible(lay)
    # training path: the same loss through NsfWideFn (tensor-core dgrad, batch-contraction wgrad)
    xg2 = x.clone().requires_grad_()
    out_t, ld_t = (lay.inverse if inverse else lay.forward)(xg2)
    gtrain = torch.autograd.grad((out_t * r).sum() + (ld_t * s).sum(), [xg2] + list(lay.parameters()))
    for n, ga, gb in zip(["x"] + names, gtrain, gall):
        scale = max(1e-6, float(gb.abs().max()))
        assert float((ga - gb).abs().max()) <= 4e-2 * scale, (n, float((ga - gb).abs().max()), scale)
    with torch.no_grad():
        o2, l2, ctx = _wide.layer_forward_saving(lay, x, inverse)
        gin = _wide.layer_backward(lay, ctx, r, s)
    assert rel_err(o2.cpu(), out.detach().double().cpu()) <= 2e-3
    assert rel_err(l2.cpu(), ld.detach().double().cpu()) <= 5e-3
    err = (gin - gref).abs() / (1.0 + gref.abs())
    assert float(err.median()) <= 2e-3, float(err.median())
    assert float((err > 3e-2).float().mean()) <= 5e-3, float((err > 3e-2).float().mean())


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("mask,H,N", [(0, 128, 1024), (1, 128, 1024), (1, 64, 900), (0, 16, 130)])
def test_fused_layer_backward_matches_fp32_autograd(inverse, mask, H, N):
    """ONE launch per layer (csrc/nsf_fused_bwd.cu: conditioner recomputed with the forward kernel's fp16 operands,
    spline adjoint in registers, three dgrad GEMMs on chip) vs autograd through the fp32 parity path of the same
    layer, for L = sum(out * r) + sum(logdet * s).  d logdet / dx jumps across knots (the spline is C1), so the
    bulk is gated, not the maximum; a partial last tile is padded."""
    from normalizingflow_b200 import _fused, flows
    torch.manual_seed(7)
    lay = flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=H, mask=[mask]).cuda()
    with torch.no_grad():
        lay.psi.network[4].weight.mul_(3.0)
    gen = torch.Generator().manual_seed(11)
    x = (1.2 * torch.randn(N, 64, generator=gen)).cuda()
    r = torch.randn(N, 64, generator=gen).cuda()
    s = torch.randn(N, generator=gen).cuda()
    lay.fused = False
    lay.psi.precision = "fp32"
    xg = x.clone().requires_grad_()
    out, ld = (lay.inverse if inverse else lay.forward)(xg)
    (gref,) = torch.autograd.grad((out * r).sum() + (ld * s).sum(), [xg])
    lay.psi.precision = "bf16"
    lay.fused = True
    assert _fused.bwd_eligible(lay)
    with torch.no_grad():
        gin = _fused.layer_backward(lay, x, r, s, 1.0, inverse)
        g_const = _fused.layer_backward(lay, x, r, None, 1.0, inverse)
        g_ones = _fused.layer_backward(lay, x, r, torch.ones_like(s), 0.0, inverse)
    assert gin.shape == (N, 64) and torch.isfinite(gin).all()
    err = (gin - gref).abs() / (1.0 + gref.abs())
    assert float(err.median()) <= 2e-3, float(err.median())
    assert float((err > 3e-2).float().mean()) <= 1e-2, float((err > 3e-2).float().mean())
    assert torch.equal(g_const, g_ones)                 # constant d logdet == a tensor of that constant
    with torch.no_grad():                               # dL/d(out) = scale * g_out (the prior's -1/var folded into the launch)
        g_scaled = _fused.layer_backward(lay, x, r * 4.0, s, 1.0, inverse, g_out_scale=0.25)
    assert torch.equal(g_scaled, gin)                   # power-of-two scale: exact
    # rows are independent: the first 128 rows alone give the same gradient bit for bit
    with torch.no_grad():
        g_head = _fused.layer_backward(lay, x[:128], r[:128], s[:128], 1.0, inverse)
    assert torch.equal(g_head, gin[:128])


def test_fused_force_path_is_used_and_matches_the_wide_path_and_the_oracle():
    """8-layer flow (H = 16 golden weights, bf16-class conditioner): FlowSimulation takes the one-launch-per-layer
    path (17 + 1 launches per evaluation); it agrees with the wide multi-launch path and with torch autograd
    through the fp32 oracle in the 1e-2 class."""
    from normalizingflow_b200 import _fused, _lib
    from normalizingflow_b200.hmc import FlowSimulation
    from oracle import nf_oracle as O
    m = _model(precision="bf16")
    g = golden("models.npz")
    x = T(g["nsf.x"])
    assert _fused.flow_logp_and_grad(m, x.cuda()) is not None
    sim = FlowSimulation(m, n_chains=x.shape[0], init_pos=x)
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    U, F = sim.potential_and_force(sim.get_position())
    torch.cuda.synchronize()
    n_launch = _lib.launch_count() - n0
    assert n_launch <= 20, n_launch
    sim.fused_grad = False
    U2, F2 = sim.potential_and_force(sim.get_position())
    scale = float(F2.abs().max())
    assert rel_err(U, U2.double().cpu()) <= 5e-3
    assert float((F - F2).abs().median()) <= 5e-3 * scale
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    specs = [dict(type="NSF_CL", size=32, dim=2, K=8, B=3.0, mask=[i % 2]) for i in range(8)]
    xr = x.clone().requires_grad_()
    z, plp, ld = O.flow_forward(specs, sd, xr)
    (gref,) = torch.autograd.grad((plp + ld).sum(), xr)
    assert rel_err(-U, (plp + ld).detach()) <= 1e-2
    err = (F.cpu() - gref).abs() / (1.0 + gref.abs())
    assert float(err.median()) <= 2e-3, float(err.median())
    assert float((err > 3e-2).float().mean()) <= 3e-2, float((err > 3e-2).float().mean())


def test_tile_chained_launches_equal_launch_by_launch_dependency_bitwise():
    """65,536 x 64 rows would be the production shape; 3 x 148 + 5 tiles are enough to have CTAs with different
    tile counts.  With tile flags a launch starts tile t when the producing launch has finished tile t (rows are
    independent): forces and log-probs must not change by a bit, with and without the log-prob reduction in the
    middle of the chain."""
    from normalizingflow_b200 import _fused
    m = _model(precision="bf16")
    N = (3 * 148 + 5) * 128
    x = 0.8 * torch.randn(N, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    try:
        _fused.TILE_CHAIN = False
        lp0, g0 = _fused.flow_logp_and_grad(m, x)
        _fused.TILE_CHAIN = "always"
        for _ in range(3):
            lp1, g1 = _fused.flow_logp_and_grad(m, x)
            lp2, g2 = _fused.flow_logp_and_grad(m, x, need_logp=False)
            assert lp2 is None
            assert torch.equal(lp0, lp1) and torch.equal(g0, g1) and torch.equal(g0, g2)
        torch.cuda.synchronize()
    finally:
        _fused.TILE_CHAIN = True


def test_force_at_the_full_chain_count_is_row_independent_and_chain_invariant():
    """BASELINE config 5's size (65,536 chains): a permuted batch gives the permuted log-probs and forces bit for bit
    (every row sees the same arithmetic wherever it sits: tile, CTA, wave), with tile-flag chains and without."""
    from normalizingflow_b200 import _fused
    m = _model(precision="bf16")
    C = 65536
    gen = torch.Generator(device="cuda").manual_seed(9)
    x = 0.8 * torch.randn(C, 64, device="cuda", generator=gen)
    perm = torch.randperm(C, device="cuda", generator=gen)
    lp, g = _fused.flow_logp_and_grad(m, x)
    lp_p, g_p = _fused.flow_logp_and_grad(m, x[perm].contiguous())
    assert torch.equal(lp[perm], lp_p) and torch.equal(g[perm], g_p)
    try:
        _fused.TILE_CHAIN = False
        lp_n, g_n = _fused.flow_logp_and_grad(m, x)
    finally:
        _fused.TILE_CHAIN = True
    assert torch.equal(lp, lp_n) and torch.equal(g, g_n)
    assert torch.isfinite(g).all() and torch.isfinite(lp).all()


def _trajectories(m, q0, p0, use_graph, fold, tile_chain, path_len, calls=2):
    from normalizingflow_b200 import _fused
    from normalizingflow_b200.hmc import FlowSimulation
    _fused.TILE_CHAIN = "always" if tile_chain else False      # "always": also below one wave of tiles
    try:
        sim = FlowSimulation(m, n_chains=q0.shape[0], init_pos=q0)
        sim.use_graph, sim.fused_leapfrog = use_graph, fold
        sim.set_velocity(p0)
        e0, out = sim.grad_evals, []
        for _ in range(calls):
            q, U = sim.integration_step(path_len=path_len, dt=0.01)
            out.append((q.clone(), U.clone(), sim.velocity.clone()))
        torch.cuda.synchronize()
        assert sim.grad_evals - e0 == calls * (path_len + 1)
        return out
    finally:
        _fused.TILE_CHAIN = True


@pytest.mark.parametrize("n_tiles", [5, 2 * 148 + 9])
def test_graph_replayed_tile_chains_reproduce_launch_by_launch_trajectories_bitwise(n_tiles):
    """Replayed as one CUDA graph the launches of an evaluation really overlap (in eager mode the host paces them): the
    tile-flag chain -- including the first backward launch hanging on the last forward launch, whose output is the
    backward's own x -- must give the launch-by-launch result bit for bit.  (Regression: the backward kernel used to
    prefetch the next tile's x before having seen that tile's flag.)"""
    m = _model(precision="bf16")
    C = n_tiles * 128              # 5 tiles: far fewer CTAs than SMs, i.e. many launches of the chain resident at once
    gen = torch.Generator(device="cuda").manual_seed(4)
    q0 = torch.randn(C, 64, device="cuda", generator=gen) * 0.5
    p0 = torch.randn(C, 64, device="cuda", generator=gen)
    ref = _trajectories(m, q0, p0, use_graph=False, fold=False, tile_chain=False, path_len=3)
    for use_graph in (False, True):
        got = _trajectories(m, q0, p0, use_graph=use_graph, fold=False, tile_chain=True, path_len=3)
        for a, b in zip(ref, got):
            assert all(torch.equal(x, y) for x, y in zip(a, b)), use_graph


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("n_tiles", [5, 148 + 7])
def test_leapfrog_folded_into_the_last_backward_launch_and_trajectory_as_one_chain(use_graph, n_tiles):
    """Kick / drift inside the launch that completes the force, the next evaluation's first forward launch hanging on
    it tile by tile (flags count evaluation epochs): (a) bit-identical to separate kick / drift launches for one step
    (no interior point); (b) over several steps bit-identical to the fold WITHOUT chaining consecutive evaluations
    (the two half kicks around an interior point are one fused multiply-add in both), also replayed as a graph with
    far fewer tiles than SMs, where launches of several evaluations are resident at once; (c) close to the separate
    launches over several steps (the dynamics amplify the rounding difference of (b))."""
    from normalizingflow_b200 import _fused
    m = _model(precision="bf16")
    C = n_tiles * 128
    gen = torch.Generator(device="cuda").manual_seed(4)
    q0 = torch.randn(C, 64, device="cuda", generator=gen) * 0.5
    p0 = torch.randn(C, 64, device="cuda", generator=gen)
    sep1 = _trajectories(m, q0, p0, use_graph=False, fold=False, tile_chain=True, path_len=1)
    sep6 = _trajectories(m, q0, p0, use_graph=False, fold=False, tile_chain=True, path_len=6)
    try:
        _fused.CHAIN_EVALS = False
        fold6 = _trajectories(m, q0, p0, use_graph=False, fold=True, tile_chain=True, path_len=6)
    finally:
        _fused.CHAIN_EVALS = True
    one1 = _trajectories(m, q0, p0, use_graph=use_graph, fold=True, tile_chain=True, path_len=1)
    one6 = _trajectories(m, q0, p0, use_graph=use_graph, fold=True, tile_chain=True, path_len=6)
    for a, b in zip(sep1, one1):
        assert all(torch.equal(x, y) for x, y in zip(a, b))
    for a, b in zip(fold6, one6):
        assert all(torch.equal(x, y) for x, y in zip(a, b))
    for a, b in zip(sep6, one6):
        for x, y in zip(a, b):
            assert torch.isfinite(y).all()
            assert float((x - y).abs().max()) <= 2e-2 * (1.0 + float(x.abs().max())), float((x - y).abs().max())


def test_single_chain_eager_rejection_restores_the_position():
    """n_chains = 1 on the eager integrator (fp32 conditioner): HMC keeps references to the position it
    last accepted (nf/hmc.py:36, :58), so the integrator must not advance that tensor in place -- a
    rejected move leaves position and potential exactly where they were."""
    from normalizingflow_b200.hmc import HMC, FlowSimulation
    m = _model()
    q0 = torch.randn(1, 64, generator=torch.Generator().manual_seed(3)) * 0.3
    sim = FlowSimulation(m, n_chains=1, nparticles=32, dim=2, init_pos=q0)
    sim.tensor_core_grad = False
    # the integrator works on copies: a reference taken before the step still holds the old position
    held = sim.get_position()
    snapshot = held.clone()
    sim.set_velocity(torch.randn(1, 64, generator=torch.Generator().manual_seed(4)))
    q1, _ = sim.integration_step(path_len=3, dt=0.2)
    assert torch.equal(held, snapshot) and not torch.equal(q1, snapshot)
    assert sim.get_position() is q1
    sim.set_position(snapshot)
    # through the driver: starting near the mode with full-temperature velocities most moves go uphill
    # and are rejected (the reference's acceptance rule ignores the kinetic energy, Q12)
    h = HMC(sim, path_len=3, dt=0.2, dim=2, beta=1.0)
    torch.manual_seed(0)
    pos, pot, logp, acc = h.hmc(epochs=10)
    assert pos.shape == (10, 64) and 0.0 <= acc <= 1.0
    n_rej = 0
    for i in range(1, 10):
        same_pos, same_pot = torch.equal(pos[i], pos[i - 1]), float(pot[i]) == float(pot[i - 1])
        assert same_pos == same_pot, (i, same_pos, same_pot)      # a recorded position moves only when accepted
        n_rej += int(same_pos)
    assert n_rej >= 1, "test needs at least one rejection to be meaningful"
    assert torch.equal(pos[0].cpu(), snapshot.flatten().cpu())
    # the simulation's own state equals HMC's accepted state
    assert torch.equal(sim.get_position().flatten().cpu(), torch.as_tensor(h.position).flatten().cpu())


def test_mass_enters_the_dynamics_once():
    """FlowSimulation(mass=m): HMC hands over velocities v ~ N(0, 1/(m beta)); the integrator must reproduce
    velocity Verlet  q += dt v + dt^2/(2m) F ...  -- i.e. a trajectory with mass m and velocity v equals the
    unit-mass trajectory with velocity v and forces F/m."""
    from normalizingflow_b200.hmc import FlowSimulation
    m = _model()
    C, mass = 64, 4.0
    gen = torch.Generator(device="cuda").manual_seed(2)
    q0 = torch.randn(C, 64, device="cuda", generator=gen) * 0.5
    v0 = torch.randn(C, 64, device="cuda", generator=gen) * 0.5
    sim = FlowSimulation(m, n_chains=C, init_pos=q0, mass=mass)
    sim.set_velocity(v0)
    dt = 0.05
    U, F = sim.potential_and_force(q0)
    q1, _ = sim.integration_step(path_len=1, dt=dt)
    expect = q0 + dt * v0 + 0.5 * dt * dt * F / mass
    assert float((q1 - expect).abs().max()) < 1e-5
    with pytest.raises(ValueError):
        FlowSimulation(m, n_chains=C, init_pos=q0, mass=torch.tensor([1.0, 2.0]))
