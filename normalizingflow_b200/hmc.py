"""HMC driver with the reference's interface (nf/hmc.py:8-65) plus a batched, flow-backed
``simulation`` object.

``HMC`` drives any object that offers the reference's duck type (hmc.py:15-19, 39-40, 48-50, 63):
``nparticles``, ``get_position()``, ``get_potential()``, ``set_position(p)``, ``set_velocity(v)``,
``integration_step(path_len, dt) -> (position, potential)``.  As in the reference, velocities
are drawn from N(0, diag(1/m)/init_beta) and the Metropolis test uses the potential only
(quirk Q12).  When the simulation exposes ``n_chains > 1`` every chain is advanced and
accepted/rejected independently in one batched step.

``FlowSimulation`` is the flow-preconditioned target of BASELINE config 5: potential
U(q) = -log p(q) under a NormalizingFlowModel, force = grad_q log p(q) through the flow's custom
backward kernels, velocity-Verlet integration with the kick/drift kernels of libnfk.  Unlike the
reference's pure-torch integrators (applications/src/systems.py:331-336, quirk Q13) the force is
re-evaluated at the new position every step.
"""
from __future__ import annotations

import math

import torch
from torch.distributions import MultivariateNormal

from . import _lib, _ops


class HMC:
    def __init__(self, simulation, init_pos=None, path_len=1, dt=None, mass=None, dim=3, beta=1.0, init_beta=None):
        self.simulation = simulation
        self.dt = dt
        self.path_len = path_len
        self.beta = beta
        self.dim = dim
        self.nparticles = simulation.nparticles
        self.n_chains = int(getattr(simulation, "n_chains", 1))
        self.position = simulation.get_position()
        self.potential = simulation.get_potential()
        if init_pos is not None:
            self.simulation.set_position(init_pos)
        self.mass = torch.ones(self.nparticles) if mass is None else mass
        if init_beta is None:
            init_beta = beta
        self.init_beta = init_beta
        # per-coordinate 1/m, particle-major like the reference (hmc.py:23)
        self._inv_mass = (1 / self.mass).expand(self.dim, self.nparticles).transpose(0, 1).flatten()
        self.v_dist = MultivariateNormal(torch.zeros(self.dim * self.nparticles),
                                         torch.diag(self._inv_mass) / init_beta)

    # ---- velocities ---------------------------------------------------------------------
    def generate_v(self):
        if self.n_chains == 1:
            v = self.v_dist.sample((1,))
            return v.flatten(), self.v_dist.log_prob(v)
        dev = self.simulation.device
        std = self._v_consts(dev)[0]
        v = torch.randn(self.n_chains, self.dim * self.nparticles, device=dev,
                        generator=getattr(self.simulation, "generator", None)) * std
        return v, self._log_prob_v(v)

    def _v_consts(self, dev):
        """(std, 1 / var, log-normaliser) of the velocity distribution on `dev`, uploaded once: a host -> device copy from
        pageable memory inside the epoch loop blocks the host until the stream has drained, i.e. once per epoch."""
        key = (str(dev), float(self.init_beta))
        c = getattr(self, "_v_cache", None)
        if c is None or c[0] != key:
            var = (self._inv_mass / self.init_beta).to(torch.float32)
            c = self._v_cache = (key, (torch.sqrt(var).to(dev), (1.0 / var).to(dev),
                                       float(-0.5 * torch.log(2 * math.pi * var).sum())))
        return c[1]

    def _log_prob_v(self, v):
        _, inv_var, log_norm = self._v_consts(v.device)
        return -0.5 * (v * v * inv_var).sum(-1) + log_norm

    def run_sim(self, v=None):
        if v is None:
            v, log_prob = self.generate_v()
        else:
            log_prob = self.v_dist.log_prob(v) if self.n_chains == 1 else self._log_prob_v(v)
        self.simulation.set_velocity(v)
        position, potential = self.simulation.integration_step(self.path_len, self.dt)
        return position, potential, log_prob

    # ---- chain --------------------------------------------------------------------------
    def hmc(self, epochs=1, init_pos=None):
        if init_pos is not None:
            self.simulation.set_position(init_pos)
            self.position = init_pos
            self.potential = self.simulation.get_potential()
        if self.n_chains > 1:
            return self._hmc_batched(epochs)
        positions, potentials, naccept, log_prob = [], [], 0, None
        for _ in range(epochs):
            positions.append(torch.as_tensor(self.position).detach().clone().float().flatten())
            potentials.append(self.potential)
            position, potential, log_prob = self.run_sim()
            acc_prob = math.exp((self.potential - potential) * self.beta)      # potential only (Q12)
            if torch.rand(1) < acc_prob:
                self.position = position.flatten()
                self.potential = potential
                naccept += 1
            else:
                self.simulation.set_position(self.position)
        return (torch.stack(positions), torch.tensor([float(p) for p in potentials]),
                torch.as_tensor(log_prob), naccept / epochs)

    def _hmc_batched(self, epochs):
        sim = self.simulation
        pos = sim.get_position().clone()
        pot = sim.get_potential().clone()
        # recorded positions / potentials go straight into their result tensors, the acceptance count stays on the device:
        # no host synchronisation inside the loop, so the next epoch's launches queue up behind the running trajectory
        positions = torch.empty((epochs,) + tuple(pos.shape), dtype=pos.dtype, device=pos.device)
        potentials = torch.empty((epochs,) + tuple(pot.shape), dtype=pot.dtype, device=pot.device)
        accepted = torch.zeros((), dtype=torch.float32, device=pos.device)
        log_prob = None
        for e in range(epochs):
            positions[e].copy_(pos)
            potentials[e].copy_(pot)
            new_pos, new_pot, log_prob = self.run_sim()
            u = torch.rand(self.n_chains, device=pos.device, generator=getattr(sim, "generator", None))
            acc = u < torch.exp((pot - new_pot) * self.beta)
            pos = torch.where(acc[:, None], new_pos, pos)
            pot = torch.where(acc, new_pot, pot)
            accepted += acc.float().mean()
            sim.set_position(pos)
        self.position, self.potential = pos, pot
        return positions, potentials, log_prob, float(accepted) / epochs


class FlowSimulation:
    """Batched ``simulation`` for ``HMC``: ``n_chains`` independent chains in the d-dimensional
    space of a NormalizingFlowModel, potential U = -log p_flow."""

    def __init__(self, model, n_chains, nparticles=None, dim=None, init_pos=None, mass=1.0, generator=None):
        self.model = model
        self.n_chains = int(n_chains)
        p = next(model.parameters())
        self.device = p.device
        self.generator = generator
        if init_pos is None:
            init_pos = model.sample(self.n_chains)[0]          # chains start from flow samples (dynamics.py:60-62)
        self.position = init_pos.to(self.device, torch.float32).reshape(self.n_chains, -1).contiguous().clone()
        self.d = self.position.shape[1]
        self.dim = dim if dim is not None else 1
        self.nparticles = nparticles if nparticles is not None else self.d // self.dim
        if isinstance(mass, torch.Tensor):
            if mass.numel() != 1 and not bool((mass == mass.flatten()[0]).all()):
                raise ValueError("FlowSimulation integrates one scalar mass for all coordinates; per-particle "
                                 "masses (HMC(mass=tensor)) must all be equal")
            mass = float(mass.flatten()[0])
        self.mass = float(mass)
        if self.mass <= 0.0:
            raise ValueError("mass must be positive")
        self.inv_mass = 1.0 / self.mass
        self.velocity = torch.zeros_like(self.position)
        self.grad_evals = 0
        self.tensor_core_grad = True      # bf16-conditioner models: hand-written forward+backward path
        self.fused_grad = True            # ... through the one-launch-per-layer kernels when hidden width <= 128
        self.fused_leapfrog = True        # ... with kick / drift folded into the last backward launch of every evaluation and
        #                                   the next evaluation's first launch hanging on it: one chain per trajectory
        self.use_graph = True             # replay whole trajectories as one CUDA graph on that path
        self._graphs = {}

    # duck type ---------------------------------------------------------------------------
    def get_position(self):
        return self.position

    def set_position(self, position):
        self.position = position.to(self.device, torch.float32).reshape(self.n_chains, -1).contiguous().clone()

    def set_velocity(self, velocity):
        """``velocity`` is what HMC.generate_v draws: v ~ N(0, 1/(m beta)) (hmc.py:24-27).  The kick / drift
        kernels integrate the MOMENTUM p = m v (kick p += dt/2 F, drift q += dt p / m), so the velocity is
        converted here; with the default unit mass the two coincide."""
        v = velocity.to(self.device, torch.float32).reshape(self.n_chains, -1).contiguous().clone()
        self.velocity = v * self.mass if self.mass != 1.0 else v

    def potential(self, x):
        return -self.model.evaluate(x)

    def get_potential(self, x=None):
        return self.potential(self.position if x is None else x)

    def potential_and_force(self, q, need_potential=True):
        """U(q) [C] and F(q) = -grad U = grad log p [C, d] — one forward + one backward through the flow.  With
        need_potential = False (the interior evaluations of a trajectory) the one-launch-per-layer path skips the
        log-prob reduction and returns None for U."""
        from . import _fused, _wide
        fast = None
        if self.tensor_core_grad:
            # hidden width <= 128: one forward and one backward launch per layer (csrc/nsf_fused2.cu,
            # csrc/nsf_fused_bwd.cu); wider conditioners: the forward keeps the hidden activations, backward =
            # spline adjoint as a GEMM epilogue + dgrad GEMMs on the tensor cores (csrc/gemm_ws.cu); no autograd graph
            fast = _fused.flow_logp_and_grad(self.model, q, need_potential) if self.fused_grad else None
            if fast is None:
                fast = _wide.flow_logp_and_grad(self.model, q)
        if fast is not None:
            logp, force = fast
            self.grad_evals += 1
            return (-logp if logp is not None else None), force
        params = [p for p in self.model.parameters() if p.requires_grad]
        for p in params:
            p.requires_grad_(False)              # dgrad only: no weight gradients in the leapfrog
        try:
            with torch.enable_grad():
                x = q.detach().requires_grad_(True)
                z, prior_lp, log_det = self.model.forward(x)
                logp = prior_lp + log_det
                (force,) = torch.autograd.grad(logp.sum(), x)
        finally:
            for p in params:
                p.requires_grad_(True)
        self.grad_evals += 1
        return -logp.detach(), force.contiguous()

    # ---- whole trajectory as one CUDA graph -------------------------------------------------
    def _fused_leapfrog_ok(self, q):
        from . import _fused
        return (self.tensor_core_grad and self.fused_grad and self.fused_leapfrog and q.dim() == 2 and q.shape[1] == 64
                and q.is_contiguous() and _lib.have("nfk_nsf_pairs_fused_bwd_leapfrog")
                and _fused.tile_chain_ok(self.model, q.shape[0]) and _fused.flow_eligible(self.model))

    def _trajectory_fused(self, q, p, path_len, dt):
        """The whole trajectory as ONE chain of layer launches: the launch that completes the force of an evaluation
        (backward of the layer nearest the data) also does the kick and the drift of its rows, and the next evaluation's
        first forward launch takes a 128-row tile as soon as that tile's position has been advanced.  Only the end
        point's potential is computed.  p receives dt/2 F(q_0), dt F(q_i) for the interior points and dt/2 F(q_L) --
        the two half kicks around an interior point are one fused multiply-add here."""
        from . import _fused
        logp = None
        for i in range(path_len + 1):
            last = i == path_len
            kick = 0.5 * dt if (i == 0 or last) else dt
            drift = 0.0 if last else dt * self.inv_mass
            logp, _ = _fused.flow_logp_and_grad(self.model, q, need_logp=last, leapfrog=(p, q, kick, drift),
                                                hang_on_previous=(i > 0 and _fused.CHAIN_EVALS),
                                                next_hangs_on_this=(not last and _fused.CHAIN_EVALS),
                                                zero_flags=(i == 0 or not _fused.CHAIN_EVALS),
                                                epoch=(i + 1 if _fused.CHAIN_EVALS else 1))
            self.grad_evals += 1
        return -logp

    def _trajectory(self, q, p, path_len, dt):
        if path_len >= 1 and self._fused_leapfrog_ok(q):
            return self._trajectory_fused(q, p, path_len, dt)
        pot, force = self.potential_and_force(q, need_potential=(path_len == 0))
        for i in range(path_len):
            _ops.leapfrog_kick_drift(q, p, force, dt, self.inv_mass)
            pot, force = self.potential_and_force(q, need_potential=(i == path_len - 1))   # only the end point's U is used
            _ops.leapfrog_kick(p, force, dt)
        return pot

    def _graph_trajectory(self, path_len, dt):
        """The leapfrog trajectory (path_len + 1 log-prob+grad evaluations: 18 launches each at hidden width <= 128, ~75 on the wide path) is a
        fixed launch sequence on fixed buffers: capture it once per (path_len, dt, parameter
        version) and replay it, so the GPU is never waiting on the host."""
        from . import _wide
        if not _wide.flow_grad_eligible(self.model):
            return None
        key = (path_len, dt, self.n_chains, self.fused_grad, self.fused_leapfrog, _lib.param_epoch(),
               tuple((p._version, p.data_ptr()) for p in self.model.parameters()))
        entry = self._graphs.get(key)
        if entry is None:
            self._graphs.clear()
            q = self.position.clone()
            p = self.velocity.clone()
            evals = self.grad_evals
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                       # warm-up: weight images, kernel attributes
                self._trajectory(q.clone(), p.clone(), 1, dt)
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                pot = self._trajectory(q, p, path_len, dt)
            self.grad_evals = evals
            entry = self._graphs[key] = (graph, q, p, pot)
        graph, q, p, pot = entry
        q.copy_(self.position)
        p.copy_(self.velocity)
        graph.replay()
        self.grad_evals += path_len + 1
        self.position = q.clone()
        self.velocity = p.clone()
        return self.position, pot.clone()

    def integration_step(self, path_len=1, dt=0.005, init_pos=None, init_velocity=None):
        if init_pos is not None:
            self.set_position(init_pos)
        if init_velocity is not None:
            self.set_velocity(init_velocity)
        if dt is None:
            dt = 0.005
        if self.tensor_core_grad and self.use_graph:
            out = self._graph_trajectory(int(path_len), float(dt))
            if out is not None:
                return out
        # integrate on copies: HMC keeps references to the position it last accepted (hmc.py:36, :58),
        # so advancing self.position in place would turn every rejection into an acceptance
        q, p = self.position.clone(), self.velocity.clone()
        pot = self._trajectory(q, p, int(path_len), float(dt))
        self.position, self.velocity = q, p
        return q, pot
