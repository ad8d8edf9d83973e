"""NormalizingFlowModel with the reference's interface (nf/models.py:5-40).

    forward(x)  -> (z, prior_logprob, log_det)        models.py:13-20
    inverse(z)  -> (x, log_det)                       models.py:22-29
    sample(n)   -> (x, log_px, z)   detached          models.py:31-35
    evaluate(x) -> log_px           detached          models.py:37-40

Differences that do not change results: under ``torch.no_grad()`` each layer accumulates its
log-det into the running [N] buffer inside its kernel (no separate add pass); consecutive
``Planar`` layers run as one fused stack kernel; a ``GaussianPrior`` is evaluated by the
log-prob reduction kernel.  Any other prior object with ``.sample((n,))`` / ``.log_prob(x)``
(e.g. torch.distributions.MultivariateNormal) is used as is.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _ops
from .flows import Planar, PlanarStack


class GaussianPrior:
    """N(0, var*I_d) — the prior the reference's hot configs build with
    MultivariateNormal(0, vars*I) (applications/src/setup.py:25-30) — on the log-prob kernel."""

    def __init__(self, dim, var=1.0, device="cuda"):
        self.dim = dim
        self.var = float(var)
        self.device = torch.device(device)
        self.generator = None

    def sample(self, sample_shape=torch.Size()):
        shape = tuple(sample_shape) + (self.dim,)
        z = torch.randn(shape, device=self.device, dtype=torch.float32, generator=self.generator)
        return z * math.sqrt(self.var) if self.var != 1.0 else z

    def log_prob(self, x):
        flat = x.reshape(-1, self.dim)
        if torch.is_grad_enabled() and flat.requires_grad:
            out = _ops.GaussLogprobFn.apply(flat, self.var)
        else:
            out = _ops.gauss_logprob(flat, self.var)
        return out.reshape(x.shape[:-1])


class NormalizingFlowModel(nn.Module):

    def __init__(self, prior, flows, device="cpu", fuse_planar=True):
        super().__init__()
        self.device = device
        self.prior = prior
        self.flows = nn.ModuleList(flows)
        self.fuse_planar = fuse_planar

    # runs of consecutive Planar layers collapse into one fused launch
    def _forward_plan(self):
        plan, run = [], []
        for f in self.flows:
            if self.fuse_planar and type(f) is Planar:
                run.append(f)
                continue
            if run:
                plan.append(PlanarStack(run) if len(run) > 1 else run[0])
                run = []
            plan.append(f)
        if run:
            plan.append(PlanarStack(run) if len(run) > 1 else run[0])
        return plan

    def forward(self, x):
        m, _ = x.shape
        log_det = torch.zeros(m, dtype=torch.float32, device=x.device)     # fp32 always (quirk Q11)
        fused = not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())))
        for flow in self._forward_plan():
            if fused and hasattr(flow, "_nfk_step"):
                x, log_det = flow._nfk_step(x, log_det, False)
            else:
                x, ld = flow.forward(x)
                log_det = log_det + ld
        z, prior_logprob = x, self.prior.log_prob(x)
        return z, prior_logprob, log_det

    def inverse(self, z):
        m, _ = z.shape
        log_det = torch.zeros(m, dtype=torch.float32, device=z.device)
        fused = not (torch.is_grad_enabled() and (z.requires_grad or any(p.requires_grad for p in self.parameters())))
        for flow in self.flows[::-1]:
            if fused and hasattr(flow, "_nfk_step"):
                z, log_det = flow._nfk_step(z, log_det, True)
            else:
                z, ld = flow.inverse(z)
                log_det = log_det + ld
        x = z
        return x, log_det

    def sample(self, n_samples):
        with torch.no_grad():
            z = self.prior.sample((n_samples,))
            x, log_det = self.inverse(z)
            log_px = self.prior.log_prob(z) - log_det
        return x.data, log_px.data, z.data

    def evaluate(self, x):
        with torch.no_grad():
            z, prior_logprob, log_det = self.forward(x)
            log_px = prior_logprob + log_det
        return log_px.data
