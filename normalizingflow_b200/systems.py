"""Priors / targets evaluated next to the flow (applications/src/systems.py) on libnfk:
``EinsteinCrystal`` (prior of the shipped LJ / Fe experiments, systems.py:340-372), ``LJ`` pair
potential with minimum image and cutoff/shift (systems.py:144-189) and ``GaussianMixture``
(systems.py:257-296).  Constructor arguments follow the reference; ``centers`` must be given as a
tensor / nested list (the reference's XYZ-file loader needs MDAnalysis, which is out of scope).
All tensors live on a CUDA device."""
from __future__ import annotations

import math

import torch

from ._lib import call, f32c, ptr, require_cuda, stream_ptr


class _EinsteinLogProb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, centers, natoms, dim, alpha, boxlength):
        dev = require_cuda(x, centers)
        xf = f32c(x.reshape(-1, natoms * dim))
        N = xf.shape[0]
        out = torch.empty(N, dtype=torch.float32, device=dev)
        gx = torch.empty_like(xf) if x.requires_grad else None
        with torch.cuda.device(dev):
            call("nfk_einstein_logprob", ptr(xf), ptr(centers), ptr(out), ptr(gx), N, natoms, dim, float(alpha),
                 float(boxlength or 0.0), stream_ptr(dev))
        ctx.gx = gx
        ctx.shape = x.shape
        return out

    @staticmethod
    def backward(ctx, g):
        return (g[:, None] * ctx.gx).reshape(ctx.shape), None, None, None, None, None


class EinsteinCrystal:
    def __init__(self, centers, dim=3, boxlength=None, alpha=50, device="cuda"):
        self.device = torch.device(device)
        if isinstance(centers, str):
            raise NotImplementedError("loading centers from an XYZ file needs MDAnalysis; pass a tensor")
        self.centers = torch.as_tensor(centers, dtype=torch.float32).reshape(-1, dim).to(self.device).contiguous()
        self.natoms = self.centers.shape[0]
        self.alpha = alpha
        self.dim = dim
        self.boxlength = boxlength
        self.generator = None

    def sample(self, nsamples, flatten=True):
        """centers + N(0, I/alpha) noise, wrapped into the box (systems.py:353-359)."""
        with torch.no_grad():
            if isinstance(nsamples, tuple):
                nsamples = nsamples[0]
            noise = torch.randn((nsamples, self.natoms, self.dim), device=self.device, generator=self.generator)
            samples = self.centers + noise * math.sqrt(1.0 / self.alpha)
            if self.boxlength is not None:
                samples = samples - (samples.abs() > 0.5 * self.boxlength) * torch.sign(samples) * self.boxlength
            return samples.reshape(nsamples, -1) if flatten else samples

    def log_prob(self, x):
        return _EinsteinLogProb.apply(x, self.centers, self.natoms, self.dim, self.alpha, self.boxlength)

    def potential(self, x):
        return -self.log_prob(x)


class _LJPotential(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, n, dim, boxlength, epsilon, sigma, cutoff, shift):
        dev = require_cuda(pos)
        pf = f32c(pos.reshape(-1, n * dim))
        N = pf.shape[0]
        out = torch.empty(N, dtype=torch.float32, device=dev)
        g = torch.empty_like(pf) if pos.requires_grad else None
        with torch.cuda.device(dev):
            call("nfk_lj_potential", ptr(pf), ptr(out), ptr(g), N, n, dim, float(boxlength or 0.0), float(epsilon),
                 float(sigma), float(cutoff or 0.0), int(bool(shift)), stream_ptr(dev))
        ctx.g = g
        ctx.shape = pos.shape
        return out

    @staticmethod
    def backward(ctx, go):
        return (go[:, None] * ctx.g).reshape(ctx.shape), None, None, None, None, None, None, None


class LJ:
    def __init__(self, pos_dir=None, boxlength=None, device="cuda", epsilon=1., sigma=1., cutoff=None, shift=True):
        if pos_dir is not None:
            raise NotImplementedError("trajectory loading needs MDAnalysis; pass positions to potential()")
        self.device = torch.device(device)
        self.epsilon, self.sigma, self.cutoff, self.shift, self.boxlength = epsilon, sigma, cutoff, shift, boxlength

    def potential(self, particle_pos):
        """particle_pos [..., nparticles, dim] -> total potential [...] (systems.py:154-189)."""
        n, dim = particle_pos.shape[-2], particle_pos.shape[-1]
        lead = particle_pos.shape[:-2]
        out = _LJPotential.apply(particle_pos.reshape(-1, n, dim), n, dim, self.boxlength, self.epsilon, self.sigma,
                                 self.cutoff, self.shift)
        return out.reshape(lead)

    def log_prob(self, x, nparticles=None, dim=3):
        if x.dim() == 2 and nparticles is not None:
            x = x.reshape(x.shape[0], nparticles, dim)
        return -self.potential(x)


class GaussianMixture:
    def __init__(self, centers, vars, npoints=None, dim=3, device="cuda"):
        self.dim = dim
        self.device = torch.device(device)
        if isinstance(centers, str):
            raise NotImplementedError("loading centers from an XYZ file needs MDAnalysis; pass a tensor")
        self.centers = torch.as_tensor(centers, dtype=torch.float32).reshape(-1, dim).to(self.device).contiguous()
        self.ncenters = len(self.centers)
        v = torch.as_tensor(vars, dtype=torch.float32).to(self.device)
        self.vars = (v.expand(self.ncenters) if v.dim() == 0 else v).contiguous()
        self.nparticles = self.ncenters if npoints is None else npoints
        self.generator = None

    def sample(self, nsamples, flatten=True):
        with torch.no_grad():
            if isinstance(nsamples, tuple):
                nsamples = nsamples[0]
            m = nsamples * self.nparticles
            which = torch.randint(0, self.ncenters, (m,), device=self.device, generator=self.generator)
            noise = torch.randn((m, self.dim), device=self.device, generator=self.generator)
            pts = self.centers[which] + noise * torch.sqrt(self.vars[which])[:, None]
            return pts.reshape(nsamples, -1) if flatten else pts.reshape(nsamples, self.nparticles, self.dim)

    def log_prob(self, x):
        dev = require_cuda(x)
        xf = f32c(x.reshape(-1, self.nparticles * self.dim))
        N = xf.shape[0]
        out = torch.empty(N, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            call("nfk_gmm_logprob", ptr(xf), ptr(self.centers), ptr(self.vars), ptr(out), N, self.nparticles, self.dim,
                 self.ncenters, stream_ptr(dev))
        return out

    def potential(self, x):
        return -self.log_prob(x)
