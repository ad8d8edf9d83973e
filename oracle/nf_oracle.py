"""CPU oracle for the normalizing-flow transform hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``normalizingflow_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and there only as the checker / the CPU arm.

What it is
----------
A restatement, in plain eager PyTorch on the host, of the arithmetic of
sherryli59/NormalizingFlow's transform path.  The reference's arithmetic *is* a chain of
ATen ops (softmax, cumsum, softplus, log, sqrt, addmm, tanh), so the restatement calls the
same ATen primitives in the same order and therefore rounds identically on the same host.
It is written against explicit tensors (no nn.Module state, no boolean-mask gathers, no
host syncs) and is dtype-agnostic: calling it with float64 tensors gives the "fp64 twin"
used for error budgeting.

Parity pin
----------
The reference ships no tests or golden vectors, so the oracle is pinned to
outputs of the reference itself: ``oracle/gen_golden.py`` imports the unmodified reference
from /root/reference (build container only), runs it on seeded inputs and stores inputs and
outputs under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this module
against those fixtures (bit-exact bins; values to 1e-6).

Reference lines followed (paths relative to the reference root):
  nf/utils.py:20-25    searchsorted           -> _count_bin
  nf/utils.py:27-56    unconstrained_RQS      -> rqs_elementwise (tails, derivative padding)
  nf/utils.py:58-152   RQS                    -> _knots, rqs_elementwise
  nf/flows.py:20-35    FCNN                   -> fcnn
  nf/flows.py:52-76    RealNVP fwd/inv        -> realnvp
  nf/flows.py:227-253  NSF_CL fwd/inv         -> nsf_cl_split, nsf_cl_transform, nsf_cl
  nf/flows_1.py:42-60  Planar.forward         -> planar
  nf/flows_1.py:85-97  Radial.forward         -> radial
  nf/models.py:13-40   NormalizingFlowModel   -> flow_forward, flow_inverse, gauss_logprob
  nf/hmc.py:8-65       HMC                    -> hmc_chain (acceptance rule Q12)
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

MIN_BIN = 1e-3      # nf/utils.py:13-14
MIN_DERIV = 1e-3    # nf/utils.py:15


def edge_derivative_constant() -> float:
    """float64 value written into the padded derivative slots (nf/utils.py:37-40)."""
    return float(np.log(np.exp(1 - MIN_DERIV) - 1))


# --------------------------------------------------------------------------------------
# RQS spline
# --------------------------------------------------------------------------------------
def _knots(unnorm: torch.Tensor, lo: float, hi: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """softmax -> min-size affine -> cumsum -> rescale -> pinned end points.

    Returns (knots[..., K+1], sizes[..., K]).  nf/utils.py:73-80 (widths), :84-91 (heights).
    """
    K = unnorm.shape[-1]
    p = F.softmax(unnorm, dim=-1)
    p = MIN_BIN + (1 - MIN_BIN * K) * p
    c = torch.cumsum(p, dim=-1)
    c = F.pad(c, pad=(1, 0), mode="constant", value=0.0)
    c = (hi - lo) * c + lo
    c[..., 0] = lo
    c[..., -1] = hi
    return c, c[..., 1:] - c[..., :-1]


def _count_bin(knots: torch.Tensor, v: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """#{j: v >= knots_j} - 1 with the last knot nudged by eps (nf/utils.py:20-25)."""
    nudged = knots.clone()
    nudged[..., -1] += eps
    return torch.sum(v[..., None] >= nudged, dim=-1) - 1


def rqs_elementwise(v: torch.Tensor, W1: torch.Tensor, H1: torch.Tensor, D1: torch.Tensor,
                    inverse: bool, B: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Unconstrained rational-quadratic spline with identity tails.

    ``v``: [...]; ``W1``, ``H1``: [..., K]; ``D1``: [..., K-1] — the values *handed to*
    ``unconstrained_RQS`` (i.e. already 2B*softmax / softplus'ed once by the layer; the
    spline normalises them a second time, quirks Q1/Q2).
    Returns (out, logabsdet, bin) with ``bin = -1`` in the tails.
    """
    inside = (v >= -B) & (v <= B)                                   # utils.py:32 (inclusive)
    # utils.py:36-40: pad the K-1 derivatives with the constant on both sides.
    D2 = F.pad(D1, pad=(1, 1))
    c = edge_derivative_constant()
    D2[..., 0] = c
    D2[..., -1] = c

    # tails are evaluated on a harmless in-range value and discarded afterwards
    vv = torch.where(inside, v, torch.zeros_like(v))

    cw, w = _knots(W1, -B, B)
    ch, h = _knots(H1, -B, B)
    d = MIN_DERIV + F.softplus(D2)                                  # utils.py:82
    k = _count_bin(ch if inverse else cw, vv)[..., None]            # utils.py:93-96
    k = k.clamp(0, W1.shape[-1] - 1)

    def pick(t):
        return t.gather(-1, k)[..., 0]

    cw_k, w_k, ch_k, h_k = pick(cw), pick(w), pick(ch), pick(h)
    delta_k = pick(h / w)                                           # utils.py:102-103
    d_k, d_k1 = pick(d), pick(d[..., 1:])

    if inverse:                                                     # utils.py:111-135
        u = vv - ch_k
        s = d_k + d_k1 - 2 * delta_k
        a = u * s + h_k * (delta_k - d_k)
        b = h_k * d_k - u * s
        cc = -delta_k * u
        disc = b.pow(2) - 4 * a * cc
        root = (2 * cc) / (-b - torch.sqrt(disc))
        out = root * w_k + cw_k
        t = root * (1 - root)
        den = delta_k + s * t
        dnum = delta_k.pow(2) * (d_k1 * root.pow(2) + 2 * delta_k * t
                                 + d_k * (1 - root).pow(2))
        lad = -(torch.log(dnum) - 2 * torch.log(den))
    else:                                                           # utils.py:137-152
        theta = (vv - cw_k) / w_k
        t = theta * (1 - theta)
        num = h_k * (delta_k * theta.pow(2) + d_k * t)
        den = delta_k + ((d_k + d_k1 - 2 * delta_k) * t)
        out = ch_k + num / den
        dnum = delta_k.pow(2) * (d_k1 * theta.pow(2) + 2 * delta_k * t
                                 + d_k * (1 - theta).pow(2))
        lad = torch.log(dnum) - 2 * torch.log(den)

    out = torch.where(inside, out, v)                               # utils.py:42-43
    lad = torch.where(inside, lad, torch.zeros_like(lad))
    bins = torch.where(inside, k[..., 0], torch.full_like(k[..., 0], -1))
    return out, lad, bins


# --------------------------------------------------------------------------------------
# NSF_CL coupling layer
# --------------------------------------------------------------------------------------
def _mask_lists(dim: int, mask: Sequence[int]) -> Tuple[List[int], List[int]]:
    mask = [int(m) for m in mask]
    return mask, [c for c in range(dim) if c not in mask]           # flows.py:224-225


def nsf_cl_split(x: torch.Tensor, size: int, dim: int, mask: Sequence[int]):
    """flows.py:229-230: conditioning ("lower") and transformed ("upper") columns."""
    m, u = _mask_lists(dim, mask)
    x3 = x.reshape(-1, size, dim)
    lower = x3[:, :, m].flatten(start_dim=1)
    upper = x3[:, :, u].flatten(start_dim=1)
    return lower, upper


def nsf_cl_transform(x: torch.Tensor, params: torch.Tensor, size: int, dim: int,
                     mask: Sequence[int], K: int, B: float, inverse: bool):
    """The part of NSF_CL.forward/inverse after ``psi`` (flows.py:231-239 / :245-253).

    ``params``: raw conditioner output [N, F_t, 3K-1] (F_t = size*(dim-len(mask))).
    Returns (out [N, size*dim] in the reference's column order (Q5), log_det [N],
    bins [N, F_t] int64).
    """
    m, u = _mask_lists(dim, mask)
    lower, upper = nsf_cl_split(x, size, dim, mask)
    Wr, Hr, Dr = torch.split(params, K, dim=2)                      # flows.py:232
    W1 = 2 * B * torch.softmax(Wr, dim=2)                           # flows.py:233-234
    H1 = 2 * B * torch.softmax(Hr, dim=2)
    D1 = F.softplus(Dr)                                             # flows.py:235
    new_upper, lad, bins = rqs_elementwise(upper, W1, H1, D1, inverse, B)
    log_det = torch.sum(lad, dim=1)                                 # flows.py:238
    out = torch.cat([lower.reshape(-1, size, len(m)),
                     new_upper.reshape(-1, size, len(u))], dim=2).flatten(start_dim=1)
    return out, log_det, bins


def fcnn(x: torch.Tensor, net: Dict[str, torch.Tensor], prefix: str = "") -> torch.Tensor:
    """Linear-Tanh-Linear-Tanh-Linear with state-dict keys network.{0,2,4}.* (flows.py:26-35)."""
    h = torch.tanh(F.linear(x, net[prefix + "network.0.weight"], net[prefix + "network.0.bias"]))
    h = torch.tanh(F.linear(h, net[prefix + "network.2.weight"], net[prefix + "network.2.bias"]))
    return F.linear(h, net[prefix + "network.4.weight"], net[prefix + "network.4.bias"])


def nsf_cl(x, net, size, dim, mask, K, B, inverse, prefix="psi."):
    """Whole NSF_CL layer.  Returns (out, log_det, bins, params)."""
    lower, _ = nsf_cl_split(x, size, dim, mask)
    n_t = size * (dim - len(mask))
    params = fcnn(lower, net, prefix).reshape(-1, n_t, 3 * K - 1)   # flows.py:231
    out, ld, bins = nsf_cl_transform(x, params, size, dim, mask, K, B, inverse)
    return out, ld, bins, params


# --------------------------------------------------------------------------------------
# NSF_AR autoregressive spline flow
# --------------------------------------------------------------------------------------
def _nsf_ar_spline(v, raw, K, B, inverse):
    """flows.py:182-188 / :204-206: split, softmax x 2B, softplus, unconstrained_RQS."""
    W, H, D = torch.split(raw, K, dim=1)
    W, H = torch.softmax(W, dim=1), torch.softmax(H, dim=1)
    W, H = 2 * B * W, 2 * B * H
    D = F.softplus(D)
    out, lad, bins = rqs_elementwise(v, W, H, D, inverse, B)
    return out, lad, bins


def nsf_ar_trig(x: torch.Tensor, B: float) -> torch.Tensor:
    """trig_transform (flows.py:172-173): cat(cos(pi x / B), sin(pi x / B)) with pi as an fp32 tensor."""
    pi = torch.tensor(math.pi, dtype=torch.float32).to(x.dtype)
    return torch.cat((torch.cos(pi * x / B), torch.sin(pi * x / B)), dim=-1)


def nsf_ar(x: torch.Tensor, net: Dict[str, torch.Tensor], dim: int, K: int, B: float, inverse: bool,
           prefix: str = ""):
    """Whole NSF_AR layer (flows.py:175-208).  Returns (out, log_det, bins [N, dim])."""
    N = x.shape[0]
    out = torch.zeros_like(x)
    log_det = torch.zeros(N, dtype=x.dtype)
    bins = []
    for i in range(dim):
        src = out if inverse else x                                    # conditioner sees x[:, :i] either way
        if i == 0:
            raw = net[prefix + "init_param"].to(x.dtype).expand(N, 3 * K - 1)
        else:
            raw = fcnn(nsf_ar_trig(src[:, :i], B), net, prefix + f"layers.{i - 1}.")
        col, ld, b = _nsf_ar_spline(x[:, i], raw, K, B, inverse)
        out = out.clone()
        out[:, i] = col
        log_det = log_det + ld
        bins.append(b)
    return out, log_det, torch.stack(bins, dim=1)


# --------------------------------------------------------------------------------------
# priors / targets next to the flow (applications/src/systems.py)
# --------------------------------------------------------------------------------------
def einstein_logprob(x: torch.Tensor, centers: torch.Tensor, dim: int, alpha: float, boxlength=None):
    """EinsteinCrystal.log_prob (systems.py:360-366) with noise = MultivariateNormal(0, I/alpha)."""
    from torch.distributions import MultivariateNormal
    natoms = centers.shape[0]
    noise = MultivariateNormal(torch.zeros(dim, dtype=x.dtype), 1 / alpha * torch.eye(dim, dtype=x.dtype))
    dev = x.reshape(-1, natoms, dim) - centers
    if boxlength is not None:
        dev = dev - (torch.abs(dev) > 0.5 * boxlength) * torch.sign(dev) * boxlength
    return torch.sum(noise.log_prob(dev.reshape(-1, dim)).reshape(-1, natoms), dim=1)


def lj_potential(pos: torch.Tensor, boxlength: float, epsilon=1.0, sigma=1.0, cutoff=None, shift=True):
    """LJ.potential (systems.py:154-189) on pos [..., n, dim]."""
    pair_dist = pos.unsqueeze(-2) - pos.unsqueeze(-3)
    pair_dist = pair_dist - (torch.abs(pair_dist) > 0.5 * boxlength) * torch.sign(pair_dist) * boxlength
    distances = torch.linalg.norm(pair_dist.float(), axis=-1)
    scaled = distances + (distances == 0)
    inv = 1 / scaled
    if cutoff is not None:
        inv = inv - (distances > cutoff) * inv
        pow_6 = torch.pow(sigma * inv, 6)
        if shift:
            s = (sigma / cutoff) ** 6
            pair = epsilon * 4 * (torch.pow(pow_6, 2) - pow_6 - s ** 2 + s)
        else:
            pair = epsilon * 4 * (torch.pow(pow_6, 2) - pow_6)
    else:
        pow_6 = torch.pow(sigma * inv, 6)
        pair = epsilon * 4 * (torch.pow(pow_6, 2) - pow_6)
    pair = pair * inv * distances
    return torch.sum(pair, axis=(-1, -2)) / 2


def gmm_logprob(x: torch.Tensor, centers: torch.Tensor, vars_: torch.Tensor, npoints: int, dim: int):
    """GaussianMixture.log_prob (systems.py:287-292): plain exp-sum over the components."""
    from torch.distributions import MultivariateNormal
    pts = x.reshape(-1, dim)
    prob = 0
    nc = centers.shape[0]
    for i in range(nc):
        d = MultivariateNormal(centers[i], vars_[i] * torch.eye(dim, dtype=x.dtype))
        prob = prob + 1 / nc * torch.exp(d.log_prob(pts))
    return torch.sum(torch.log(prob).reshape(-1, npoints), axis=1)


# --------------------------------------------------------------------------------------
# BAR free-energy estimator (applications/src/bar.py)
# --------------------------------------------------------------------------------------
def _logsum(a):
    mx = torch.max(a)
    return torch.log(torch.sum(torch.exp(a - mx))) + mx                 # bar.py:3-14


def bar_zero(w_F: torch.Tensor, w_R: torch.Tensor, DeltaF: float) -> float:
    """BARzero (bar.py:16-59) in the dtype of the inputs."""
    T_F, T_R = float(w_F.numel()), float(w_R.numel())
    M = math.log(T_F / T_R)
    arg_F = M + w_F - DeltaF
    max_F = torch.where(arg_F < 0.0, torch.zeros_like(arg_F), arg_F)
    log_f_F = -max_F - torch.log(torch.exp(-max_F) + torch.exp(arg_F - max_F))
    log_numer = _logsum(log_f_F) - math.log(T_F)
    arg_R = M - w_R - DeltaF
    max_R = torch.where(arg_R < 0.0, torch.zeros_like(arg_R), arg_R)
    log_f_R = -max_R - torch.log(torch.exp(-max_R) + torch.exp(arg_R - max_R)) - w_R
    log_denom = _logsum(log_f_R) - math.log(T_R)
    return float(DeltaF - (log_denom - log_numer))


def bar(w_F: torch.Tensor, w_R: torch.Tensor, DeltaF: float = 0.0, maximum_iterations: int = 1000,
        relative_tolerance: float = 1.0e-5) -> float:
    """BAR (bar.py:61-67)."""
    for iteration in range(maximum_iterations):
        old = DeltaF
        DeltaF = -bar_zero(w_F, w_R, DeltaF) + DeltaF
        if iteration > 0 and abs((DeltaF - old) / DeltaF) < relative_tolerance:
            break
    return DeltaF


# --------------------------------------------------------------------------------------
# RealNVP, Planar, Radial
# --------------------------------------------------------------------------------------
def realnvp(x: torch.Tensor, net: Dict[str, torch.Tensor], inverse: bool, prefix: str = ""):
    """Two affine half-couplings (flows.py:52-63 forward, :65-76 inverse)."""
    half = x.shape[1] // 2
    lower, upper = x[:, :half], x[:, half:]
    if not inverse:
        t1 = fcnn(lower, net, prefix + "t1.")
        s1 = fcnn(lower, net, prefix + "s1.")
        upper = t1 + upper * torch.exp(s1)
        t2 = fcnn(upper, net, prefix + "t2.")
        s2 = fcnn(upper, net, prefix + "s2.")
        lower = t2 + lower * torch.exp(s2)
        ld = torch.sum(s1, dim=1) + torch.sum(s2, dim=1)
    else:
        t2 = fcnn(upper, net, prefix + "t2.")
        s2 = fcnn(upper, net, prefix + "s2.")
        lower = (lower - t2) * torch.exp(-s2)
        t1 = fcnn(lower, net, prefix + "t1.")
        s1 = fcnn(lower, net, prefix + "s1.")
        upper = (upper - t1) * torch.exp(-s1)
        ld = torch.sum(-s1, dim=1) + torch.sum(-s2, dim=1)
    return torch.cat([lower, upper], dim=1), ld


def planar(x: torch.Tensor, w: torch.Tensor, u: torch.Tensor, b: torch.Tensor):
    """tanh planar layer with the reference's u-hat and +1e-4 (flows_1.py:49-60, Q8)."""
    wu = w @ u
    scal = torch.log(1 + torch.exp(wu)) - wu - 1
    uhat = u + scal * w / torch.norm(w) ** 2
    lin = torch.unsqueeze(x @ w, 1) + b
    th = torch.tanh(lin)
    z = x + uhat * th
    phi = (1 - torch.pow(th, 2)) * w
    return z, torch.log(torch.abs(1 + phi @ uhat) + 1e-4)


def radial(x: torch.Tensor, x0: torch.Tensor, log_alpha: torch.Tensor, beta_raw: torch.Tensor,
           per_sample: bool = False):
    """Radial layer (flows_1.py:85-97).

    ``per_sample=False`` reproduces the reference (Q9): r is ONE Frobenius norm over the
    whole batch and log_det has shape [1].  ``per_sample=True`` is the corrected
    per-row-norm variant, log_det shape [N].
    """
    n = x.shape[1]
    diff = x - x0
    r = torch.norm(diff, dim=1, keepdim=True) if per_sample else torch.norm(diff)
    alpha = torch.exp(log_alpha)
    h = 1 / (alpha + r)
    beta = -alpha + torch.log(1 + torch.exp(beta_raw))
    z = x + beta * h * diff
    ld = (n - 1) * torch.log(1 + beta * h) + torch.log(1 + beta * h - beta * r / (alpha + r) ** 2)
    return z, (ld[:, 0] if per_sample else ld)


# --------------------------------------------------------------------------------------
# model level
# --------------------------------------------------------------------------------------
def gauss_logprob(z: torch.Tensor, var: float = 1.0) -> torch.Tensor:
    """log N(z; 0, var*I) — the prior the hot configs use (applications/src/setup.py:25-30)."""
    d = z.shape[1]
    return -0.5 * torch.sum(z * z, dim=1) / var - 0.5 * d * math.log(2 * math.pi * var)


def apply_layer(spec: dict, sd: Dict[str, torch.Tensor], i: int, x: torch.Tensor, inverse: bool):
    """One layer of a flow described by ``spec`` (see ``flow_forward``)."""
    kind = spec["type"]
    pre = f"flows.{i}."
    if kind == "NSF_CL":
        out, ld, _, _ = nsf_cl(x, sd, spec["size"], spec["dim"], spec["mask"], spec["K"],
                               spec["B"], inverse, prefix=pre + "psi.")
        return out, ld
    if kind == "NSF_AR":
        out, ld, _ = nsf_ar(x, sd, spec["dim"], spec["K"], spec["B"], inverse, prefix=pre)
        return out, ld
    if kind == "RealNVP":
        return realnvp(x, sd, inverse, prefix=pre)
    if kind == "Planar":
        if inverse:
            raise NotImplementedError("Planar flow has no algebraic inverse.")  # flows_1.py:62
        return planar(x, sd[pre + "w"], sd[pre + "u"], sd[pre + "b"])
    if kind == "Radial":
        if inverse:
            raise NotImplementedError("Radial flow has no inverse in the reference.")
        return radial(x, sd[pre + "x0"], sd[pre + "log_alpha"], sd[pre + "beta"],
                      per_sample=spec.get("per_sample", False))
    raise ValueError(kind)


def flow_forward(specs: List[dict], sd: Dict[str, torch.Tensor], x: torch.Tensor, var: float = 1.0):
    """NormalizingFlowModel.forward (models.py:13-20): (z, prior_logprob, log_det)."""
    log_det = torch.zeros(x.shape[0], dtype=x.dtype)
    for i, spec in enumerate(specs):
        x, ld = apply_layer(spec, sd, i, x, inverse=False)
        log_det = log_det + ld
    return x, gauss_logprob(x, var), log_det


def flow_inverse(specs: List[dict], sd: Dict[str, torch.Tensor], z: torch.Tensor):
    """NormalizingFlowModel.inverse (models.py:22-29): (x, log_det)."""
    log_det = torch.zeros(z.shape[0], dtype=z.dtype)
    for i in reversed(range(len(specs))):
        z, ld = apply_layer(specs[i], sd, i, z, inverse=True)
        log_det = log_det + ld
    return z, log_det


def flow_fwd_inv_pass(specs, sd, x, z):
    """One unit of the BASELINE metric: one forward and one inverse pass, each with log-det."""
    a = flow_forward(specs, sd, x)
    b = flow_inverse(specs, sd, z)
    return a, b


# --------------------------------------------------------------------------------------
# HMC acceptance logic (hmc.py:43-65) on a batch of independent chains
# --------------------------------------------------------------------------------------
def hmc_accept(u_old: torch.Tensor, u_new: torch.Tensor, beta: float, uniform: torch.Tensor):
    """Metropolis test on the potential only (Q12): accept iff uniform < exp((U_old-U_new)*beta)."""
    return uniform < torch.exp((u_old - u_new) * beta)


def leapfrog(q, p, force_fn, n_steps: int, dt: float, inv_mass: float = 1.0):
    """Velocity-Verlet with the force re-evaluated at the NEW position (intentional fix of Q13).

    ``force_fn(q) -> (U, F)``.  Returns (q, p, U(q_final)).
    """
    U, Fq = force_fn(q)
    for _ in range(n_steps):
        p = p + 0.5 * dt * Fq
        q = q + dt * inv_mass * p
        U, Fq = force_fn(q)
        p = p + 0.5 * dt * Fq
    return q, p, U
