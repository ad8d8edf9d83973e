"""Flow layers with the reference's class API (nf/flows.py, nf/flows_1.py:21-97) on libnfk.

Constructor signatures, attribute names and state-dict keys follow the reference so that its
checkpoints load with ``load_state_dict``:
  FCNN      network.{0,2,4}.{weight,bias}          nf/flows.py:20-35
  RealNVP   {t1,s1,t2,s2}.network.*                nf/flows.py:38-76
  NSF_AR    init_param, layers.{i}.network.*       nf/flows.py:152-209
  NSF_CL    psi.network.*                          nf/flows.py:210-253
  Planar    w, u, b                                nf/flows_1.py:21-63
  Radial    x0, log_alpha, beta                    nf/flows_1.py:66-97
Every ``forward``/``inverse`` returns ``(z, log_det)``.  Tensors must be CUDA tensors.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.init as init

from . import _ops

__all__ = ["FCNN", "RealNVP", "NSF_AR", "NSF_CL", "Planar", "Radial", "PlanarStack", "RadialStack"]


class FCNN(nn.Module):
    """Linear-Tanh-Linear-Tanh-Linear conditioner (nf/flows.py:20-35).

    ``precision``: "fp32" = fp32-class GEMMs (3xTF32 on the tensor cores for inference, CUDA-core
    kernels under autograd; parity mode, matches the reference's addmm to fp32 round-off);
    "bf16" = tcgen05 tensor-core GEMMs with 16-bit operands and fp32 accumulation (throughput mode,
    1e-2 parity class: fp16 operands on the inference forward, bf16 on the gradient paths);
    "fp32x3" = as "fp32", and eligible NSF_CL layers run the split-operand fused layer kernel
    (fp16 hi + lo operands, three tensor-core MMAs per product: fp32-class accuracy in one launch).
    """

    def __init__(self, in_dim, out_dim, hidden_dim, precision="fp32"):
        super().__init__()
        self.network = nn.Sequential(
            nn.Linear(in_dim, hidden_dim),
            nn.Tanh(),
            nn.Linear(hidden_dim, hidden_dim),
            nn.Tanh(),
            nn.Linear(hidden_dim, out_dim),
        )
        self.precision = precision
        self._bf16_cache = None

    def forward(self, x):
        l0, l2, l4 = self.network[0], self.network[2], self.network[4]
        if self.precision == "bf16":
            from . import _bf16
            return _bf16.mlp3(self, x)
        if self.precision not in ("fp32", "fp32x3"):
            raise ValueError(f"unknown conditioner precision {self.precision!r}")
        if not torch.is_grad_enabled():
            # inference: the three GEMMs on the tensor cores with fp32-class accuracy (3xTF32,
            # csrc/linear_tf32.cu) wherever the shape allows; same 1e-5 parity class as the CUDA-core kernel
            h = _ops.linear_f32(x, l0.weight, l0.bias, 1)
            h = _ops.linear_f32(h, l2.weight, l2.bias, 1)
            return _ops.linear_f32(h, l4.weight, l4.bias, 0)
        h = _ops.LinearF32Fn.apply(x, l0.weight, l0.bias, 1)
        h = _ops.LinearF32Fn.apply(h, l2.weight, l2.bias, 1)
        return _ops.LinearF32Fn.apply(h, l4.weight, l4.bias, 0)


class RealNVP(nn.Module):
    """Two affine half-couplings (nf/flows.py:38-76); only even ``dim`` works, as in the
    reference (quirk Q10)."""

    def __init__(self, dim, hidden_dim=800, base_network=FCNN):
        super().__init__()
        self.dim = dim
        self.t1 = base_network(dim // 2, dim // 2, hidden_dim)
        self.s1 = base_network(dim // 2, dim // 2, hidden_dim)
        self.t2 = base_network(dim // 2, dim // 2, hidden_dim)
        self.s2 = base_network(dim // 2, dim // 2, hidden_dim)

    def forward(self, x):
        h = self.dim // 2
        lower, upper = x[:, :h], x[:, h:]
        t1, s1 = self.t1(lower), self.s1(lower)
        upper, ld1 = _ops.AffineHalfFn.apply(upper, s1, t1, False)       # flows.py:56
        t2, s2 = self.t2(upper), self.s2(upper)
        lower, ld2 = _ops.AffineHalfFn.apply(lower, s2, t2, False)       # flows.py:59
        return torch.cat([lower, upper], dim=1), ld1 + ld2               # flows.py:60-62

    def inverse(self, z):
        h = self.dim // 2
        lower, upper = z[:, :h], z[:, h:]
        t2, s2 = self.t2(upper), self.s2(upper)
        lower, ld2 = _ops.AffineHalfFn.apply(lower, s2, t2, True)        # flows.py:69
        t1, s1 = self.t1(lower), self.s1(lower)
        upper, ld1 = _ops.AffineHalfFn.apply(upper, s1, t1, True)        # flows.py:72
        return torch.cat([lower, upper], dim=1), ld1 + ld2               # flows.py:73-75


class NSF_AR(nn.Module):
    """Autoregressive neural spline flow (nf/flows.py:152-209): dimension i is transformed by an RQS
    whose 3K-1 parameters come from ``layers[i-1]`` applied to [cos, sin](pi x[:, :i] / B)
    (``trig_transform``, flows.py:172-173); dimension 0 uses the learned ``init_param``.

    forward (x -> z) needs only known inputs, so all dim conditioners run back to back and ONE
    element-wise spline launch transforms every column; inverse (z -> x) is inherently sequential
    (flows.py:196-208): dim conditioner + spline steps.
    """

    def __init__(self, dim, K=32, B=3, hidden_dim=800, base_network=FCNN, device="cpu", arith=_ops.DEFAULT_ARITH):
        super().__init__()
        self.dim = dim
        self.K = K
        self.B = B
        self.device = device
        self.arith = arith
        self.layers = nn.ModuleList()
        self.init_param = nn.Parameter(torch.Tensor(3 * K - 1))
        for i in range(1, dim):
            self.layers += [base_network(2 * i, 3 * K - 1, hidden_dim)]
        self.reset_parameters()

    def reset_parameters(self):
        init.uniform_(self.init_param, -1 / 2, 1 / 2)

    def trig_transform(self, x):
        pi = torch.tensor(math.pi, dtype=torch.float32, device=x.device)
        return torch.cat((torch.cos(pi * x / self.B), torch.sin(pi * x / self.B)), dim=-1)

    def _spline(self, v, params, inverse):
        if torch.is_grad_enabled() and (v.requires_grad or params.requires_grad):
            return _ops.RqsElementwiseFn.apply(v, params, self.K, float(self.B), inverse, self.arith)
        out, lad, _ = _ops.rqs_elementwise(v, params, self.K, float(self.B), inverse, self.arith)
        return out, lad

    def forward(self, x):
        N = x.shape[0]
        P = 3 * self.K - 1
        if not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))) \
                and x.dtype == torch.float32 and x.is_cuda:
            from . import _wide
            if _wide.nsf_ar_grouped_ok(self):
                # every conditioner in three grouped tensor-core launches + one spline launch
                z, lad = self._spline(x, _wide.nsf_ar_params(self, x), False)
                return z, lad.sum(dim=1)
        pi = torch.tensor(math.pi, dtype=torch.float32, device=x.device)
        ang = pi * x / self.B
        c, s = torch.cos(ang), torch.sin(ang)
        cols = [self.init_param.to(x.device).expand(N, P)]
        for i in range(1, self.dim):
            cols.append(self.layers[i - 1](torch.cat((c[:, :i], s[:, :i]), dim=-1)))      # flows.py:186
        params = torch.stack(cols, dim=1)                                                # [N, dim, 3K-1]
        z, lad = self._spline(x, params, False)
        return z, lad.sum(dim=1)

    def inverse(self, z):
        N = z.shape[0]
        P = 3 * self.K - 1
        xs = []
        log_det = torch.zeros(N, dtype=torch.float32, device=z.device)
        for i in range(self.dim):
            if i == 0:
                params = self.init_param.to(z.device).expand(N, P)
            else:
                params = self.layers[i - 1](self.trig_transform(torch.stack(xs, dim=1)))  # flows.py:203
            xi, ld = self._spline(z[:, i].contiguous(), params.contiguous(), True)
            xs.append(xi)
            log_det = log_det + ld
        return torch.stack(xs, dim=1), log_det


class NSF_CL(nn.Module):
    """Neural spline flow coupling layer (nf/flows.py:210-253).

    ``x`` is viewed as [N, size, dim]; columns ``mask`` condition, the other columns go through
    an unconstrained RQS whose 3K-1 parameters per feature come from ``psi``.  The output keeps
    the reference's column order: conditioning columns first inside each dim-group (quirk Q5).
    ``arith``: "hybrid" (default; bins bit-exact), "exact" or "fast" — see include/nfk.h.
    """

    def __init__(self, size, dim=3, K=32, B=3, hidden_dim=800, base_network=FCNN, device="cpu", mask=[1],
                 arith=_ops.DEFAULT_ARITH):
        super().__init__()
        self.size = size
        self.dim = dim
        self.K = K
        self.B = B
        self.device = device
        self.mask = torch.Tensor(mask).long()                            # plain attributes, not buffers (Q14)
        self.unmasked = torch.Tensor([x for x in range(self.dim) if x not in self.mask]).long()
        self._mask = [int(m) for m in mask]
        self._unmasked = [c for c in range(dim) if c not in self._mask]
        self.arith = arith
        self.fused = True          # use the fused layer kernel when the layer is eligible (see _fused.py)
        self.psi = base_network(len(mask) * self.size,
                                (3 * K - 1) * (self.dim - len(self.mask)) * self.size, hidden_dim).to(self.device)

    # conditioner input x[:, :, mask].flatten(1) (flows.py:230)
    def _lower(self, x):
        if x.requires_grad and torch.is_grad_enabled():
            # device-resident index (a Python list would be uploaded on every call, which also breaks
            # CUDA-graph capture of a training step)
            idx = getattr(self, "_mask_dev", None)
            if idx is None or idx.device != x.device:
                idx = self._mask_dev = torch.tensor(self._mask, dtype=torch.long, device=x.device)
            return x.reshape(-1, self.size, self.dim).index_select(2, idx).flatten(start_dim=1)
        if getattr(self.psi, "precision", None) == "bf16":
            # gather straight into the padded bf16 operand of the first tensor-core GEMM
            w = self.size * len(self._mask)
            return _ops.gather_cols(x, self.size, self.dim, self._mask, bf16=True, ld_out=(w + 7) // 8 * 8)
        return _ops.gather_cols(x, self.size, self.dim, self._mask)

    def _transform(self, x, inverse, logdet=None):
        no_grad = not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())))
        if no_grad and self.fused:
            from . import _fused
            if _fused.eligible(self):
                # conditioner GEMMs + spline in ONE kernel; the parameter tensor never reaches HBM
                return _fused.run(self, x, inverse, logdet)
        n_t = self.size * (self.dim - len(self._mask))
        if no_grad and x.dtype == torch.float32:
            from . import _wide
            if self.fused and _wide.rqs_eligible(self):
                # wide conditioner with the spline transform as the epilogue of its last GEMM
                return _wide.run_layer(self, x, inverse, logdet)
            if _wide.usable(self.psi):
                # wide conditioner: gather + 3 persistent tcgen05 GEMMs over image-layout operands
                params = _wide.mlp3(self.psi, x, self.size, self.dim, self._mask).reshape(-1, n_t, 3 * self.K - 1)
                out, ld, _ = _ops.rqs_coupling(x, params, self.size, self.dim, self._mask, self.K, float(self.B),
                                               inverse, self.arith, logdet=logdet)
                return out, ld
        if not no_grad and self.fused and x.dtype == torch.float32:
            from . import _wide
            if _wide.grad_eligible(self):
                # training / gradient path on the tensor cores (forward keeps the activation images)
                net = self.psi.network
                out, ld = _wide.NsfWideFn.apply(x, net[0].weight, net[0].bias, net[2].weight, net[2].bias,
                                                net[4].weight, net[4].bias, self, bool(inverse))
                if logdet is not None:
                    ld = logdet + ld
                return out, ld
        params = self.psi(self._lower(x)).reshape(-1, n_t, 3 * self.K - 1)   # flows.py:231
        if torch.is_grad_enabled() and (x.requires_grad or params.requires_grad):
            out, ld = _ops.RqsCouplingFn.apply(x, params, self.size, self.dim, tuple(self._mask), self.K,
                                               float(self.B), inverse, self.arith)
            if logdet is not None:
                ld = logdet + ld
            return out, ld
        out, ld, _ = _ops.rqs_coupling(x, params, self.size, self.dim, self._mask, self.K, float(self.B),
                                       inverse, self.arith, logdet=logdet)
        return out, ld

    def forward(self, x):
        return self._transform(x, False)

    def inverse(self, z):
        return self._transform(z, True)

    # used by NormalizingFlowModel: accumulate the layer's log-det into ``logdet`` in-kernel
    def _nfk_step(self, x, logdet, inverse):
        return self._transform(x, inverse, logdet)


class Planar(nn.Module):
    """Planar flow z = x + uhat * tanh(w.x + b) (nf/flows_1.py:21-63, quirk Q8).  Only tanh has a
    working derivative in the reference, and only tanh is implemented here."""

    def __init__(self, dim, nonlinearity=torch.tanh):
        super().__init__()
        if nonlinearity is not torch.tanh:
            raise NotImplementedError("only torch.tanh is supported (the reference's other "
                                      "derivatives are CPU-only and sign-buggy, flows_1.py:12-18)")
        self.h = nonlinearity
        self.w = nn.Parameter(torch.Tensor(dim))
        self.u = nn.Parameter(torch.Tensor(dim))
        self.b = nn.Parameter(torch.Tensor(1))
        self.reset_parameters(dim)

    def reset_parameters(self, dim):
        init.uniform_(self.w, -math.sqrt(1 / dim), math.sqrt(1 / dim))
        init.uniform_(self.u, -math.sqrt(1 / dim), math.sqrt(1 / dim))
        init.uniform_(self.b, -math.sqrt(1 / dim), math.sqrt(1 / dim))

    def forward(self, x):
        return _ops.PlanarStackFn.apply(x, self.w[None, :], self.u[None, :], self.b.reshape(1))

    def inverse(self, z):
        raise NotImplementedError("Planar flow has no algebraic inverse.")  # flows_1.py:62-63

    def _nfk_planar(self):
        return self.w, self.u, self.b


class PlanarStack(nn.Module):
    """Fuses a run of consecutive ``Planar`` layers into one kernel launch (one pass over HBM
    instead of L).  Built by NormalizingFlowModel; parameters stay owned by the layers."""

    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)

    def forward(self, x):
        w = torch.stack([l.w for l in self.layers])
        u = torch.stack([l.u for l in self.layers])
        b = torch.cat([l.b.reshape(1) for l in self.layers])
        return _ops.PlanarStackFn.apply(x, w, u, b)


class Radial(nn.Module):
    """Radial flow z = x + beta*h*(x - x0) (nf/flows_1.py:66-97).

    ``per_sample=False`` (default) reproduces the reference: r is ONE Frobenius norm over the
    whole batch and log_det has shape [1] (quirk Q9; across ranks the sum of squares is
    all-reduced).  ``per_sample=True`` uses the per-row norm and returns log_det [N].
    The reference never initialises its parameters (its reset_parameters is broken);
    here they start at U(-1/sqrt(d), 1/sqrt(d))."""

    def __init__(self, dim, per_sample=False):
        super().__init__()
        self.x0 = nn.Parameter(torch.Tensor(dim))
        self.log_alpha = nn.Parameter(torch.Tensor(1))
        self.beta = nn.Parameter(torch.Tensor(1))
        self.per_sample = per_sample
        self.reset_parameters(dim)

    def reset_parameters(self, dim):
        b = math.sqrt(1 / dim)
        init.uniform_(self.x0, -b, b)
        init.uniform_(self.log_alpha, -b, b)
        init.uniform_(self.beta, -b, b)

    def forward(self, x):
        return _ops.RadialFn.apply(x, self.x0, self.log_alpha, self.beta, self.per_sample)

    def inverse(self, z):
        raise NotImplementedError("Radial flow has no inverse in the reference.")


class RadialStack(nn.Module):
    """Fuses a run of consecutive ``Radial`` layers of the same mode (built by NormalizingFlowModel; parameters
    stay owned by the layers).  Without autograd: per-sample runs are ONE pass over the batch, batch-global
    runs (the reference's behaviour) one read + one write per layer; with autograd each layer runs through its
    own ``RadialFn``."""

    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        self.per_sample = bool(layers[0].per_sample)

    def forward(self, x):
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        d = x.shape[1]
        if needs_grad or not _ops.radial_stack_ok(d, self.per_sample) or x.dtype != torch.float32:
            ld = None
            for l in self.layers:
                x, l_ld = l.forward(x)
                ld = l_ld if ld is None else ld + l_ld
            return x, ld
        x0 = torch.stack([l.x0.detach() for l in self.layers])
        la = torch.cat([l.log_alpha.detach().reshape(1) for l in self.layers])
        be = torch.cat([l.beta.detach().reshape(1) for l in self.layers])
        return _ops.radial_stack(x, x0, la, be, self.per_sample)
