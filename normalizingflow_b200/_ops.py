"""Thin operator layer over libnfk: raw launchers plus the torch.autograd.Functions the layer
classes use.  Every launcher allocates its outputs with torch (the library never allocates),
passes raw device pointers + the current CUDA stream, and turns error codes into exceptions.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import call, f32c, i32_array, ptr, require_cuda, stream_ptr

DEFAULT_ARITH = "hybrid"


class KernelTimer:
    """CUDA-event timer for individual launches (bench.py's live roofline measurement):
    events are recorded on the launching stream around each wrapped launch."""

    def __init__(self):
        self.events = {}

    def start(self, name, dev):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(dev))
        self.events.setdefault(name, []).append((e0, e1))
        return e1

    def stop(self, e1, dev):
        e1.record(torch.cuda.current_stream(dev))

    def summary(self):
        """name -> (launches, total_ms); call after a synchronize."""
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in self.events.items()}


KERNEL_TIMER = None


def _arith(a) -> int:
    if isinstance(a, int):
        return a
    try:
        return _lib.ARITH[a]
    except KeyError:
        raise ValueError(f"arith must be one of {sorted(_lib.ARITH)}, got {a!r}") from None


# --------------------------------------------------------------------------------------
# RQS coupling (nf/flows.py:232-239 + nf/utils.py:20-152)
# --------------------------------------------------------------------------------------
def rqs_coupling(x: torch.Tensor, params: torch.Tensor, size: int, dim: int, mask: Sequence[int],
                 K: int, B: float, inverse: bool, arith=DEFAULT_ARITH,
                 logdet: Optional[torch.Tensor] = None, want_bins: bool = False):
    """Raw launch (no autograd).  ``params`` [N, F_t, 3K-1] raw conditioner output.
    Returns (out [N, size*dim], logdet [N], bins [N, F_t] int8 | None).  When ``logdet`` is
    given it is accumulated into in place."""
    dev = require_cuda(x, params, logdet)
    N = x.shape[0]
    d = size * dim
    n_t = size * (dim - len(mask))
    P = 3 * K - 1
    if x.numel() != N * d:
        raise ValueError(f"x has {x.numel()} elements, expected N*size*dim = {N}*{d}")
    if params.numel() != N * n_t * P:
        raise ValueError(f"params has {params.numel()} elements, expected {N}*{n_t}*{P}")
    x = f32c(x)
    params = f32c(params)
    out = torch.empty((N, d), dtype=torch.float32, device=dev)
    accumulate = logdet is not None
    if accumulate:
        if logdet.dtype != torch.float32 or not logdet.is_contiguous() or logdet.numel() != N:
            raise ValueError("logdet accumulator must be a contiguous fp32 [N] tensor")
    else:
        logdet = torch.empty((N,), dtype=torch.float32, device=dev)
    bins = torch.empty((N, n_t), dtype=torch.int8, device=dev) if want_bins else None
    m = i32_array(mask)
    with torch.cuda.device(dev):
        tm = KERNEL_TIMER
        ev = tm.start("rqs_coupling", dev) if tm is not None else None
        call("nfk_rqs_coupling", ptr(x), ptr(params), ptr(out), ptr(logdet), ptr(bins), N, size, dim, m,
             len(mask), K, float(B), int(bool(inverse)), int(accumulate), _arith(arith), stream_ptr(dev))
        if ev is not None:
            tm.stop(ev, dev)
    return out, logdet, bins


def rqs_coupling_bwd(x, params, grad_out, grad_logdet, size, dim, mask, K, B, inverse):
    dev = require_cuda(x, params, grad_out, grad_logdet)
    N = x.shape[0]
    x, params = f32c(x), f32c(params)
    grad_out = f32c(grad_out) if grad_out is not None else torch.zeros_like(x)
    grad_logdet = f32c(grad_logdet) if grad_logdet is not None else None
    gx = torch.empty_like(x)
    gp = torch.empty_like(params)
    m = i32_array(mask)
    with torch.cuda.device(dev):
        call("nfk_rqs_coupling_bwd", ptr(x), ptr(params), ptr(grad_out), ptr(grad_logdet), ptr(gx), ptr(gp),
             N, size, dim, m, len(mask), K, float(B), int(bool(inverse)), stream_ptr(dev))
    return gx, gp


class RqsCouplingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, params, size, dim, mask, K, B, inverse, arith):
        out, ld, _ = rqs_coupling(x, params, size, dim, mask, K, B, inverse, arith)
        ctx.save_for_backward(x, params)
        ctx.cfg = (size, dim, tuple(mask), K, B, inverse)
        return out, ld

    @staticmethod
    def backward(ctx, g_out, g_ld):
        x, params = ctx.saved_tensors
        size, dim, mask, K, B, inverse = ctx.cfg
        gx, gp = rqs_coupling_bwd(x, params, g_out, g_ld, size, dim, mask, K, B, inverse)
        return gx.view_as(x), gp.view_as(params), None, None, None, None, None, None, None


def rqs_elementwise(inputs, params, K: int, B: float, inverse: bool, arith=DEFAULT_ARITH, want_bins=False):
    """Element-wise spline on raw conditioner outputs (NSF_AR, nf/flows.py:178-190 / :196-208):
    inputs [...], params [..., 3K-1] -> (out [...], lad [...], bins | None)."""
    dev = require_cuda(inputs, params)
    P = 3 * K - 1
    M = inputs.numel()
    if params.numel() != M * P:
        raise ValueError(f"params has {params.numel()} elements, expected {M}*{P}")
    shape = inputs.shape
    inputs, params = f32c(inputs), f32c(params)
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    lad = torch.empty(shape, dtype=torch.float32, device=dev)
    bins = torch.empty(shape, dtype=torch.int8, device=dev) if want_bins else None
    with torch.cuda.device(dev):
        call("nfk_rqs_elementwise", ptr(inputs), ptr(params), ptr(out), ptr(lad), ptr(bins), M, K, float(B),
             int(bool(inverse)), _arith(arith), stream_ptr(dev))
    return out, lad, bins


class RqsElementwiseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, params, K, B, inverse, arith):
        out, lad, _ = rqs_elementwise(inputs, params, K, B, inverse, arith)
        ctx.save_for_backward(inputs, params)
        ctx.cfg = (K, B, inverse)
        return out, lad

    @staticmethod
    def backward(ctx, g_out, g_lad):
        inputs, params = ctx.saved_tensors
        K, B, inverse = ctx.cfg
        dev = inputs.device
        xi, pp = f32c(inputs), f32c(params)
        g_out = f32c(g_out) if g_out is not None else torch.zeros_like(xi)
        g_lad = f32c(g_lad) if g_lad is not None else None
        gi = torch.empty_like(xi)
        gp = torch.empty_like(pp)
        with torch.cuda.device(dev):
            call("nfk_rqs_elementwise_bwd", ptr(xi), ptr(pp), ptr(g_out), ptr(g_lad), ptr(gi), ptr(gp), xi.numel(), K,
                 float(B), int(bool(inverse)), stream_ptr(dev))
        return gi.view_as(inputs), gp.view_as(params), None, None, None, None


def unconstrained_rqs(inputs, W, H, D, inverse: bool, B: float, arith=DEFAULT_ARITH, want_bins=False):
    """nf/utils.py:27-56 on tensors inputs [...], W,H [...,K], D [...,K-1]."""
    dev = require_cuda(inputs, W, H, D)
    K = W.shape[-1]
    if H.shape[-1] != K or D.shape[-1] != K - 1:
        raise ValueError("W, H need K and D K-1 entries in the last dimension")
    shape = inputs.shape
    M = inputs.numel()
    if W.numel() != M * K or H.numel() != M * K or D.numel() != M * (K - 1):
        raise ValueError("parameter tensors do not match the shape of inputs")
    inputs, W, H, D = f32c(inputs), f32c(W), f32c(H), f32c(D)
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    lad = torch.empty(shape, dtype=torch.float32, device=dev)
    bins = torch.empty(shape, dtype=torch.int8, device=dev) if want_bins else None
    with torch.cuda.device(dev):
        call("nfk_unconstrained_rqs", ptr(inputs), ptr(W), ptr(H), ptr(D), ptr(out), ptr(lad), ptr(bins), M, K,
             float(B), int(bool(inverse)), _arith(arith), stream_ptr(dev))
    return out, lad, bins


def debug_knots(logits, B: float, layer_norm: bool, exact: bool):
    dev = require_cuda(logits)
    logits = f32c(logits)
    M, K = logits.shape
    out = torch.empty((M, K + 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("nfk_debug_knots", ptr(logits), ptr(out), M, K, float(B), int(layer_norm), int(exact), stream_ptr(dev))
    return out


# --------------------------------------------------------------------------------------
# conditioner layers (nf/flows.py:26-35)
# --------------------------------------------------------------------------------------
def gemm_f32(A, ta: bool, Bm, tb: bool, out=None, accumulate=False):
    """C[M,N] (+)= op(A) op(B) on row-major 2-D fp32 tensors (inner stride 1)."""
    dev = require_cuda(A, Bm, out)
    M, Kd = (A.shape[1], A.shape[0]) if ta else A.shape
    K2, N = (Bm.shape[1], Bm.shape[0]) if tb else Bm.shape
    if Kd != K2:
        raise ValueError(f"inner dimensions differ: {Kd} vs {K2}")
    for t in (A, Bm):
        if t.dtype != torch.float32 or t.stride(1) != 1:
            raise ValueError("gemm_f32 needs fp32 operands with unit inner stride")
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=dev)
        accumulate = False
    with torch.cuda.device(dev):
        call("nfk_gemm_f32", ptr(A), A.stride(0), int(ta), ptr(Bm), Bm.stride(0), int(tb), ptr(out),
             out.stride(0), M, N, Kd, int(accumulate), stream_ptr(dev))
    return out


TF32X3 = True             # parity-mode forward GEMMs on the tensor cores (3xTF32) when the shape allows


def linear_f32(x, weight, bias, act: int, allow_tc: bool = True):
    """act(x @ weight.T + bias); x may be a row-strided 2-D view (inner stride 1).  ``allow_tc=False`` keeps the
    CUDA-core fp32 kernel (training: the saved activations feed the hand-written backward)."""
    dev = require_cuda(x, weight, bias)
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        x = f32c(x.reshape(x.shape[0], -1))
    weight = f32c(weight)
    bias = f32c(bias) if bias is not None else None
    M, Kd = x.shape
    Nout = weight.shape[0]
    if weight.shape[1] != Kd:
        raise ValueError(f"weight is {tuple(weight.shape)}, input has {Kd} features")
    y = torch.empty((M, Nout), dtype=torch.float32, device=dev)
    ldx = x.stride(0) if M > 1 else max(Kd, x.stride(0))
    with torch.cuda.device(dev):
        if (TF32X3 and allow_tc and Kd % 4 == 0 and ldx % 4 == 0 and x.data_ptr() % 16 == 0
                and weight.data_ptr() % 16 == 0):
            # fp32-class accuracy on the tensor cores (3xTF32, csrc/linear_tf32.cu)
            call("nfk_linear_tf32x3", ptr(x), ldx, ptr(weight), weight.stride(0), ptr(bias), ptr(y), Nout, M, Kd,
                 Nout, act, stream_ptr(dev))
        else:
            call("nfk_linear_f32", ptr(x), ldx, ptr(weight), ptr(bias), ptr(y), M, Kd, Nout, act, stream_ptr(dev))
    return y


class LinearF32Fn(torch.autograd.Function):
    """One nn.Linear (+Tanh) of the conditioner on the fp32 CUDA-core kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias, act):
        y = linear_f32(x, weight, bias, act, allow_tc=False)      # the saved activations feed the hand-written backward
        ctx.save_for_backward(x, weight, y)
        ctx.act = act
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, y = ctx.saved_tensors
        gy = f32c(gy)
        if ctx.act == 1:
            gy = gy * (1 - y * y)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = gemm_f32(gy, False, weight, False)              # [M,N] @ [N,K]
        if ctx.needs_input_grad[1]:
            xx = x if (x.stride(1) == 1 and x.dtype == torch.float32) else f32c(x)
            gw = gemm_f32(gy, True, xx, False)                   # [N,M] @ [M,K]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(0)
        return gx, gw, gb, None


# --------------------------------------------------------------------------------------
# affine half-coupling (nf/flows.py:52-76)
# --------------------------------------------------------------------------------------
def affine_half(x, v_off, s, t, out, y_off, logdet, inverse, accumulate):
    dev = require_cuda(x, s, t, out, logdet)
    N, h = s.shape
    with torch.cuda.device(dev):
        call("nfk_affine_halfcoupling", ptr(x), x.stride(0), v_off, ptr(s), ptr(t), ptr(out), out.stride(0),
             y_off, ptr(logdet), N, h, int(bool(inverse)), int(bool(accumulate)), stream_ptr(dev))


class AffineHalfFn(torch.autograd.Function):
    """y = t + v*exp(s) (forward) or (v - t)*exp(-s) (inverse); returns (y, +-sum s)."""

    @staticmethod
    def forward(ctx, v, s, t, inverse):
        dev = require_cuda(v, s, t)
        v = v if (v.dtype == torch.float32 and v.stride(1) == 1) else f32c(v)
        s, t = f32c(s), f32c(t)
        N, h = s.shape
        y = torch.empty((N, h), dtype=torch.float32, device=dev)
        ld = torch.empty((N,), dtype=torch.float32, device=dev)
        affine_half(v, 0, s, t, y, 0, ld, inverse, False)
        ctx.save_for_backward(v, s, t)
        ctx.inverse = inverse
        return y, ld

    @staticmethod
    def backward(ctx, gy, gld):
        v, s, t = ctx.saved_tensors
        dev = v.device
        N, h = s.shape
        gy = f32c(gy) if gy is not None else torch.zeros((N, h), dtype=torch.float32, device=dev)
        gld = f32c(gld) if gld is not None else None
        gv = torch.empty((N, h), dtype=torch.float32, device=dev)
        gs = torch.empty_like(s)
        gt = torch.empty_like(t)
        with torch.cuda.device(dev):
            call("nfk_affine_halfcoupling_bwd", ptr(v), v.stride(0), 0, ptr(s), ptr(t), ptr(gy), gy.stride(0), 0,
                 ptr(gld), ptr(gv), gv.stride(0), 0, ptr(gs), ptr(gt), N, h, int(bool(ctx.inverse)),
                 stream_ptr(dev))
        return gv, gs, gt, None


# --------------------------------------------------------------------------------------
# planar / radial (nf/flows_1.py:42-60, :85-97)
# --------------------------------------------------------------------------------------
def planar_prepare(w, u):
    dev = require_cuda(w, u)
    w, u = f32c(w), f32c(u)
    L, d = w.shape
    uhat = torch.empty_like(w)
    wuhat = torch.empty((L,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("nfk_planar_prepare", ptr(w), ptr(u), ptr(uhat), ptr(wuhat), d, L, stream_ptr(dev))
    return uhat, wuhat


def planar_stack(x, w, u, b, logdet=None):
    """L planar layers in one pass.  w,u [L,d], b [L].  Returns (z, logdet)."""
    dev = require_cuda(x, w, u, b, logdet)
    x, w, b = f32c(x), f32c(w), f32c(b.reshape(-1))
    L, d = w.shape
    N = x.shape[0]
    uhat, wuhat = planar_prepare(w, u)
    accumulate = logdet is not None
    if not accumulate:
        logdet = torch.empty((N,), dtype=torch.float32, device=dev)
    if d == 128 and PLANAR_MMA and _lib.have("nfk_planar_stack_mma") and x.data_ptr() % 8 == 0:
        # Gram-matrix form on the tensor cores (csrc/planar_mma.cu), 32 layers per launch
        cur = x
        with torch.cuda.device(dev):
            for l0 in range(0, L, 32):
                l1 = min(L, l0 + 32)
                wl, ul, bl = w[l0:l1].contiguous(), uhat[l0:l1].contiguous(), b[l0:l1].contiguous()
                gram = torch.empty((l1 - l0, l1 - l0), dtype=torch.float32, device=dev)
                call("nfk_planar_gram", ptr(wl), ptr(ul), ptr(gram), d, l1 - l0, stream_ptr(dev))
                out = torch.empty_like(x)
                call("nfk_planar_stack_mma", ptr(cur), ptr(wl), ptr(ul), ptr(gram), ptr(bl), ptr(out), ptr(logdet), N, d,
                     l1 - l0, int(accumulate or l0 > 0), stream_ptr(dev))
                cur = out
        return cur, logdet
    out = torch.empty_like(x)
    with torch.cuda.device(dev):
        call("nfk_planar_stack", ptr(x), ptr(w), ptr(uhat), ptr(wuhat), ptr(b), ptr(out), ptr(logdet), N, d, L,
             int(accumulate), stream_ptr(dev))
    return out, logdet


PLANAR_MMA = True        # False: the per-layer register-resident kernel for every shape (cross-check)


def radial(x, x0, log_alpha, beta, per_sample=False, logdet=None):
    """Radial layer.  per_sample=False reproduces the reference's batch-global norm (Q9): when a
    process group is initialised the sum of squares is all-reduced, so a sharded batch gives the
    result of the whole batch."""
    dev = require_cuda(x, x0, log_alpha, beta)
    x, x0, log_alpha, beta = f32c(x), f32c(x0), f32c(log_alpha), f32c(beta)
    N, d = x.shape
    out = torch.empty_like(x)
    sumsq = None
    with torch.cuda.device(dev):
        if not per_sample:
            sumsq = torch.zeros((1,), dtype=torch.float32, device=dev)
            call("nfk_radial_sumsq", ptr(x), ptr(x0), ptr(sumsq), N, d, stream_ptr(dev))
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                torch.distributed.all_reduce(sumsq)
        accumulate = logdet is not None
        if not accumulate:
            logdet = torch.empty((N if per_sample else 1,), dtype=torch.float32, device=dev)
        call("nfk_radial", ptr(x), ptr(x0), ptr(log_alpha), ptr(beta), ptr(sumsq), ptr(out), ptr(logdet), N, d,
             int(bool(per_sample)), int(accumulate), stream_ptr(dev))
    return out, logdet


def radial_stack(x, x0, log_alpha, beta, per_sample, logdet=None):
    """A run of L radial layers (x0 [L,d], log_alpha [L], beta [L]) without autograd.
    per_sample: ONE pass over x for the whole run (row kept in registers).  Batch-global (the reference's
    Frobenius norm over the whole batch, quirk Q9): one read + one write per layer -- the pass that applies layer l
    also accumulates the sum of squares layer l+1 needs (all-reduced across ranks between launches).
    Returns (z, logdet): logdet [N] per sample, [1] in the batch-global mode."""
    dev = require_cuda(x, x0, log_alpha, beta, logdet)
    x, x0, log_alpha, beta = f32c(x), f32c(x0), f32c(log_alpha.reshape(-1)), f32c(beta.reshape(-1))
    N, d = x.shape
    L = x0.shape[0]
    accumulate = logdet is not None
    if not accumulate:
        logdet = torch.empty((N if per_sample else 1,), dtype=torch.float32, device=dev)
    if per_sample:
        out = torch.empty_like(x)
        with torch.cuda.device(dev):
            call("nfk_radial_stack", ptr(x), ptr(x0), ptr(log_alpha), ptr(beta), ptr(out), ptr(logdet), N, d, L,
                 int(accumulate), stream_ptr(dev))
        return out, logdet
    multi = torch.distributed.is_available() and torch.distributed.is_initialized()
    sums = torch.zeros((L + 1,), dtype=torch.float32, device=dev)
    bufs = [torch.empty_like(x), torch.empty_like(x) if L > 1 else None]
    with torch.cuda.device(dev):
        call("nfk_radial_global", ptr(x), ptr(None), ptr(None), ptr(None), ptr(None), ptr(x0[0]), ptr(sums[0:1]),
             ptr(None), ptr(None), N, d, 0, stream_ptr(dev))
        cur = x
        for l in range(L):
            if multi:
                torch.distributed.all_reduce(sums[l:l + 1])
            out = bufs[l & 1]
            nxt = x0[l + 1] if l + 1 < L else None
            call("nfk_radial_global", ptr(cur), ptr(x0[l]), ptr(log_alpha[l:l + 1]), ptr(beta[l:l + 1]), ptr(sums[l:l + 1]),
                 ptr(nxt), ptr(sums[l + 1:l + 2]), ptr(out), ptr(logdet), N, d, int(accumulate or l > 0), stream_ptr(dev))
            cur = out
    return cur, logdet


def radial_stack_ok(d: int, per_sample: bool) -> bool:
    if per_sample:
        return _lib.have("nfk_radial_stack") and d in (32, 64, 128, 256)
    return _lib.have("nfk_radial_global") and d % 4 == 0 and 1024 % d == 0


# --------------------------------------------------------------------------------------
# log-prob reduction, gather, leapfrog
# --------------------------------------------------------------------------------------
def gauss_logprob(z, var: float = 1.0, add=None, add_sign: float = 1.0):
    """log N(z; 0, var I) (+ add_sign * add)  —  nf/models.py:19-20, :34, :39."""
    dev = require_cuda(z, add)
    z = f32c(z)
    N, d = z.shape
    add = f32c(add) if add is not None else None
    out = torch.empty((N,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("nfk_gauss_logprob", ptr(z), ptr(add), float(add_sign), ptr(out), N, d, float(var), stream_ptr(dev))
    return out


class GaussLogprobFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, var):
        ctx.save_for_backward(z)
        ctx.var = var
        return gauss_logprob(z, var)

    @staticmethod
    def backward(ctx, g):
        (z,) = ctx.saved_tensors
        return -(g[:, None] * z) / ctx.var, None


def gather_cols(x, size, dim, cols, bf16=False, ld_out=None):
    """x[:, :, cols].flatten(1) of x viewed [N, size, dim] (nf/flows.py:230)."""
    dev = require_cuda(x)
    x = f32c(x)
    N = x.shape[0]
    width = size * len(cols)
    ld_out = width if ld_out is None else ld_out
    out = torch.zeros((N, ld_out), dtype=torch.bfloat16 if bf16 else torch.float32, device=dev) \
        if ld_out != width else torch.empty((N, width), dtype=torch.bfloat16 if bf16 else torch.float32, device=dev)
    c = i32_array(cols)
    with torch.cuda.device(dev):
        call("nfk_gather_cols", ptr(x), ptr(out), N, size, dim, c, len(cols), int(bf16), ld_out, stream_ptr(dev))
    return out


def leapfrog_kick_drift(q, p, force, dt, inv_mass=1.0):
    dev = require_cuda(q, p, force)
    with torch.cuda.device(dev):
        call("nfk_leapfrog_kick_drift", ptr(q), ptr(p), ptr(force), q.numel(), float(dt), float(inv_mass),
             stream_ptr(dev))


def leapfrog_kick(p, force, dt):
    dev = require_cuda(p, force)
    with torch.cuda.device(dev):
        call("nfk_leapfrog_kick", ptr(p), ptr(force), p.numel(), float(dt), stream_ptr(dev))


class PlanarStackFn(torch.autograd.Function):
    """(z, log_det) = L fused planar layers; w,u [L,d], b [L]."""

    @staticmethod
    def forward(ctx, x, w, u, b):
        out, ld = planar_stack(x, w, u, b)
        ctx.save_for_backward(x, w, u, b)
        return out, ld

    @staticmethod
    def backward(ctx, g_out, g_ld):
        x, w, u, b = ctx.saved_tensors
        return planar_stack_bwd(x, w, u, b, g_out, g_ld)


def planar_stack_bwd(x, w, u, b, g_out, g_ld):
    """Backward of the fused stack: the kernel returns grads w.r.t. x, w (direct path), uhat,
    b and w.uhat; the parameter-sized chain uhat(w,u), w.uhat -> (w,u) is closed here with
    autograd on [L,d] tensors."""
    dev = require_cuda(x, w, u, b)
    x = f32c(x)
    N, d = x.shape
    L = w.shape[0]
    with torch.enable_grad():
        w_ = f32c(w).detach().requires_grad_(True)
        u_ = f32c(u).detach().requires_grad_(True)
        wu = (w_ * u_).sum(1, keepdim=True)
        scal = torch.log(1 + torch.exp(wu)) - wu - 1
        uhat = u_ + scal * w_ / (torch.norm(w_, dim=1, keepdim=True) ** 2)   # flows_1.py:52-53
        wuhat = (w_ * uhat).sum(1)
    g_out = f32c(g_out) if g_out is not None else torch.zeros_like(x)
    g_ld = f32c(g_ld) if g_ld is not None else None
    gx = torch.empty_like(x)
    gw = torch.zeros((L, d), dtype=torch.float32, device=dev)
    guh = torch.zeros((L, d), dtype=torch.float32, device=dev)
    gb = torch.zeros((L,), dtype=torch.float32, device=dev)
    gwuh = torch.zeros((L,), dtype=torch.float32, device=dev)
    bb = f32c(b.reshape(-1))
    with torch.cuda.device(dev):
        call("nfk_planar_stack_bwd", ptr(x), ptr(w_.detach()), ptr(uhat.detach().contiguous()),
             ptr(wuhat.detach().contiguous()), ptr(bb), ptr(g_out), ptr(g_ld), ptr(gx), ptr(gw), ptr(guh),
             ptr(gb), ptr(gwuh), N, d, L, stream_ptr(dev))
    gw2, gu2 = torch.autograd.grad([uhat, wuhat], [w_, u_], [guh, gwuh])
    return gx, (gw + gw2).view_as(w), gu2.view_as(u), gb.view_as(b)


class RadialFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, x0, log_alpha, beta, per_sample):
        out, ld = radial(x, x0, log_alpha, beta, per_sample)
        ctx.save_for_backward(x, x0, log_alpha, beta)
        ctx.per_sample = per_sample
        return out, ld

    @staticmethod
    def backward(ctx, g_out, g_ld):
        x, x0, log_alpha, beta = ctx.saved_tensors
        dev = x.device
        xx, x0c, la, bb = f32c(x), f32c(x0), f32c(log_alpha), f32c(beta)
        N, d = xx.shape
        g_out = f32c(g_out) if g_out is not None else torch.zeros_like(xx)
        g_ld = f32c(g_ld) if g_ld is not None else None
        gx = torch.empty_like(xx)
        gx0 = torch.zeros(d, dtype=torch.float32, device=dev)
        gla = torch.zeros(1, dtype=torch.float32, device=dev)
        gb = torch.zeros(1, dtype=torch.float32, device=dev)
        sumsq = dot = None
        with torch.cuda.device(dev):
            if not ctx.per_sample:
                sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
                dot = torch.zeros(1, dtype=torch.float32, device=dev)
                call("nfk_radial_sumsq", ptr(xx), ptr(x0c), ptr(sumsq), N, d, stream_ptr(dev))
                call("nfk_radial_dot", ptr(xx), ptr(x0c), ptr(g_out), ptr(dot), N, d, stream_ptr(dev))
                if torch.distributed.is_available() and torch.distributed.is_initialized():
                    both = torch.cat([sumsq, dot])
                    torch.distributed.all_reduce(both)
                    sumsq, dot = both[:1].contiguous(), both[1:].contiguous()
            call("nfk_radial_bwd", ptr(xx), ptr(x0c), ptr(la), ptr(bb), ptr(sumsq), ptr(dot), ptr(g_out), ptr(g_ld),
                 ptr(gx), ptr(gx0), ptr(gla), ptr(gb), N, d, int(bool(ctx.per_sample)), stream_ptr(dev))
        return gx.view_as(x), gx0.view_as(x0), gla.view_as(log_alpha), gb.view_as(beta), None
