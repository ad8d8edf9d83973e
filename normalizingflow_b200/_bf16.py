"""bf16 tensor-core path of the conditioner MLP (FCNN, nf/flows.py:20-35): three launches of the
tcgen05 GEMM kernel nfk_linear_bf16 with fused bias (+tanh).  Hidden activations stay bf16
(padded to a multiple of 8 columns = 16 bytes), the last layer writes the fp32 spline
parameters.  Weights are converted to padded bf16 once per parameter version."""
from __future__ import annotations

import torch

from . import _lib, _ops
from ._lib import call, ptr, require_cuda, stream_ptr


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _packed_weights(fcnn):
    """[(W_bf16 [out, pad8(in)], bias fp32)] for the three Linear layers, cached on the module."""
    layers = (fcnn.network[0], fcnn.network[2], fcnn.network[4])
    key = (_lib.param_epoch(),) + tuple(
        (l.weight._version, l.weight.data_ptr(), l.bias._version if l.bias is not None else -1) for l in layers)
    cache = getattr(fcnn, "_bf16_cache", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    packed = []
    k_in = None
    for l in layers:
        w = l.weight.detach()
        out_f, in_f = w.shape
        k_pad = _pad8(in_f) if k_in is None else k_in       # layer input width = previous padded width
        wp = torch.zeros((out_f, k_pad), dtype=torch.bfloat16, device=w.device)
        wp[:, :in_f] = w.to(torch.bfloat16)
        packed.append((wp, l.bias.detach().float().contiguous() if l.bias is not None else None))
        k_in = _pad8(out_f)
    fcnn._bf16_cache = (key, packed)
    return packed


def linear_bf16(x_bf16, w_bf16, bias, act: int, out_f32: bool):
    """act(x @ w.T + bias): x [M, Kp] bf16 (Kp % 8 == 0, zero padded), w [N, Kp] bf16."""
    dev = require_cuda(x_bf16, w_bf16, bias)
    M, Kp = x_bf16.shape
    N = w_bf16.shape[0]
    if w_bf16.shape[1] != Kp:
        raise ValueError(f"weight has K={w_bf16.shape[1]}, input has K={Kp}")
    if out_f32:
        y = torch.empty((M, N), dtype=torch.float32, device=dev)
        ldy = N
    else:
        ldy = _pad8(N)
        y = torch.empty((M, ldy), dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        call("nfk_linear_bf16", ptr(x_bf16), x_bf16.stride(0), ptr(w_bf16), w_bf16.stride(0), ptr(bias), ptr(y),
             ldy, M, Kp, N, act, int(out_f32), stream_ptr(dev))
    return y


def to_bf16_padded(x):
    """fp32 [M, K] -> bf16 [M, pad8(K)] with zero padding (one gather/cast kernel)."""
    if x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0 and x.is_contiguous():
        return x
    M, K = x.shape
    return _ops.gather_cols(x, K, 1, [0], bf16=True, ld_out=_pad8(K))


class MLP3Bf16Fn(torch.autograd.Function):
    """Cross-check path (tests; taken in product code only if libnfk lacks nfk_wgrad_ws, which build()
    rejects): FCNN forward on nfk_linear_bf16, backward through library GEMMs.  The product path for
    bf16 training is _wide.Mlp3WideFn."""

    @staticmethod
    def forward(ctx, x, w0, b0, w2, b2, w4, b4, fcnn):
        (p0, c0), (p2, c2), (p4, c4) = _packed_weights(fcnn)
        xb = to_bf16_padded(x)
        h1 = linear_bf16(xb, p0, c0, 1, False)
        h2 = linear_bf16(h1, p2, c2, 1, False)
        out = linear_bf16(h2, p4, c4, 0, True)
        ctx.save_for_backward(xb, h1, h2, w0, w2, w4)
        return out

    @staticmethod
    def backward(ctx, g):
        xb, h1, h2, w0, w2, w4 = ctx.saved_tensors
        H = w2.shape[0]
        g = g.float()
        h2f, h1f, xf = h2[:, :H].float(), h1[:, :H].float(), xb[:, :w0.shape[1]].float()
        gw4, gb4 = g.t() @ h2f, g.sum(0)
        g2 = (g @ w4.float()) * (1 - h2f * h2f)
        gw2, gb2 = g2.t() @ h1f, g2.sum(0)
        g1 = (g2 @ w2.float()) * (1 - h1f * h1f)
        gw0, gb0 = g1.t() @ xf, g1.sum(0)
        gx = g1 @ w0.float() if ctx.needs_input_grad[0] else None
        return gx, gw0, gb0, gw2, gb2, gw4, gb4, None


def mlp3(fcnn, x):
    l0, l2, l4 = fcnn.network[0], fcnn.network[2], fcnn.network[4]
    if x.dim() != 2:
        x = x.reshape(x.shape[0], -1)
    if torch.is_grad_enabled() and (x.requires_grad or l0.weight.requires_grad):
        from . import _wide
        if x.dtype == torch.float32 and x.is_cuda and _wide.mlp3_grad_ok(fcnn):
            # forward AND backward on the hand-written tensor-core GEMMs (images, dgrad with fused tanh
            # backward, nfk_wgrad_ws); MLP3Bf16Fn below is the library cross-check kept for tests
            return _wide.Mlp3WideFn.apply(x, l0.weight, l0.bias, l2.weight, l2.bias, l4.weight, l4.bias, fcnn)
        return MLP3Bf16Fn.apply(x, l0.weight, l0.bias, l2.weight, l2.bias, l4.weight, l4.bias, fcnn)
    if x.dtype == torch.float32:
        from . import _wide
        if _wide.usable(fcnn):
            return _wide.mlp3(fcnn, x)
    (p0, c0), (p2, c2), (p4, c4) = _packed_weights(fcnn)
    xb = to_bf16_padded(x)
    h1 = linear_bf16(xb, p0, c0, 1, False)
    h2 = linear_bf16(h1, p2, c2, 1, False)
    return linear_bf16(h2, p4, c4, 0, True)
