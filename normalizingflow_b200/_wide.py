"""Host side of the wide conditioner path (csrc/gemm_ws.cu): FCNN (nf/flows.py:20-35) with a hidden
width above 128 — e.g. the class default 800 (nf/flows.py:216) — as three launches of the
persistent warp-specialised tcgen05 GEMM.  Every operand lives in HBM as the shared-memory image
the tensor core reads (128 x 64 bf16 blocks, K-major SWIZZLE_128B), so the kernels move whole
16 KB blocks with 1-D TMA bulk copies; weights are packed once per parameter version."""
from __future__ import annotations

import torch

from . import _lib, _ops
from ._lib import call, f32c, i32_array, ptr, require_cuda, stream_ptr

ROWS = 128
MIN_HIDDEN = 129      # narrower conditioners use the fused layer kernel / nfk_linear_bf16
BF16, F16 = _lib.IMG_BF16, _lib.IMG_F16
# Element format of the operand images on the no-grad (inference) forward: IEEE fp16 -- the activations
# are tanh outputs in (-1, 1), the weights O(1) and the conditioning inputs saturate at +-65504, so fp16's
# range suffices and its 11-bit significand carries 8x less quantisation noise than bf16 at the same
# tensor-core rate.  The gradient paths (training, log-prob gradients) keep every image in bf16: gradient
# images need bf16's range and the backward kernels read the saved activations as bf16.
INFER_FMT = F16
_TORCH_FMT = {BF16: torch.bfloat16, F16: torch.float16}


def available() -> bool:
    return _lib.have("nfk_gemm_ws") and _lib.have("nfk_pack_a_img")


def usable(fcnn) -> bool:
    net = getattr(fcnn, "network", None)
    return (available() and getattr(fcnn, "precision", None) == "bf16" and net is not None and len(net) == 5
            and net[0].out_features >= MIN_HIDDEN)


def blocks(n: int) -> int:
    return (n + 63) // 64


def plan_tiles(n_blocks: int):
    """Split n_blocks 64-column blocks into the fewest N tiles of <= 4 blocks, evenly sized."""
    n_tiles = (n_blocks + 3) // 4
    base, rem = divmod(n_blocks, n_tiles)
    return [base + (1 if t < rem else 0) for t in range(n_tiles)]


def _swizzle_image(mat):
    """[rows, 64*KB] bf16 -> [KB, rows, 8, 8]: 16-byte chunk j of row r holds source chunk
    j ^ (r % 8) (the canonical K-major SWIZZLE_128B layout tcgen05.mma reads)."""
    rows, kp = mat.shape
    kb = kp // 64
    blk = mat.reshape(rows, kb, 8, 8).permute(1, 0, 2, 3)
    r = torch.arange(rows, device=mat.device)[:, None]
    j = torch.arange(8, device=mat.device)[None, :]
    return blk[:, r, j ^ (r & 7), :].contiguous()


def weight_image(w, bias, kb: int, tiles, fmt=BF16):
    """(w_img, bias_pad) for W [n_out, k_in] and the N-tile plan ``tiles``."""
    n_out, k_in = w.shape
    ob = sum(tiles)
    dev = w.device
    dt = _TORCH_FMT[fmt]
    wp = torch.zeros((ob * 64, kb * 64), dtype=dt, device=dev)
    wd = w.detach().float()
    wp[:n_out, :k_in] = (wd.clamp(-65504.0, 65504.0) if fmt == F16 else wd).to(dt)
    parts, r0 = [], 0
    for nb in tiles:
        parts.append(_swizzle_image(wp[r0:r0 + nb * 64]).reshape(-1))
        r0 += nb * 64
    bp = torch.zeros(ob * 64, dtype=torch.float32, device=dev)
    if bias is not None:
        bp[:n_out] = bias.detach().float()
    return torch.cat(parts).contiguous(), bp


def pack_weight(w, kb, tiles, transposed=False, pad_rows=False, pad_k=False, fmt=BF16):
    """16-bit W image of the fp32 matrix ``w`` in ONE launch (nfk_pack_w_img)."""
    dev = require_cuda(w)
    w = f32c(w.detach())
    img = torch.empty(sum(tiles) * 64 * kb * 64, dtype=_TORCH_FMT[fmt], device=dev)
    with torch.cuda.device(dev):
        call("nfk_pack_w_img", ptr(w), w.stride(0), w.shape[0], w.shape[1], ptr(img), kb, i32_array(tiles), len(tiles),
             int(transposed), int(pad_rows), int(pad_k), fmt, stream_ptr(dev))
    return img


def _pad_bias(bias, n, dev):
    bp = torch.zeros(n, dtype=torch.float32, device=dev)
    if bias is not None:
        bp[:bias.numel()] = bias.detach().float()
    return bp


def packed(fcnn, fmt=BF16):
    """Per-layer (w_img, bias, KB, kmma_last, tiles, n_out, fmt) cached on the module, one entry per format."""
    layers = (fcnn.network[0], fcnn.network[2], fcnn.network[4])
    key = (_lib.param_epoch(),) + tuple(
        (l.weight._version, l.weight.data_ptr(), l.bias._version if l.bias is not None else -1) for l in layers)
    caches = fcnn.__dict__.setdefault("_wide_cache", {})
    cache = caches.get(fmt)
    if cache is not None and cache[0] == key:
        return cache[1]
    out = []
    kb = blocks(layers[0].in_features)
    for l in layers:
        n_out, k_in = l.weight.shape
        tiles = plan_tiles(blocks(n_out))
        if l.weight.is_cuda:
            w_img, b = pack_weight(l.weight, kb, tiles, fmt=fmt), _pad_bias(l.bias, sum(tiles) * 64, l.weight.device)
        else:
            w_img, b = weight_image(l.weight, l.bias, kb, tiles, fmt)
        kmma_last = (k_in - 64 * (kb - 1) + 15) // 16
        out.append(dict(w=w_img, b=b, KB=kb, kmma_last=kmma_last, tiles=tiles, tiles_c=i32_array(tiles),
                        n_out=n_out, fmt=fmt))
        kb = sum(tiles)
    caches[fmt] = (key, out)
    return out


def pack_input(x, size, dim, cols, kb, fmt=BF16):
    """x[:, :, cols].flatten(1) (nf/flows.py:230) -> 16-bit A image."""
    dev = require_cuda(x)
    x = f32c(x)
    N = x.shape[0]
    m_tiles = (N + ROWS - 1) // ROWS
    img = torch.empty((m_tiles, kb, ROWS, 64), dtype=_TORCH_FMT[fmt], device=dev)
    with torch.cuda.device(dev):
        call("nfk_pack_a_img", ptr(x), ptr(img), N, size, dim, i32_array(cols), len(cols), kb, fmt, stream_ptr(dev))
    return img


def gemm(a_img, layer, M, act, out_f32, tag="gemm_ws", aux=None):
    dev = a_img.device
    m_tiles = (M + ROWS - 1) // ROWS
    ob = sum(layer["tiles"])
    fmt = layer.get("fmt", BF16)
    if a_img.dtype != _TORCH_FMT[fmt]:
        raise ValueError(f"gemm: activation image is {a_img.dtype}, weight image is {_TORCH_FMT[fmt]}")
    if out_f32:
        out = torch.empty((M, layer["n_out"]), dtype=torch.float32, device=dev)
        ldy = layer["n_out"]
    else:
        out = torch.empty((m_tiles, ob, ROWS, 64), dtype=_TORCH_FMT[fmt], device=dev)
        ldy = 0
    with torch.cuda.device(dev):
        tm = _ops.KERNEL_TIMER
        ev = tm.start(tag, dev) if tm is not None else None
        call("nfk_gemm_ws", ptr(a_img), ptr(layer["w"]), ptr(layer["b"]), ptr(out), M, layer["KB"],
             layer["kmma_last"], layer["tiles_c"], len(layer["tiles"]), act, int(out_f32), layer["n_out"], ldy,
             ptr(aux), fmt, stream_ptr(dev))
        if ev is not None:
            tm.stop(ev, dev)
    return out


def mlp3(fcnn, x, size=None, dim=1, cols=(0,)):
    """FCNN(x[:, :, cols].flatten(1)) -> fp32 [N, out_dim]; with the defaults x is the plain
    [N, in_dim] conditioner input."""
    l1, l2, l3 = packed(fcnn, INFER_FMT)
    if size is None:
        size = x.shape[1]
    N = x.shape[0]
    a0 = pack_input(x, size, dim, list(cols), l1["KB"], INFER_FMT)
    h1 = gemm(a0, l1, N, 1, False, "gemm_ws_l1")
    h2 = gemm(h1, l2, N, 1, False, "gemm_ws_l2")
    return gemm(h2, l3, N, 0, True, "gemm_ws_l3")


def rqs_eligible(layer) -> bool:
    """Layers whose last GEMM can carry the spline transform as its epilogue (nfk_gemm_ws_rqs)."""
    n_t = layer.size * (layer.dim - len(layer._mask))
    return (_lib.have("nfk_gemm_ws_rqs") and layer.K == 8 and 2 <= layer.dim <= 4 and n_t <= 128
            and 1 <= len(layer._mask) < layer.dim and usable(layer.psi)
            and layer.psi.network[0].in_features == layer.size * len(layer._mask)
            and layer.psi.network[4].out_features == 23 * n_t)


def packed_rqs(layer, fmt=BF16):
    """(l1, l2, l3) of the layer's conditioner with l3 in the per-feature padded (23 -> 24 rows)
    layout of the fused spline epilogue: N tiles of 8 features."""
    fcnn = layer.psi
    l1, l2, _ = packed(fcnn, fmt)
    last = fcnn.network[4]
    key = (_lib.param_epoch(), last.weight._version, last.weight.data_ptr(), last.bias._version)
    caches = layer.__dict__.setdefault("_wide_rqs_cache", {})
    cache = caches.get(fmt)
    if cache is None or cache[0] != key:
        H = last.in_features
        dev = last.weight.device
        n_t = last.out_features // 23
        n_tiles = (n_t + 7) // 8
        b3 = torch.zeros((n_tiles * 8, 24), dtype=torch.float32, device=dev)
        b3[:n_t, :23] = last.bias.detach().float().reshape(n_t, 23)
        kb = sum(l2["tiles"])
        w_img = pack_weight(last.weight, kb, [3] * n_tiles, pad_rows=True, fmt=fmt)      # 23 -> 24 rows per feature
        l3 = dict(w=w_img, b=b3.reshape(-1), KB=kb, kmma_last=(H - 64 * (kb - 1) + 15) // 16, fmt=fmt)
        caches[fmt] = cache = (key, l3)
    return l1, l2, cache[1]


def run_layer(layer, x, inverse, logdet=None, debug=False):
    """(out, logdet) of one eligible NSF_CL layer: pack + 2 GEMMs + GEMM-with-spline-epilogue.
    ``debug`` (test hook): also return the raw spline parameters [N, F_t, 23] the epilogue computed and the
    bins [N, F_t] it used."""
    dev = require_cuda(x, logdet)
    x = f32c(x)
    N = x.shape[0]
    l1, l2, l3 = packed_rqs(layer, INFER_FMT)
    a0 = pack_input(x, layer.size, layer.dim, layer._mask, l1["KB"], INFER_FMT)
    h1 = gemm(a0, l1, N, 1, False, "gemm_ws_l1")
    h2 = gemm(h1, l2, N, 1, False, "gemm_ws_l2")
    out = torch.empty((N, layer.size * layer.dim), dtype=torch.float32, device=dev)
    accumulate = logdet is not None
    if not accumulate:
        logdet = torch.empty((N,), dtype=torch.float32, device=dev)
    n_t = layer.size * (layer.dim - len(layer._mask))
    ftp = (n_t + 7) // 8 * 8
    dbg_p = torch.zeros((N, ftp, 24), dtype=torch.float32, device=dev) if debug else None
    dbg_b = torch.full((N, ftp), -1, dtype=torch.int8, device=dev) if debug else None
    with torch.cuda.device(dev):
        tm = _ops.KERNEL_TIMER
        ev = tm.start("gemm_ws_rqs", dev) if tm is not None else None
        call("nfk_gemm_ws_rqs", ptr(h2), ptr(l3["w"]), ptr(l3["b"]), ptr(x), ptr(out), ptr(logdet), N, l3["KB"],
             l3["kmma_last"], layer.size, layer.dim, i32_array(layer._mask), len(layer._mask), float(layer.B),
             int(bool(inverse)), int(accumulate), _ops._arith(layer.arith), INFER_FMT, ptr(dbg_p), ptr(dbg_b),
             stream_ptr(dev))
        if ev is not None:
            tm.stop(ev, dev)
    if debug:
        return out, logdet, dbg_p[:, :n_t, :23].contiguous(), dbg_b[:, :n_t].contiguous()
    return out, logdet


# ---------------------------------------------------------------------------------------------
# log-prob gradient path (flow-preconditioned HMC, nf/hmc.py + SURVEY 8(a) row H): forward that
# keeps the hidden-activation images, backward = spline adjoint as a GEMM epilogue + 3 dgrad GEMMs
# ---------------------------------------------------------------------------------------------
def grad_eligible(layer) -> bool:
    """Layers whose d(out, logdet)/dx runs on the tensor-core path (any hidden width)."""
    psi = layer.psi
    net = getattr(psi, "network", None)
    n_t = layer.size * (layer.dim - len(layer._mask))
    return (_lib.have("nfk_gemm_ws_rqs_bwd") and available() and getattr(psi, "precision", None) == "bf16"
            and net is not None and len(net) == 5 and layer.K == 8 and 2 <= layer.dim <= 4 and n_t <= 128
            and 1 <= len(layer._mask) < layer.dim and net[0].in_features == layer.size * len(layer._mask)
            and net[4].out_features == 23 * n_t)


def _transposed(w, kb, pad_k=False):
    """Layer dict for out = A @ w (w [k_in, n_out] as stored, nn.Linear weight of the forward layer):
    the dgrad GEMM of y = x w^T.  ``pad_k``: the K index runs over 24-per-feature padded spline
    parameters.  No bias."""
    n_out = w.shape[1]
    k_log = kb * 64 if pad_k else w.shape[0]
    tiles = plan_tiles(blocks(n_out))
    w_img = pack_weight(w, kb, tiles, transposed=True, pad_k=pad_k)
    return dict(w=w_img, b=torch.zeros(sum(tiles) * 64, dtype=torch.float32, device=w.device), KB=kb,
                kmma_last=min(4, (k_log - 64 * (kb - 1) + 15) // 16), tiles=tiles, tiles_c=i32_array(tiles),
                n_out=n_out)


def packed_bwd(layer):
    """dgrad operands (W3p^T, W2^T, W1^T images), cached per parameter version."""
    net = layer.psi.network
    l0, l2, l4 = net[0], net[2], net[4]
    key = (_lib.param_epoch(),) + tuple((l.weight._version, l.weight.data_ptr()) for l in (l0, l2, l4))
    cache = getattr(layer, "_wide_bwd_cache", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    n_t = l4.out_features // 23
    n_tiles = (n_t + 7) // 8
    t3 = _transposed(l4.weight, 3 * n_tiles, pad_k=True)          # dH2 = G @ W3p
    t2 = _transposed(l2.weight, blocks(l2.out_features))          # dH1 = dZ2 @ W2
    t1 = _transposed(l0.weight, blocks(l0.out_features))          # dXc = dZ1 @ W1
    layer._wide_bwd_cache = (key, (t3, t2, t1))
    return t3, t2, t1


def layer_forward_saving(layer, x, inverse, logdet=None):
    """run_layer that also returns what the backward needs: (out, logdet, (x, h1, h2))."""
    dev = require_cuda(x, logdet)
    x = f32c(x)
    N = x.shape[0]
    l1, l2, l3 = packed_rqs(layer)
    a0 = pack_input(x, layer.size, layer.dim, layer._mask, l1["KB"])
    h1 = gemm(a0, l1, N, 1, False, "gemm_ws_l1")
    h2 = gemm(h1, l2, N, 1, False, "gemm_ws_l2")
    out = torch.empty((N, layer.size * layer.dim), dtype=torch.float32, device=dev)
    accumulate = logdet is not None
    if not accumulate:
        logdet = torch.empty((N,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("nfk_gemm_ws_rqs", ptr(h2), ptr(l3["w"]), ptr(l3["b"]), ptr(x), ptr(out), ptr(logdet), N, l3["KB"],
             l3["kmma_last"], layer.size, layer.dim, i32_array(layer._mask), len(layer._mask), float(layer.B),
             int(bool(inverse)), int(accumulate), _ops._arith(layer.arith), BF16, ptr(None), ptr(None), stream_ptr(dev))
    return out, logdet, (x, h1, h2, bool(inverse))


def layer_backward(layer, ctx, g_out, g_logdet=None, g_logdet_const=1.0, keep=False):
    """dL/dx of one layer from dL/dout [N, d] and dL/dlogdet ([N] tensor, or a constant for all rows).
    ``keep``: also return the gradient images (dL/dparams, dL/dz2, dL/dz1) the weight gradients need."""
    x, h1, h2, inverse = ctx
    dev = require_cuda(x, g_out, g_logdet)
    g_out = f32c(g_out)
    N, d = x.shape
    _, _, l3 = packed_rqs(layer)
    t3, t2, t1 = packed_bwd(layer)
    m_tiles = (N + ROWS - 1) // ROWS
    g_in = torch.empty((N, d), dtype=torch.float32, device=dev)
    g_img = torch.empty((m_tiles, t3["KB"], ROWS, 64), dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        call("nfk_gemm_ws_rqs_bwd", ptr(h2), ptr(l3["w"]), ptr(l3["b"]), ptr(x), ptr(g_out),
             ptr(f32c(g_logdet)) if g_logdet is not None else ptr(None), float(g_logdet_const), ptr(g_in),
             ptr(g_img), N, l3["KB"], l3["kmma_last"], layer.size, layer.dim, i32_array(layer._mask),
             len(layer._mask), float(layer.B), int(inverse), stream_ptr(dev))
    dz2 = gemm(g_img, t3, N, 2, False, "gemm_ws_d3", aux=h2)        # (G W3p) * (1 - h2^2)
    dz1 = gemm(dz2, t2, N, 2, False, "gemm_ws_d2", aux=h1)          # (dZ2 W2) * (1 - h1^2)
    dxc = gemm(dz1, t1, N, 0, True, "gemm_ws_d1")                   # dZ1 W1  [N, size*n_mask] fp32
    with torch.cuda.device(dev):
        call("nfk_scatter_add_cols", ptr(g_in), ptr(dxc), N, layer.size, layer.dim, i32_array(layer._mask),
             len(layer._mask), stream_ptr(dev))
    if keep:
        return g_in, (g_img, dz2, dz1)
    return g_in


def unpack_rows(img, M, ncols):
    """bf16 image -> row-major bf16 [M, pad8(ncols)] (view of the first ncols columns returned)."""
    dev = img.device
    kb = img.shape[1]
    ld = (ncols + 7) // 8 * 8
    rows = torch.empty((M, ld), dtype=torch.bfloat16, device=dev)
    if ld > (ncols // 8) * 8:
        rows[:, (ncols // 8) * 8:] = 0
    with torch.cuda.device(dev):
        call("nfk_unpack_img_rows", ptr(img), ptr(rows), M, kb, min(ld, kb * 64), ld, stream_ptr(dev))
    return rows[:, :ncols]


def _mm_f32(a_t, b):
    """TEST-ONLY cross-check of nfk_wgrad_ws (reached only with ``NATIVE_WGRAD = False``, which no product
    path sets): a_t [N, P], b [N, Q] -> a_t.T @ b [P, Q] through the library GEMM."""
    try:
        return torch.mm(a_t.t(), b, out_dtype=torch.float32)
    except TypeError:
        return torch.mm(a_t.t(), b).float()


class NsfWideFn(torch.autograd.Function):
    """One grad_eligible NSF_CL layer for training: forward = pack + 2 GEMMs + GEMM-with-spline
    epilogue (hidden activations stay in HBM as bf16 images), backward = spline adjoint as a GEMM
    epilogue + dgrad GEMMs with fused tanh backward on the tensor cores; weight gradients contract
    the saved images over the batch with library GEMMs."""

    @staticmethod
    def forward(ctx, x, w0, b0, w2, b2, w4, b4, layer, inverse):
        out, ld, saved = layer_forward_saving(layer, x.detach(), inverse)
        xs, h1, h2, inv = saved
        ctx.save_for_backward(xs, h1, h2)
        ctx.layer, ctx.inv = layer, inv
        return out, ld

    @staticmethod
    def backward(ctx, g_out, g_ld):
        xs, h1, h2 = ctx.saved_tensors
        layer = ctx.layer
        N = xs.shape[0]
        if g_out is None:
            g_out = torch.zeros_like(xs)
        if g_ld is None:
            g_ld = torch.zeros(N, dtype=torch.float32, device=xs.device)
        g_in, (g_img, dz2, dz1) = layer_backward(layer, (xs, h1, h2, ctx.inv), g_out, g_ld, keep=True)
        need_w = any(ctx.needs_input_grad[1:7])
        if not need_w:
            return (g_in,) + (None,) * 8
        net = layer.psi.network
        H, n_t, n_in = net[4].in_features, net[4].out_features // 23, net[0].in_features
        n_tiles = (n_t + 7) // 8
        l1 = packed(layer.psi)[0]
        a0 = pack_input(xs, layer.size, layer.dim, layer._mask, l1["KB"])
        if NATIVE_WGRAD:
            if not _lib.have("nfk_wgrad_ws"):
                raise RuntimeError("libnfk.so does not export nfk_wgrad_ws; rebuild the library")
            # batch contraction straight from the images (MN-major tcgen05 operands, split-K)
            gw4 = wgrad(g_img, h2, N, n_t * 23, H, pad_p=True)
            gw2 = wgrad(dz2, h1, N, H, H)
            gw0 = wgrad(dz1, a0, N, H, n_in)
            gb4 = image_colsum(g_img, n_tiles * 192).reshape(n_tiles * 8, 24)[:n_t, :23].reshape(-1)
            gb2 = image_colsum(dz2, H)
            gb0 = image_colsum(dz1, H)
            return g_in, gw0, gb0, gw2, gb2, gw4, gb4, None, None
        G = unpack_rows(g_img, N, n_tiles * 192)                     # dL/dparams, 24 columns per feature
        h2r, h1r = unpack_rows(h2, N, H), unpack_rows(h1, N, H)
        dz2r, dz1r = unpack_rows(dz2, N, H), unpack_rows(dz1, N, H)
        xc = unpack_rows(a0, N, n_in)
        gw4 = _mm_f32(G, h2r).reshape(n_tiles * 8, 24, H)[:n_t, :23].reshape(n_t * 23, H)
        gb4 = torch.sum(G, dim=0, dtype=torch.float32).reshape(n_tiles * 8, 24)[:n_t, :23].reshape(-1)
        gw2 = _mm_f32(dz2r, h1r)
        gb2 = torch.sum(dz2r, dim=0, dtype=torch.float32)
        gw0 = _mm_f32(dz1r, xc)
        gb0 = torch.sum(dz1r, dim=0, dtype=torch.float32)
        return g_in, gw0, gb0, gw2, gb2, gw4, gb4, None, None


class Mlp3WideFn(torch.autograd.Function):
    """A whole FCNN (nf/flows.py:20-35) with bf16 tensor-core GEMMs in BOTH directions, for layers that
    are not grad_eligible NSF_CL layers (RealNVP s/t nets, NSF_AR conditioners, K != 8): forward = pack +
    3 x nfk_gemm_ws keeping the activation images; backward = the gradient packed as an image, two
    act = 2 dgrad GEMMs (tanh backward fused), one fp32-row dgrad GEMM, and nfk_wgrad_ws for the three
    weight gradients (bias gradients = column sums of the gradient images)."""

    @staticmethod
    def forward(ctx, x, w0, b0, w2, b2, w4, b4, fcnn):
        l1, l2, l3 = packed(fcnn)
        xs = f32c(x.detach())
        N, n_in = xs.shape
        a0 = pack_input(xs, n_in, 1, [0], l1["KB"])
        h1 = gemm(a0, l1, N, 1, False, "gemm_ws_l1")
        h2 = gemm(h1, l2, N, 1, False, "gemm_ws_l2")
        out = gemm(h2, l3, N, 0, True, "gemm_ws_l3")
        ctx.save_for_backward(a0, h1, h2)
        ctx.fcnn, ctx.N = fcnn, N
        return out

    @staticmethod
    def backward(ctx, g):
        a0, h1, h2 = ctx.saved_tensors
        fcnn, N = ctx.fcnn, ctx.N
        net = fcnn.network
        l0, l2, l4 = net[0], net[2], net[4]
        H, n_out, n_in = l4.in_features, l4.out_features, l0.in_features
        key = (_lib.param_epoch(),) + tuple((l.weight._version, l.weight.data_ptr()) for l in (l0, l2, l4))
        cache = getattr(fcnn, "_wide_bwd_cache", None)
        if cache is None or cache[0] != key:
            cache = fcnn._wide_bwd_cache = (key, (_transposed(l4.weight, blocks(n_out)),
                                                  _transposed(l2.weight, blocks(H)),
                                                  _transposed(l0.weight, blocks(H))))
        t3, t2, t1 = cache[1]
        g_img = pack_input(f32c(g), n_out, 1, [0], t3["KB"])
        dz2 = gemm(g_img, t3, N, 2, False, "gemm_ws_d3", aux=h2)
        dz1 = gemm(dz2, t2, N, 2, False, "gemm_ws_d2", aux=h1)
        gx = gemm(dz1, t1, N, 0, True, "gemm_ws_d1") if ctx.needs_input_grad[0] else None
        if not any(ctx.needs_input_grad[1:7]):
            return (gx,) + (None,) * 7
        gw4, gb4 = wgrad(g_img, h2, N, n_out, H), image_colsum(g_img, n_out)
        gw2, gb2 = wgrad(dz2, h1, N, H, H), image_colsum(dz2, H)
        gw0, gb0 = wgrad(dz1, a0, N, H, n_in), image_colsum(dz1, H)
        return gx, gw0, gb0, gw2, gb2, gw4, gb4, None


def mlp3_grad_ok(fcnn) -> bool:
    net = getattr(fcnn, "network", None)
    return (available() and _lib.have("nfk_wgrad_ws") and _lib.have("nfk_pack_w_img")
            and getattr(fcnn, "precision", None) == "bf16" and net is not None and len(net) == 5
            and all(net[i].bias is not None for i in (0, 2, 4)) and net[0].weight.is_cuda)


NATIVE_WGRAD = True      # False: weight gradients through image -> rows + library GEMMs (cross-check path)


def wgrad(a_img, b_img, M, P, Q, pad_p=False):
    """C [P, Q] = sum over batch rows of A[n, p] * B[n, q] from two bf16 images (nfk_wgrad_ws)."""
    dev = a_img.device
    c = torch.zeros((P, Q), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        tm = _ops.KERNEL_TIMER
        ev = tm.start("wgrad_ws", dev) if tm is not None else None
        call("nfk_wgrad_ws", ptr(a_img), ptr(b_img), ptr(c), Q, M, a_img.shape[1], b_img.shape[1], P, Q, int(pad_p),
             stream_ptr(dev))
        if ev is not None:
            tm.stop(ev, dev)
    return c


def image_colsum(img, ncols):
    """fp32 column sums of a bf16 image over all rows (bias gradients): one reduction over the image,
    then the 16-byte-chunk swizzle is undone on the [KB, 8, 8, 8] result."""
    mt, kb, rows, _ = img.shape
    t = img.view(mt, kb, rows // 8, 8, 8, 8).sum(dim=(0, 2), dtype=torch.float32)      # [kb, r%8, slot, e]
    r7 = torch.arange(8, device=img.device)[:, None]
    ch = torch.arange(8, device=img.device)[None, :]
    sel = t[:, r7, ch ^ r7, :]                                                       # chunk c sits at slot c ^ (r%8)
    return sel.sum(dim=1).reshape(kb * 64)[:ncols]


def flow_grad_eligible(model) -> bool:
    """All layers are grad_eligible NSF_CL layers under a GaussianPrior."""
    from .flows import NSF_CL
    return (hasattr(model, "_prior_var") and model._prior_var() is not None and len(model.flows) > 0
            and all(isinstance(f, NSF_CL) and grad_eligible(f) for f in model.flows))


def flow_logp_and_grad(model, x):
    """log p(x) [N] and d log p / dx [N, d] for a NormalizingFlowModel whose layers are all
    grad_eligible NSF_CL layers under a GaussianPrior; None when the model does not qualify."""
    if not flow_grad_eligible(model):
        return None
    var = model._prior_var()
    h = f32c(x.detach())
    logdet = torch.zeros(h.shape[0], dtype=torch.float32, device=h.device)
    ctxs = []
    for f in model.flows:
        h, logdet, ctx = layer_forward_saving(f, h, False, logdet)
        ctxs.append(ctx)
    logp = _ops.gauss_logprob(h, var, add=logdet, add_sign=1.0)
    g = h * (-1.0 / var)                                      # d log N(z; 0, var I) / dz
    for f, ctx in zip(reversed(model.flows), reversed(ctxs)):
        g = layer_backward(f, ctx, g, None, 1.0)
    return logp, g


# ---------------------------------------------------------------------------------------------
# NSF_AR (nf/flows.py:152-209): the dim-1 per-dimension conditioners as THREE grouped launches
# ---------------------------------------------------------------------------------------------
def nsf_ar_grouped_ok(layer) -> bool:
    ls = list(layer.layers)
    return (_lib.have("nfk_gemm_ws_grouped") and _lib.have("nfk_nsf_ar_pack") and available() and 2 <= layer.dim <= 128
            and len(ls) == layer.dim - 1 and all(getattr(f, "precision", None) == "bf16" for f in ls)
            and all(hasattr(f, "network") and len(f.network) == 5 for f in ls)
            and len({f.network[0].out_features for f in ls}) == 1)


def _group_table(records, dev):
    """device array of WsGroup records {a_img, w_img, bias, out, KB, kmma_last, a_kb, pad}"""
    import numpy as np
    dt = np.dtype([("a", "<u8"), ("w", "<u8"), ("b", "<u8"), ("o", "<u8"), ("kb", "<i4"), ("km", "<i4"),
                   ("akb", "<i4"), ("pad", "<i4")])
    assert dt.itemsize == _lib.lib.nfk_gemm_ws_group_bytes()
    arr = np.array(records, dtype=dt)
    return torch.from_numpy(arr.view(np.uint8).copy()).to(dev)


def _nsf_ar_packs(layer):
    """Per conditioner (l1, l2, l3) with l1 packed for the INTERLEAVED [cos_0, sin_0, cos_1, ...] input
    order of the shared feature image (the reference order is [cos_0..cos_{i-1}, sin_0..sin_{i-1}])."""
    key = (_lib.param_epoch(),) + tuple(
        (f.network[0].weight._version, f.network[2].weight._version, f.network[4].weight._version,
         f.network[0].weight.data_ptr()) + tuple(f.network[i].bias._version for i in (0, 2, 4))
        for f in layer.layers)
    cache = getattr(layer, "_ar_cache", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    out = []
    for g, f in enumerate(layer.layers):
        i = g + 1
        l0 = f.network[0]
        perm = torch.arange(2 * i, device=l0.weight.device).reshape(2, i).t().reshape(-1)   # [0, i, 1, i+1, ...]
        kb = blocks(2 * i)
        tiles = plan_tiles(blocks(l0.out_features))
        l1 = dict(w=pack_weight(l0.weight.detach()[:, perm].contiguous(), kb, tiles, fmt=INFER_FMT), fmt=INFER_FMT,
                  b=_pad_bias(l0.bias, sum(tiles) * 64, l0.weight.device), KB=kb,
                  kmma_last=(2 * i - 64 * (kb - 1) + 15) // 16, tiles=tiles, tiles_c=i32_array(tiles),
                  n_out=l0.out_features)
        _, l2, l3 = packed(f, INFER_FMT)
        out.append((l1, l2, l3))
    layer._ar_cache = (key, out)
    layer._ar_tables = {}                 # group tables hold raw addresses of the old weight images
    return out


def nsf_ar_params(layer, x, max_rows=None):
    """Raw spline parameters [N, dim, 3K-1] of an NSF_AR layer for known inputs x (forward direction):
    one launch builds the shared trig-feature image, three grouped GEMM launches run the dim-1
    conditioners (work items interleave the groups), row 0 is ``init_param``."""
    dev = require_cuda(x)
    x = f32c(x)
    N, dim = x.shape
    P = 3 * layer.K - 1
    G = dim - 1
    packs = _nsf_ar_packs(layer)
    obh = sum(packs[0][0]["tiles"])
    out = torch.empty((N, dim, P), dtype=torch.float32, device=dev)
    out[:, 0, :] = layer.init_param.detach().to(dev).float()
    if max_rows is None:                                           # bound the two hidden-activation buffers to ~8 GB
        max_rows = max(ROWS, int(4e9 // (G * obh * 64 * 2)) // ROWS * ROWS)
    kb_img = blocks(2 * dim)
    tables = getattr(layer, "_ar_tables", None)
    if tables is None:
        tables = layer._ar_tables = {}
    with torch.cuda.device(dev):
        for r0 in range(0, N, max_rows):
            r1 = min(N, r0 + max_rows)
            n = r1 - r0
            mt = (n + ROWS - 1) // ROWS
            a0 = torch.empty((mt, kb_img, ROWS, 64), dtype=_TORCH_FMT[INFER_FMT], device=dev)
            call("nfk_nsf_ar_pack", ptr(x[r0:r1]), ptr(a0), n, dim, float(layer.B), INFER_FMT, stream_ptr(dev))
            h1 = torch.empty((G, mt, obh, ROWS, 64), dtype=_TORCH_FMT[INFER_FMT], device=dev)
            h2 = torch.empty_like(h1)
            ob = out[r0:r1]
            hs = mt * obh * ROWS * 64 * 2
            # the group tables hold raw addresses: rebuilt only when a buffer or a weight image moved
            key = (a0.data_ptr(), h1.data_ptr(), h2.data_ptr(), ob.data_ptr(), n, id(packs))
            tabs = tables.get(key)
            if tabs is None:
                if len(tables) > 8:
                    tables.clear()
                t1 = _group_table([(a0.data_ptr(), packs[g][0]["w"].data_ptr(), packs[g][0]["b"].data_ptr(),
                                    h1.data_ptr() + g * hs, packs[g][0]["KB"], packs[g][0]["kmma_last"], kb_img, 0)
                                   for g in range(G)], dev)
                t2 = _group_table([(h1.data_ptr() + g * hs, packs[g][1]["w"].data_ptr(), packs[g][1]["b"].data_ptr(),
                                    h2.data_ptr() + g * hs, packs[g][1]["KB"], packs[g][1]["kmma_last"], packs[g][1]["KB"], 0)
                                   for g in range(G)], dev)
                t3 = _group_table([(h2.data_ptr() + g * hs, packs[g][2]["w"].data_ptr(), packs[g][2]["b"].data_ptr(),
                                    ob.data_ptr() + (g + 1) * P * 4, packs[g][2]["KB"], packs[g][2]["kmma_last"],
                                    packs[g][2]["KB"], 0) for g in range(G)], dev)
                tabs = tables[key] = (t1, t2, t3)
            tm = _ops.KERNEL_TIMER
            for tab, lay, act, f32, tag in ((tabs[0], packs[0][0], 1, 0, "gemm_ws_grouped_l1"),
                                            (tabs[1], packs[0][1], 1, 0, "gemm_ws_grouped_l2"),
                                            (tabs[2], packs[0][2], 0, 1, "gemm_ws_grouped_l3")):
                ev = tm.start(tag, dev) if tm is not None else None
                call("nfk_gemm_ws_grouped", ptr(tab), G, n, lay["tiles_c"], len(lay["tiles"]), act, f32, P, dim * P,
                     INFER_FMT, stream_ptr(dev))
                if ev is not None:
                    tm.stop(ev, dev)
    return out


def image_to_rows(img, n_rows, n_cols):
    """Inverse of the image layout (tests): [m_tiles, KB, 128, 64] bf16 image -> [n_rows, n_cols]."""
    m_tiles, kb, rows, _ = img.shape
    r = torch.arange(rows, device=img.device)[:, None]
    j = torch.arange(8, device=img.device)[None, :]
    blk = img.reshape(m_tiles, kb, rows, 8, 8)
    un = blk[:, :, r, j ^ (r & 7), :]                    # chunk j of the source sits at slot j ^ (r % 8)
    return un.permute(0, 2, 1, 3, 4).reshape(m_tiles * rows, kb * 64)[:n_rows, :n_cols]
