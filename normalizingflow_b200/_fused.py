"""Host side of the fused NSF coupling-layer kernel (csrc/nsf_fused.cu): packs the conditioner
weights once per parameter version into the fp16 SWIZZLE_128B shared-memory images the kernel
bulk-copies, and launches it.  Eligible layers: size = 32, dim = 2, one masked column, K = 8,
hidden width <= 128 (padded with zero rows/columns to 128, which leaves the MLP unchanged:
tanh(0) = 0 feeds zero weights)."""
from __future__ import annotations

import torch

from . import _lib, _ops
from ._lib import call, f32c, ptr, require_cuda, stream_ptr

ROWS = 128
HP, K1P, NF, PC, CF = 128, 64, 32, 24, 4


# Which kernel generation runs the layer: 2 = csrc/nsf_fused2.cu (three overlapping warp sets), 1 = csrc/nsf_fused.cu.
# The split-operand (fp32-class, precision "fp32x3") configuration exists in generation 2 only.
GENERATION = 2 if _lib.have("nfk_nsf_pairs_fused2") else 1


def eligible(layer) -> bool:
    prec = getattr(layer.psi, "precision", None)
    return (layer.size == 32 and layer.dim == 2 and len(layer._mask) == 1 and layer.K == 8
            and (prec == "bf16" or (prec == "fp32x3" and _lib.have("nfk_nsf_pairs_fused2")))
            and _lib.have("nfk_nsf_pairs_fused")
            and hasattr(layer.psi, "network") and layer.psi.network[0].out_features <= HP
            and layer.psi.network[0].in_features == 32)


def _swizzle_image(mat):
    """[rows, 64*KB] bf16 -> [KB, rows, 8, 8]: K blocks of 128-byte rows whose 16-byte chunk j holds
    source chunk j ^ (row % 8) (the canonical K-major SWIZZLE_128B layout tcgen05.mma reads)."""
    rows, kp = mat.shape
    kb = kp // 64
    blk = mat.reshape(rows, kb, 8, 8).permute(1, 0, 2, 3)
    r = torch.arange(rows, device=mat.device)[:, None]
    j = torch.arange(8, device=mat.device)[None, :]
    src = j ^ (r & 7)
    return blk[:, r, src, :].contiguous()


def _hi_lo(w):
    """fp32 -> (hi, lo) fp16 pair with hi + lo = w to ~22 bits (lo = fp16(w - hi))."""
    w = w.detach().float().clamp(-65504.0, 65504.0)
    hi = w.to(torch.float16)
    lo = (w - hi.float()).to(torch.float16)
    return hi, lo


def packed_split(layer):
    """Operand images of the split-operand (fp32-class) kernel: see nfk_nsf_pairs_fused2 in include/nfk.h."""
    net = layer.psi.network
    l0, l2, l4 = net[0], net[2], net[4]
    key = (_lib.param_epoch(),) + tuple((l.weight._version, l.weight.data_ptr(), l.bias._version)
                                        for l in (l0, l2, l4))
    cache = getattr(layer, "_fused_split_cache", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    dev = l0.weight.device
    H = l0.out_features
    f16 = torch.float16
    h1, o1 = _hi_lo(l0.weight)
    w1 = torch.zeros((HP, K1P), dtype=f16, device=dev)
    w1[:H, :32] = h1                                   # K columns 0..31: hi, 32..63: lo
    w1[:H, 32:64] = o1
    h2, o2 = _hi_lo(l2.weight)
    w2h = torch.zeros((HP, HP), dtype=f16, device=dev)
    w2l = torch.zeros((HP, HP), dtype=f16, device=dev)
    w2h[:H, :H], w2l[:H, :H] = h2, o2
    i2h, i2l = _swizzle_image(w2h), _swizzle_image(w2l)                       # [2 K blocks, 128, 8, 8] each
    w2 = torch.stack([i2h[0], i2l[0], i2h[1], i2l[1]]).contiguous()           # consumption order of the ring
    h3, o3 = _hi_lo(l4.weight)
    w3h = torch.zeros((NF, PC, HP), dtype=f16, device=dev)
    w3l = torch.zeros((NF, PC, HP), dtype=f16, device=dev)
    w3h[:, :23, :H], w3l[:, :23, :H] = h3.reshape(NF, 23, H), o3.reshape(NF, 23, H)
    w3h, w3l = w3h.reshape(NF // CF, CF * PC, HP), w3l.reshape(NF // CF, CF * PC, HP)
    w3 = torch.stack([torch.stack([_swizzle_image(w3h[c]), _swizzle_image(w3l[c])]) for c in range(NF // CF)]).contiguous()
    b1 = torch.zeros(HP, dtype=torch.float32, device=dev)
    b1[:H] = l0.bias.detach().float()
    b2 = torch.zeros(HP, dtype=torch.float32, device=dev)
    b2[:H] = l2.bias.detach().float()
    b3 = torch.zeros((NF, PC), dtype=torch.float32, device=dev)
    b3[:, :23] = l4.bias.detach().float().reshape(NF, 23)
    pk = dict(w1=_swizzle_image(w1), w2=w2, w3=w3, b1=b1, b2=b2, b3=b3.reshape(-1).contiguous())
    layer._fused_split_cache = (key, pk)
    return pk


def _launch(layer, pk, x, out, logdet, n, inverse, accumulate, split, dbg_p=None, dbg_b=None, flags_in=None, flags_out=None,
            epoch=1):
    dev = x.device
    if flags_in is not None or flags_out is not None:
        call("nfk_nsf_pairs_fused2_chain", ptr(x), ptr(out), ptr(logdet), ptr(pk["w1"]), ptr(pk["w2"]), ptr(pk["w3"]),
             ptr(pk["b1"]), ptr(pk["b2"]), ptr(pk["b3"]), n, layer._mask[0], float(layer.B), int(bool(inverse)),
             int(accumulate), _ops._arith(layer.arith), int(split), ptr(dbg_p), ptr(dbg_b), ptr(flags_in), ptr(flags_out),
             int(epoch), stream_ptr(dev))
    elif GENERATION == 2 or split:
        call("nfk_nsf_pairs_fused2", ptr(x), ptr(out), ptr(logdet), ptr(pk["w1"]), ptr(pk["w2"]), ptr(pk["w3"]),
             ptr(pk["b1"]), ptr(pk["b2"]), ptr(pk["b3"]), n, layer._mask[0], float(layer.B), int(bool(inverse)),
             int(accumulate), _ops._arith(layer.arith), int(split), ptr(dbg_p), ptr(dbg_b), stream_ptr(dev))
    else:
        call("nfk_nsf_pairs_fused", ptr(x), ptr(out), ptr(logdet), ptr(pk["w1"]), ptr(pk["w2"]), ptr(pk["w3"]),
             ptr(pk["b1"]), ptr(pk["b2"]), ptr(pk["b3"]), n, layer._mask[0], float(layer.B), int(bool(inverse)),
             int(accumulate), _ops._arith(layer.arith), ptr(dbg_p), ptr(dbg_b), stream_ptr(dev))


def _is_split(layer) -> bool:
    return getattr(layer.psi, "precision", None) == "fp32x3"


def packed(layer):
    if _is_split(layer):
        return packed_split(layer)
    net = layer.psi.network
    l0, l2, l4 = net[0], net[2], net[4]
    key = (_lib.param_epoch(),) + tuple((l.weight._version, l.weight.data_ptr(), l.bias._version)
                                        for l in (l0, l2, l4))
    cache = getattr(layer, "_fused_cache", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    dev = l0.weight.device
    H = l0.out_features
    # IEEE fp16 operands (csrc/nsf_fused.cu): 11-bit significands; weights are O(1), far inside fp16's range
    bf = torch.float16

    def cast(w):
        return w.detach().float().clamp(-65504.0, 65504.0).to(bf)
    w1 = torch.zeros((HP, K1P), dtype=bf, device=dev)
    w1[:H, :32] = cast(l0.weight)
    w2 = torch.zeros((HP, HP), dtype=bf, device=dev)
    w2[:H, :H] = cast(l2.weight)
    w3 = torch.zeros((NF, PC, HP), dtype=bf, device=dev)
    w3[:, :23, :H] = cast(l4.weight).reshape(NF, 23, H)
    w3 = w3.reshape(NF // CF, CF * PC, HP)
    b1 = torch.zeros(HP, dtype=torch.float32, device=dev)
    b1[:H] = l0.bias.detach().float()
    b2 = torch.zeros(HP, dtype=torch.float32, device=dev)
    b2[:H] = l2.bias.detach().float()
    b3 = torch.zeros((NF, PC), dtype=torch.float32, device=dev)
    b3[:, :23] = l4.bias.detach().float().reshape(NF, 23)
    pk = dict(w1=_swizzle_image(w1), w2=_swizzle_image(w2),
              w3=torch.stack([_swizzle_image(w3[c]) for c in range(NF // CF)]).contiguous(),
              b1=b1, b2=b2, b3=b3.reshape(-1).contiguous())
    layer._fused_cache = (key, pk)
    return pk


def run_debug(layer, x, inverse):
    """Test hook: (out, logdet, params [N, 32, 23], bins [N, 32] int8) — the raw spline parameters the fused
    kernel computed for every element (accumulator + b3) and the bin it used.  N must be a multiple of 128."""
    dev = require_cuda(x)
    x = f32c(x)
    n_real = x.shape[0]
    if n_real % ROWS:                       # a partial last tile is zero-padded, as in run()
        xp = torch.zeros(((n_real + ROWS - 1) // ROWS * ROWS, 64), dtype=torch.float32, device=dev)
        xp[:n_real] = x
        x = xp
    N = x.shape[0]
    pk = packed(layer)
    out = torch.empty((N, 64), dtype=torch.float32, device=dev)
    logdet = torch.empty((N,), dtype=torch.float32, device=dev)
    params = torch.empty((N, NF, PC), dtype=torch.float32, device=dev)
    bins = torch.empty((N, NF), dtype=torch.int8, device=dev)
    with torch.cuda.device(dev):
        _launch(layer, pk, x, out, logdet, N, inverse, False, _is_split(layer), params, bins)
    return out[:n_real], logdet[:n_real], params[:n_real, :, :23].contiguous(), bins[:n_real]


def run(layer, x, inverse, logdet=None, flags_in=None, flags_out=None, epoch=1):
    """(out, logdet) of one NSF_CL layer through the fused kernel (whole 128-row tiles; a partial last
    tile is padded).  flags_in / flags_out ([N / 128] int32, N a multiple of 128): per-tile dependency between the
    launches of a chain, see nfk_nsf_pairs_fused2_chain in include/nfk.h."""
    dev = require_cuda(x, logdet)
    x = f32c(x)
    N = x.shape[0]
    pk = packed(layer)
    out = torch.empty((N, 64), dtype=torch.float32, device=dev)
    accumulate = logdet is not None
    if not accumulate:
        logdet = torch.empty((N,), dtype=torch.float32, device=dev)
    n_main = N // ROWS * ROWS
    if n_main:
        with torch.cuda.device(dev):
            tm = _ops.KERNEL_TIMER
            ev = tm.start("nsf_pairs_fused3x" if _is_split(layer) else "nsf_pairs_fused", dev) if tm is not None else None
            _launch(layer, pk, x, out, logdet, n_main, inverse, accumulate, _is_split(layer), None, None, flags_in, flags_out, epoch)
            if ev is not None:
                tm.stop(ev, dev)
    if n_main < N:
        # the last, partial tile goes through the same kernel on a zero-padded copy, so every row of a
        # batch sees the same arithmetic (fp16 operands) whatever the batch size
        nt = N - n_main
        xp = torch.zeros((ROWS, 64), dtype=torch.float32, device=dev)
        xp[:nt] = x[n_main:]
        op = torch.empty((ROWS, 64), dtype=torch.float32, device=dev)
        lp = torch.zeros((ROWS,), dtype=torch.float32, device=dev)
        if accumulate:
            lp[:nt] = logdet[n_main:]
        with torch.cuda.device(dev):
            _launch(layer, pk, xp, op, lp, ROWS, inverse, accumulate, _is_split(layer))
        out[n_main:] = op[:nt]
        logdet[n_main:] = lp[:nt]
    return out, logdet


# ---------------------------------------------------------------------------------------------
# dL/dx of a fused layer in one launch (csrc/nsf_fused_bwd.cu): the HMC force path for hidden <= 128
# ---------------------------------------------------------------------------------------------
# consecutive backward launches of a flow depend on each other tile by tile instead of launch by launch (tools / tests)
TILE_CHAIN = True
# ... and consecutive evaluations of a trajectory too (the first forward launch hangs on the previous evaluation's last
# backward launch, which has advanced the position)
CHAIN_EVALS = True
_DEBUG_KEEP = None               # a list: every buffer of every evaluation stays allocated (no reuse across evaluations)
_DEBUG_NO_FLAGS = False          # tools/ubench/dbg_traj.py: leapfrog fold with launch-by-launch dependencies


def bwd_eligible(layer) -> bool:
    return (_lib.have("nfk_nsf_pairs_fused_bwd") and GENERATION == 2 and eligible(layer)
            and getattr(layer.psi, "precision", None) == "bf16")


# order of a feature's 24 parameter rows in the backward kernel's W3 images: width / height logits interleaved
_BWD_ROW_ORDER = [i // 2 + 8 * (i % 2) for i in range(16)] + list(range(16, 24))


def packed_bwd(layer):
    """Operand images of the one-launch backward (see nfk_nsf_pairs_fused_bwd in include/nfk.h): the forward W3 / b3
    with permuted rows (fp16) and the transposed bf16 dgrad operands W3^T per pair of chunks, W2^T, W1^T."""
    net = layer.psi.network
    l0, l2, l4 = net[0], net[2], net[4]
    key = (_lib.param_epoch(),) + tuple((l.weight._version, l.weight.data_ptr(), l.bias._version) for l in (l0, l2, l4))
    cache = getattr(layer, "_fused_bwd_cache", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    dev = l0.weight.device
    H = l0.out_features
    bf, f16 = torch.bfloat16, torch.float16
    order = torch.tensor(_BWD_ROW_ORDER, device=dev)
    w3p = torch.zeros((NF, PC, HP), dtype=torch.float32, device=dev)
    w3p[:, :23, :H] = l4.weight.detach().float().reshape(NF, 23, H)
    w3p = w3p[:, order, :]                                                          # [32, 24, 128], rows permuted
    w3f = w3p.clamp(-65504.0, 65504.0).to(f16).reshape(NF // CF, CF * PC, HP)
    w3 = torch.stack([_swizzle_image(w3f[c]) for c in range(NF // CF)]).contiguous()
    b3 = torch.zeros((NF, PC), dtype=torch.float32, device=dev)
    b3[:, :23] = l4.bias.detach().float().reshape(NF, 23)
    b3 = b3[:, order].reshape(-1).contiguous()
    w3b = w3p.to(bf)
    w3t = torch.stack([_swizzle_image(w3b[8 * p:8 * p + 8].reshape(8 * PC, HP).t().contiguous())
                       for p in range(NF // 8)]).contiguous()                    # [4, 3, 128, 8, 8]
    w2p = torch.zeros((HP, HP), dtype=bf, device=dev)
    w2p[:H, :H] = l2.weight.detach().to(bf)
    w2t = _swizzle_image(w2p.t().contiguous())                                      # [2, 128, 8, 8]
    w1p = torch.zeros((HP, 32), dtype=bf, device=dev)
    w1p[:H, :] = l0.weight.detach().to(bf)
    w1t = _swizzle_image(w1p.t().contiguous())                                      # [2, 32, 8, 8]
    pk = dict(w3=w3, b3=b3, w3t=w3t, w2t=w2t, w1t=w1t)
    layer._fused_bwd_cache = (key, pk)
    return pk


def layer_backward(layer, x, g_out, g_logdet=None, g_logdet_const=1.0, inverse=False, g_out_scale=1.0, flags_in=None,
                   flags_out=None, keep_padding=False, leapfrog=None, epoch=1, out_epoch=0):
    """dL/dx [N, 64] of one bwd_eligible layer from the layer input x, dL/d(out) = g_out_scale * g_out and dL/d(log_det)
    (a per-row tensor, or a constant for every row).  A partial last tile is zero-padded.  flags_in / flags_out
    ([N / 128] int32): per-tile dependency between the launches of a chain, see nfk_nsf_pairs_fused_bwd in include/nfk.h.
    leapfrog = (momentum, position, kick, drift): momentum += kick * dL/dx; position += drift * momentum, row by row
    inside the same launch (nfk_nsf_pairs_fused_bwd_leapfrog; N a multiple of 128, position may be x)."""
    dev = require_cuda(x, g_out, g_logdet)
    x, g_out = f32c(x), f32c(g_out)
    N = x.shape[0]
    pk, pb = packed(layer), packed_bwd(layer)
    n_pad = (N + ROWS - 1) // ROWS * ROWS
    if n_pad != N:
        xp = torch.zeros((n_pad, 64), dtype=torch.float32, device=dev)
        xp[:N] = x
        gp = torch.zeros((n_pad, 64), dtype=torch.float32, device=dev)
        gp[:N] = g_out
        x, g_out = xp, gp
        if g_logdet is not None:
            glp = torch.zeros((n_pad,), dtype=torch.float32, device=dev)
            glp[:N] = g_logdet
            g_logdet = glp
    g_in = torch.empty((n_pad, 64), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        tm = _ops.KERNEL_TIMER
        ev = tm.start("nsf_pairs_fused_bwd", dev) if tm is not None else None
        if leapfrog is not None:
            lp, lq, kick, drift = leapfrog
            if n_pad != N or lp.shape != (N, 64) or lq.shape != (N, 64) or not (lp.is_contiguous() and lq.is_contiguous()):
                raise ValueError("leapfrog fold needs whole 128-row tiles and contiguous [N, 64] momentum / position")
            call("nfk_nsf_pairs_fused_bwd_leapfrog", ptr(x), ptr(g_out), float(g_out_scale),
                 ptr(f32c(g_logdet)) if g_logdet is not None else ptr(None), float(g_logdet_const), ptr(g_in), ptr(pk["w1"]),
                 ptr(pk["w2"]), ptr(pb["w3"]), ptr(pb["w3t"]), ptr(pb["w2t"]), ptr(pb["w1t"]), ptr(pk["b1"]), ptr(pk["b2"]),
                 ptr(pb["b3"]), n_pad, layer._mask[0], float(layer.B), int(bool(inverse)), ptr(flags_in), ptr(flags_out),
                 int(epoch), int(out_epoch), ptr(lp), ptr(lq), float(kick), float(drift), stream_ptr(dev))
        else:
            call("nfk_nsf_pairs_fused_bwd", ptr(x), ptr(g_out), float(g_out_scale),
                 ptr(f32c(g_logdet)) if g_logdet is not None else ptr(None), float(g_logdet_const), ptr(g_in), ptr(pk["w1"]),
                 ptr(pk["w2"]), ptr(pb["w3"]), ptr(pb["w3t"]), ptr(pb["w2t"]), ptr(pb["w1t"]), ptr(pk["b1"]), ptr(pk["b2"]),
                 ptr(pb["b3"]), n_pad, layer._mask[0], float(layer.B), int(bool(inverse)), ptr(flags_in), ptr(flags_out),
                 int(epoch), stream_ptr(dev))
        if ev is not None:
            tm.stop(ev, dev)
    return g_in if keep_padding else g_in[:N]


def flow_eligible(model) -> bool:
    """Every layer is a bwd_eligible NSF_CL on the fused path and the prior is an isotropic zero-mean Gaussian."""
    from .flows import NSF_CL
    var = model._prior_var() if hasattr(model, "_prior_var") else None
    return (var is not None and len(model.flows) > 0
            and all(isinstance(f, NSF_CL) and f.fused and bwd_eligible(f) for f in model.flows))


def tile_chain_ok(model, n_rows, dev=None) -> bool:
    """Tile-flag chains pay when a launch has more tiles than the GPU has SMs (the chain turns the last, partial wave of
    every launch into the first wave of the next); with a single partial wave the flag traffic is a net loss
    (8,192 chains: 0.405 ms per evaluation without, 0.411 with)."""
    if not (TILE_CHAIN and n_rows % ROWS == 0 and len(model.flows) > 1 and GENERATION == 2
            and _lib.have("nfk_nsf_pairs_fused2_chain")):
        return False
    if TILE_CHAIN == "always":                      # tests: exercise the chains at any size
        return True
    if dev is None:
        dev = next(model.parameters()).device
    return n_rows // ROWS > _sm_count(dev)


_SM_COUNT = {}


def _sm_count(dev):
    key = str(dev)
    if key not in _SM_COUNT:
        _SM_COUNT[key] = torch.cuda.get_device_properties(dev).multi_processor_count
    return _SM_COUNT[key]


def _tile_flags(model, dev, n_tiles, L):
    """[3, L, n_tiles] int32: forward-chain flags, backward-chain flags, and (row 0 of the third plane) the flags that
    hang the next evaluation's first forward launch on this evaluation's last backward launch."""
    key = (dev, n_tiles, L)
    cache = getattr(model, "_fused_tile_flags", None)
    if cache is None or cache[0] != key:
        cache = model._fused_tile_flags = (key, torch.zeros((3, L, n_tiles), dtype=torch.int32, device=dev))
    return cache[1]


def flow_logp_and_grad(model, x, need_logp=True, leapfrog=None, hang_on_previous=False, next_hangs_on_this=False,
                       zero_flags=True, epoch=1):
    """log p(x) [N] and d log p / dx [N, d] through a flow whose layers are all bwd_eligible, under an isotropic Gaussian
    prior: one forward launch per layer (keeping each layer's input), the log-prob reduction, one backward launch per
    layer; None when the model does not qualify.
    leapfrog = (momentum, position, kick, drift): the last backward launch also advances momentum and position
    (layer_backward).  hang_on_previous / next_hangs_on_this: consecutive evaluations of a trajectory form one tile-flag
    chain (the first forward launch of an evaluation takes tile t as soon as the previous evaluation's last backward
    launch has advanced the position of tile t); such evaluations share the flag arrays without a fill kernel between
    them (zero_flags = False) and therefore count epochs: flags only grow, evaluation number e of a trajectory waits
    for and stores e + 1, so a launch that is resident early cannot take an earlier evaluation's flag for its own."""
    if not flow_eligible(model):
        return None
    var = model._prior_var()
    h = f32c(x.detach())
    N, L = h.shape[0], len(model.flows)
    for f in model.flows:          # every operand image exists BEFORE the first launch of the chain: a flagged launch does
        packed(f)                  # not wait for the stream, so nothing but flagged tiles may be produced inside the chain
        packed_bwd(f)
    fw = bw = link = None
    if tile_chain_ok(model, N, h.device) and not _DEBUG_NO_FLAGS:
        # per-tile dependency between consecutive launches of the forward chain and of the backward chain (rows are
        # independent): a launch starts on the SMs the previous one has already left instead of waiting for its last wave
        flags = _tile_flags(model, h.device, N // ROWS, L)
        if zero_flags:
            flags.zero_()
        fw, bw, link = flags[0], flags[1], flags[2][0]
    elif hang_on_previous or next_hangs_on_this:
        raise ValueError("chained evaluations need the tile-flag path (whole 128-row tiles, TILE_CHAIN)")
    # without the log-prob reduction between them (need_logp = False: a leapfrog step inside a trajectory only needs the
    # force) the first backward launch hangs on the last forward launch tile by tile as well
    join = fw is not None and not need_logp
    inputs = []
    logdet = None                  # the first layer writes its log-det, the others accumulate: no zero-fill launch
    for i, f in enumerate(model.flows):
        inputs.append(h)
        f_in = (fw[i - 1] if i > 0 else (link if hang_on_previous else None)) if fw is not None else None
        h, logdet = run(f, h, False, logdet, f_in, fw[i] if (fw is not None and (i + 1 < L or join)) else None, epoch)
    logp = _ops.gauss_logprob(h, var, add=logdet, add_sign=1.0) if need_logp else None
    g, scale = h, -1.0 / var                                  # d log N(z; 0, var I) / dz = -z / var, applied by the first launch
    keep = []                      # in a chain the launches overlap: no buffer of the evaluation is handed back to the
    #                                allocator (and possibly re-issued at another offset) before the last launch is queued
    for i, (f, xin) in enumerate(zip(reversed(model.flows), reversed(inputs))):
        f_in = (bw[i - 1] if i > 0 else (fw[L - 1] if join else None)) if bw is not None else None
        f_out = (bw[i] if i + 1 < L else (link if next_hangs_on_this else None)) if bw is not None else None
        keep.append(g)
        g = layer_backward(f, xin, g, None, 1.0, False, scale, f_in, f_out, leapfrog=leapfrog if i + 1 == L else None,
                           epoch=epoch, out_epoch=(epoch + 1 if (i + 1 == L and next_hangs_on_this) else 0))
        scale = 1.0
    if _DEBUG_KEEP is not None:
        _DEBUG_KEEP.append((inputs, keep, logdet, h, g))
    del keep
    return logp, g
