"""ctypes binding of libnfk.so — the C-ABI declared in include/nfk.h.

There is no CPU path and no other backend: if the library is missing this module raises at
import, and every op raises if it is handed a tensor that is not a CUDA tensor.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NFK_LIB") or os.path.join(_HERE, "libnfk.so")     # NFK_LIB: timing experiments (tools/ubench)

NFK_OK, NFK_EINVAL, NFK_ECUDA, NFK_EUNSUPPORTED = 0, 1, 2, 3
ARITH_EXACT, ARITH_HYBRID, ARITH_FAST = 0, 1, 2
ARITH = {"exact": ARITH_EXACT, "hybrid": ARITH_HYBRID, "fast": ARITH_FAST}
IMG_BF16, IMG_F16 = 0, 1          # NFK_IMG_*: element format of the wide path's operand images

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C normalizingflow_b200/csrc` (needs nvcc; sm_100a only, there is no fallback)")

lib = ctypes.CDLL(LIB_PATH)

_P = c_void_p
# name -> (restype, argtypes); mirrors include/nfk.h one to one
SIGNATURES = {
    "nfk_abi_version": (c_int, []),
    "nfk_last_error": (c_char_p, []),
    "nfk_launch_count": (c_int64, []),
    "nfk_set_scan_order": (c_int, [c_int]),
    "nfk_get_scan_order": (c_int, []),
    "nfk_set_tuning": (c_int, [c_int, c_int, c_int, c_int]),
    "nfk_debug_knots": (c_int, [_P, _P, c_int64, c_int, c_float, c_int, c_int, _P]),
    "nfk_rqs_coupling": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, _P, c_int, c_int, c_float,
                                 c_int, c_int, c_int, _P]),
    "nfk_unconstrained_rqs": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_float, c_int, c_int, _P]),
    "nfk_rqs_elementwise": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_float, c_int, c_int, _P]),
    "nfk_rqs_elementwise_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_float, c_int, _P]),
    "nfk_affine_halfcoupling": (c_int, [_P, c_int64, c_int, _P, _P, _P, c_int64, c_int, _P, c_int64, c_int,
                                        c_int, c_int, _P]),
    "nfk_affine_halfcoupling_bwd": (c_int, [_P, c_int64, c_int, _P, _P, _P, c_int64, c_int, _P, _P, c_int64,
                                            c_int, _P, _P, c_int64, c_int, c_int, _P]),
    "nfk_planar_prepare": (c_int, [_P, _P, _P, _P, c_int, c_int, _P]),
    "nfk_planar_stack": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "nfk_planar_gram": (c_int, [_P, _P, _P, c_int, c_int, _P]),
    "nfk_planar_stack_mma": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "nfk_planar_stack_bwd": (c_int, [_P] * 12 + [c_int64, c_int, c_int, _P]),
    "nfk_rqs_coupling_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, _P, c_int, c_int, c_float, c_int, _P]),
    "nfk_radial_sumsq": (c_int, [_P, _P, _P, c_int64, c_int, _P]),
    "nfk_radial": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "nfk_radial_stack": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "nfk_radial_global": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "nfk_radial_dot": (c_int, [_P, _P, _P, _P, c_int64, c_int, _P]),
    "nfk_radial_bwd": (c_int, [_P] * 12 + [c_int64, c_int, c_int, _P]),
    "nfk_gauss_logprob": (c_int, [_P, _P, c_float, _P, c_int64, c_int, c_float, _P]),
    "nfk_linear_f32": (c_int, [_P, c_int64, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "nfk_linear_tf32x3": (c_int, [_P, c_int64, _P, c_int64, _P, _P, c_int64, c_int64, c_int, c_int, c_int, _P]),
    "nfk_linear_bf16": (c_int, [_P, c_int64, _P, c_int64, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, _P]),
    "nfk_nsf_fused_rows_per_tile": (c_int, []),
    "nfk_set_fused_trace": (c_int, [_P]),
    "nfk_nsf_pairs_fused": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_float, c_int, c_int,
                                    c_int, _P, _P, _P]),
    "nfk_set_fused2_pdl": (c_int, [c_int]),
    "nfk_nsf_pairs_fused2": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_float, c_int, c_int,
                                     c_int, c_int, _P, _P, _P]),
    "nfk_nsf_pairs_fused2_chain": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_float, c_int, c_int,
                                           c_int, c_int, _P, _P, _P, _P, c_int, _P]),
    "nfk_nsf_pairs_fused_bwd_leapfrog": (c_int, [_P, _P, c_float, _P, c_float, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64,
                                                 c_int, c_float, c_int, _P, _P, c_int, c_int, _P, _P, c_float, c_float, _P]),
    "nfk_nsf_pairs_fused_bwd": (c_int, [_P, _P, c_float, _P, c_float, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int,
                                        c_float, c_int, _P, _P, c_int, _P]),
    "nfk_gemm_f32": (c_int, [_P, c_int64, c_int, _P, c_int64, c_int, _P, c_int64, c_int64, c_int64, c_int64,
                             c_int, _P]),
    "nfk_gemm_ws_rows_per_tile": (c_int, []),
    "nfk_set_gemm_ws_pair_mode": (c_int, [c_int]),
    "nfk_gemm_ws_last_clusters": (c_int, []),
    "nfk_gemm_ws": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int64, _P, c_int,
                            _P]),
    "nfk_gemm_ws_rqs_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_float, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P,
                                    c_int, c_float, c_int, _P]),
    "nfk_wgrad_ws": (c_int, [_P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, _P]),
    "nfk_pack_w_img": (c_int, [_P, c_int64, c_int, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "nfk_unpack_img_rows": (c_int, [_P, _P, c_int64, c_int, c_int, c_int64, _P]),
    "nfk_scatter_add_cols": (c_int, [_P, _P, c_int64, c_int, c_int, _P, c_int, _P]),
    "nfk_gemm_ws_group_bytes": (c_int, []),
    "nfk_gemm_ws_grouped": (c_int, [_P, c_int, c_int64, _P, c_int, c_int, c_int, c_int, c_int64, c_int, _P]),
    "nfk_nsf_ar_pack": (c_int, [_P, _P, c_int64, c_int, c_float, c_int, _P]),
    "nfk_gemm_ws_rqs": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P, c_int, c_float, c_int,
                                c_int, c_int, c_int, _P, _P, _P]),
    "nfk_pack_a_img": (c_int, [_P, _P, c_int64, c_int, c_int, _P, c_int, c_int, c_int, _P]),
    "nfk_gather_cols": (c_int, [_P, _P, c_int64, c_int, c_int, _P, c_int, c_int, c_int64, _P]),
    "nfk_cast_f32_bf16": (c_int, [_P, _P, c_int64, _P]),
    "nfk_einstein_logprob": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_float, c_float, _P]),
    "nfk_lj_potential": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_float, c_float, c_float, c_float, c_int, _P]),
    "nfk_gmm_logprob": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "nfk_bar": (c_int, [_P, _P, c_int64, c_int64, c_int, ctypes.c_double, c_int, ctypes.c_double, _P, _P]),
    "nfk_log_mean_exp": (c_int, [_P, _P, c_int64, c_int64, _P]),
    "nfk_leapfrog_kick_drift": (c_int, [_P, _P, _P, c_int64, c_float, c_float, _P]),
    "nfk_leapfrog_kick": (c_int, [_P, _P, c_int64, c_float, _P]),
}


def _bind():
    missing = []
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    return missing


MISSING = _bind()


def have(name: str) -> bool:
    return name in SIGNATURES and name not in MISSING


def last_error() -> str:
    return (lib.nfk_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    """0 -> None; NFK_EINVAL -> ValueError (as nf/utils.py:64-71 raise); else RuntimeError."""
    if rc == NFK_OK:
        return
    msg = f"{what}: {last_error()}" if what else last_error()
    if rc == NFK_EINVAL:
        raise ValueError(msg)
    raise RuntimeError(msg)


def call(name: str, *args) -> None:
    if name in MISSING:
        raise RuntimeError(f"libnfk.so does not export {name}; rebuild the library")
    check(getattr(lib, name)(*args), name)


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "normalizingflow_b200 has no CPU path: tensors must live on a CUDA (sm_100a) device; "
                f"got a {t.device} tensor")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} and {t.device}")
    return dev


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32, contiguous, 16-byte aligned view/copy of ``t``."""
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def ptr(t) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())


def stream_ptr(dev: torch.device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def i32_array(values):
    arr = (c_int32 * len(values))(*[int(v) for v in values])
    return arr


def launch_count() -> int:
    return int(lib.nfk_launch_count())


# Weight-image caches (fused / wide / bf16 / NSF_AR packs, host chunk graphs, HMC trajectory graphs) are
# keyed on Parameter._version and data_ptr.  Writes that bypass the version counter -- a replayed CUDA
# graph that contains the optimizer step, ``p.data`` writes such as dist.broadcast -- must bump this
# epoch, which is part of every cache key.
_PARAM_EPOCH = 0


def param_epoch() -> int:
    return _PARAM_EPOCH


def invalidate_caches() -> None:
    """Drop every cached weight image / captured graph on next use (call after writing parameters in a
    way autograd's version counter does not see)."""
    global _PARAM_EPOCH
    _PARAM_EPOCH += 1
