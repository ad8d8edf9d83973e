"""Free-energy estimators on the device: consumers of ``sample`` / ``evaluate``
outputs that already live on the GPU.

``BAR`` follows applications/src/bar.py:61-67 (Bennett acceptance ratio, self-consistent
iteration with the max-shifted log-sums of ``BARzero``, bar.py:16-59): same arguments, returns
DeltaF.  ``log_mean_exp`` is the reweighting estimate of applications/src/test.py:66-68 and the
``logsumexp - log(npoints)`` of ``integrate_out_v`` (dynamics.py:36)."""
from __future__ import annotations

import torch

from ._lib import call, ptr, require_cuda, stream_ptr


def BAR(w_F, w_R, DeltaF=0.0, maximum_iterations=1000, relative_tolerance=1.0e-5, return_iterations=False):
    dev = require_cuda(w_F, w_R)
    f64 = w_F.dtype == torch.float64 or w_R.dtype == torch.float64
    dt = torch.float64 if f64 else torch.float32
    wf = w_F.detach().to(dt).contiguous().reshape(-1)
    wr = w_R.detach().to(dt).contiguous().reshape(-1)
    out = torch.empty(2, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        call("nfk_bar", ptr(wf), ptr(wr), wf.numel(), wr.numel(), int(f64), float(DeltaF), int(maximum_iterations),
             float(relative_tolerance), ptr(out), stream_ptr(dev))
    res = out.cpu()
    return (float(res[0]), int(res[1])) if return_iterations else float(res[0])


def log_mean_exp(a, dim=0):
    """log(mean(exp(a), dim)) without overflow; ``a`` fp32 CUDA tensor, any shape."""
    dev = require_cuda(a)
    a = a.detach().float()
    moved = a.movedim(dim, 0).contiguous()
    rows = moved.shape[0]
    cols = moved.numel() // max(rows, 1)
    out = torch.empty(moved.shape[1:], dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("nfk_log_mean_exp", ptr(moved), ptr(out), rows, cols, stream_ptr(dev))
    return out
