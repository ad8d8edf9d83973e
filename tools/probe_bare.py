"""Prints the BARE parity figures (max |a-b| / max(1,|b|), no noise allowance) of the fp32 paths against the
golden fixtures: stand-alone spline kernels per arithmetic, whole fp32 layers, so that the tests can assert
the bare north-star 1e-5 wherever it holds.  python tools/probe_bare.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from normalizingflow_b200 import _ops, flows                              # noqa: E402
from tests.helpers import T, golden, parse_masks, rel_err, sub_sd          # noqa: E402

for name in ("nsfcl_d64.npz", "nsfcl_d64_stress.npz", "nsfcl_lj38.npz", "nsfcl_k32.npz"):
    g = golden(name)
    size, dim, K, B, H = int(g["size"]), int(g["dim"]), int(g["K"]), float(g["B"]), int(g["H"])
    for arith in ("exact", "hybrid", "fast"):
        ez = el = 0.0
        for mi, mask in enumerate(parse_masks(g)):
            p = f"m{mi}."
            for inv, xk, pk, ok, lk in ((False, "x", "params", "z", "ld"), (True, "zin", "params_inv", "x_inv", "ld_inv")):
                out, ld, _ = _ops.rqs_coupling(T(g[p + xk]).cuda(), T(g[p + pk]).cuda(), size, dim, mask, K, B, inv, arith)
                ez, el = max(ez, rel_err(out, g[p + ok])), max(el, rel_err(ld, g[p + lk]))
        print(f"transform {name:22s} {arith:6s} z {ez:.2e} log_det {el:.2e}")
    ez = el = 0.0
    for mi, mask in enumerate(parse_masks(g)):
        p = f"m{mi}."
        layer = flows.NSF_CL(size, dim=dim, K=K, B=B, hidden_dim=H, mask=mask)
        layer.load_state_dict(sub_sd(g, p + "sd."))
        layer = layer.cuda()
        with torch.no_grad():
            z, ld = layer.forward(T(g[p + "x"]).cuda())
            xi, ldi = layer.inverse(T(g[p + "zin"]).cuda())
        ez = max(ez, rel_err(z, g[p + "z"]), rel_err(xi, g[p + "x_inv"]))
        el = max(el, rel_err(ld, g[p + "ld"]), rel_err(ldi, g[p + "ld_inv"]))
    print(f"fp32 layer {name:22s} hybrid z {ez:.2e} log_det {el:.2e}  (H={H})")
