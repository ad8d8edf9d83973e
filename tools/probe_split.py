"""Error budget of the split-operand fused kernel against the golden fixtures: parameters, z, log_det per
fixture / direction / arithmetic (max and quantiles of |a-b| / max(1,|b|)).  python tools/probe_split.py"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from normalizingflow_b200 import _fused, flows                                   # noqa: E402
from tests.helpers import T, golden, parse_masks, rel_vec, sub_sd               # noqa: E402

for name in ("nsfcl_d64.npz", "nsfcl_d64_stress.npz"):
    g = golden(name)
    size, dim, K, B, H = int(g["size"]), int(g["dim"]), int(g["K"]), float(g["B"]), int(g["H"])
    for prec in ("fp32x3", "bf16", "fp32"):
        for arith in ("hybrid", "fast"):
            for mi, mask in enumerate(parse_masks(g)):
                p = f"m{mi}."
                lay = flows.NSF_CL(size, dim=dim, K=K, B=B, hidden_dim=H, mask=mask, arith=arith)
                lay.load_state_dict(sub_sd(g, p + "sd."))
                lay.psi.precision = prec
                lay = lay.cuda()
                for inv, xk, pk, ok, lk in ((False, "x", "params", "z", "ld"), (True, "zin", "params_inv", "x_inv", "ld_inv")):
                    x = T(g[p + xk]).cuda()
                    with torch.no_grad():
                        if prec == "fp32":
                            out, ld = (lay.inverse if inv else lay.forward)(x)
                            ep = float("nan")
                        else:
                            out, ld, params, bins = _fused.run_debug(lay, x, inv)
                            ep = float(rel_vec(params, g[p + pk]).max())
                    ez, el = rel_vec(out, g[p + ok]), rel_vec(ld, g[p + lk])
                    print(f"{name:22s} {prec:7s} {arith:6s} mask {mask} inv {int(inv)}: params {ep:.2e}  z max {float(ez.max()):.2e} "
                          f"q99 {float(ez.flatten().quantile(0.99)):.2e}  log_det max {float(el.max()):.2e} med {float(el.median()):.2e}")
