#!/usr/bin/env python
"""GPU probe: how far apart are (a) the reference chain run by ATen on CUDA vs on the CPU and
(b) each libnfk arithmetic mode vs both, on the parity-test inputs?  Used to set the fp32 gate."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200 import _ops
from oracle import nf_oracle as O
from tests.helpers import T, golden, parse_masks, rel_vec

def report(tag, mine, cpu, cuda, f64):
    e_dev = rel_vec(cuda, cpu); e_mc = rel_vec(mine, cpu); e_mg = rel_vec(mine, cuda); e_ref = rel_vec(cpu, f64); e_m64 = rel_vec(mine, f64)
    well = e_dev <= 2.5e-6
    print(f"{tag:55s} dev {e_dev.max():.1e} | mine-cpu {e_mc.max():.1e} mine-cuda {e_mg.max():.1e} | cpu-64 {e_ref.max():.1e} mine-64 {e_m64.max():.1e}"
          f" | frac>1e-5: dev {(e_dev>1e-5).double().mean():.1e} mine-cpu {(e_mc>1e-5).double().mean():.1e} mine-cuda {(e_mg>1e-5).double().mean():.1e}"
          f" | well-cond max mine-cpu {e_mc[well].max():.1e} mine-cuda {e_mg[well].max():.1e}")

def main():
    for name in ["nsfcl_d64.npz", "nsfcl_d64_stress.npz", "nsfcl_lj38.npz", "nsfcl_k32.npz"]:
        g = golden(name); size, dim, K, B = int(g["size"]), int(g["dim"]), int(g["K"]), float(g["B"])
        for mi, mask in enumerate(parse_masks(g)):
            p = f"m{mi}."
            for inv, xk, pk in ((False, "x", "params"), (True, "zin", "params_inv")):
                x, pr = T(g[p + xk]), T(g[p + pk])
                cpu = O.nsf_cl_transform(x, pr, size, dim, mask, K, B, inv)
                f64 = O.nsf_cl_transform(x.double(), pr.double(), size, dim, mask, K, B, inv)
                cu = O.nsf_cl_transform(x.cuda(), pr.cuda(), size, dim, mask, K, B, inv)
                for mode in ("exact", "hybrid", "fast"):
                    out, ld, _ = _ops.rqs_coupling(x.cuda(), pr.cuda(), size, dim, mask, K, B, inv, mode)
                    report(f"{name[:-4]} m{mask} inv{int(inv)} {mode} z", out, cpu[0], cu[0], f64[0])
                    report(f"{name[:-4]} m{mask} inv{int(inv)} {mode} ld", ld, cpu[1], cu[1], f64[1])
    gen = torch.Generator().manual_seed(77); N = 65536
    x = torch.randn(N, 64, generator=gen) * 1.5; pr = torch.randn(N, 32, 23, generator=gen) * 1.5
    for inv in (False, True):
        cpu = O.nsf_cl_transform(x, pr, 32, 2, [1], 8, 3.0, inv)
        f64 = O.nsf_cl_transform(x.double(), pr.double(), 32, 2, [1], 8, 3.0, inv)
        cu = O.nsf_cl_transform(x.cuda(), pr.cuda(), 32, 2, [1], 8, 3.0, inv)
        for mode in ("exact", "hybrid", "fast"):
            out, ld, bins = _ops.rqs_coupling(x.cuda(), pr.cuda(), 32, 2, [1], 8, 3.0, inv, mode, want_bins=True)
            report(f"large inv{int(inv)} {mode} z", out, cpu[0], cu[0], f64[0])
            report(f"large inv{int(inv)} {mode} ld", ld, cpu[1], cu[1], f64[1])
            print("   bins != cuda:", int((bins.long() != cu[2]).sum()), " != cpu:", int((bins.long().cpu() != cpu[2]).sum()))

if __name__ == "__main__":
    main()
