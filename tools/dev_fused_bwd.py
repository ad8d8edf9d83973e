"""Development check of the one-launch layer backward (csrc/nsf_fused_bwd.cu): dL/dx against autograd through the
unfused kernels of the same layer and against the wide (multi-launch) tensor-core backward, then the time of a
whole log-prob + gradient evaluation at 65,536 chains.  Run on a GPU box under `timeout`."""
import sys
import time

import torch

sys.path.insert(0, ".")
from normalizingflow_b200 import _fused, _wide, flows, models  # noqa: E402


def one_layer(mask, inverse, H, N):
    torch.manual_seed(7)
    lay = flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=H, mask=[mask]).cuda()
    with torch.no_grad():
        lay.psi.network[4].weight.mul_(3.0)
    lay.psi.precision = "bf16"
    gen = torch.Generator().manual_seed(11)
    x = (1.2 * torch.randn(N, 64, generator=gen)).cuda()
    r = torch.randn(N, 64, generator=gen).cuda()
    s = torch.randn(N, generator=gen).cuda()
    # autograd through the fp32 (parity) conditioner + stand-alone spline kernels
    lay.fused = False
    lay.psi.precision = "fp32"
    xg = x.clone().requires_grad_()
    out, ld = (lay.inverse if inverse else lay.forward)(xg)
    (gref,) = torch.autograd.grad((out * r).sum() + (ld * s).sum(), [xg])
    lay.psi.precision = "bf16"
    lay.fused = True
    assert _fused.bwd_eligible(lay), "layer not eligible"
    with torch.no_grad():
        gin = _fused.layer_backward(lay, x, r, s, 1.0, inverse)
        torch.cuda.synchronize()
        o2, l2, ctx = _wide.layer_forward_saving(lay, x, inverse)
        gw = _wide.layer_backward(lay, ctx, r, s)
    torch.cuda.synchronize()
    err = (gin - gref).abs() / (1.0 + gref.abs())
    errw = (gw - gref).abs() / (1.0 + gref.abs())
    print(f"mask {mask} inv {int(inverse)} H {H} N {N}: fused-bwd vs fp32 autograd median {float(err.median()):.2e} "
          f"p99 {float(err.flatten().kthvalue(int(0.99 * err.numel())).values):.2e} max {float(err.max()):.2e} "
          f"frac>3e-2 {float((err > 3e-2).float().mean()):.2e} | wide: median {float(errw.median()):.2e} "
          f"max {float(errw.max()):.2e} frac>3e-2 {float((errw > 3e-2).float().mean()):.2e} | nan {int(torch.isnan(gin).sum())}",
          flush=True)
    # constant d logdet, no tensor
    with torch.no_grad():
        g1 = _fused.layer_backward(lay, x, r, None, 1.0, inverse)
        g2 = _fused.layer_backward(lay, x, r, torch.ones_like(s), 0.0, inverse)
    print("   const-vs-tensor logdet gradient equal:", bool(torch.equal(g1, g2)), flush=True)


def whole_flow(C=65536, H=128):
    torch.manual_seed(0)
    dev = torch.device("cuda")
    fl = [flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=H, mask=[i % 2]) for i in range(8)]
    m = models.NormalizingFlowModel(models.GaussianPrior(64, device=dev), fl, device=dev).to(dev)
    for f in fl:
        f.psi.precision = "bf16"
    x = torch.randn(C, 64, device=dev)
    with torch.no_grad():
        a = _fused.flow_logp_and_grad(m, x)
        b = _wide.flow_logp_and_grad(m, x)
    assert a is not None and b is not None
    torch.cuda.synchronize()
    _fused.TILE_CHAIN = False
    with torch.no_grad():
        a2 = _fused.flow_logp_and_grad(m, x)
    torch.cuda.synchronize()
    print("tile-chained launches == launch-by-launch dependency, bitwise:", bool(torch.equal(a[1], a2[1])), flush=True)
    for chain in (False, True):
        _fused.TILE_CHAIN = chain
        with torch.no_grad():
            for _ in range(3):
                _fused.flow_logp_and_grad(m, x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                r = _fused.flow_logp_and_grad(m, x)
            e1.record()
            torch.cuda.synchronize()
        print(f"   TILE_CHAIN={chain}: {e0.elapsed_time(e1) / 20:.3f} ms per evaluation; equal to first: {bool(torch.equal(r[1], a[1]))}", flush=True)
    _fused.TILE_CHAIN = True
    sc = float(b[1].abs().max())
    print(f"flow: logp fused vs wide max abs {float((a[0] - b[0]).abs().max()):.2e}; force max abs diff "
          f"{float((a[1] - b[1]).abs().max()):.2e} (scale {sc:.2e}), median {float((a[1] - b[1]).abs().median()):.2e}", flush=True)
    for name, fn in (("fused (17 launches)", _fused.flow_logp_and_grad), ("wide", _wide.flow_logp_and_grad)):
        with torch.no_grad():
            for _ in range(3):
                fn(m, x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn(m, x)
            e1.record()
            torch.cuda.synchronize()
        print(f"   {name}: {e0.elapsed_time(e1) / 20:.3f} ms per log-prob + gradient at {C} chains", flush=True)
    # backward launch alone
    lay = fl[0]
    g = torch.randn(C, 64, device=dev)
    with torch.no_grad():
        for _ in range(3):
            _fused.layer_backward(lay, x, g, None, 1.0, False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            _fused.layer_backward(lay, x, g, None, 1.0, False)
        e1.record()
        torch.cuda.synchronize()
        print(f"   backward launch: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us at {C} rows", flush=True)
        for _ in range(3):
            _fused.run(lay, x, False)
        e0.record()
        for _ in range(50):
            _fused.run(lay, x, False)
        e1.record()
        torch.cuda.synchronize()
        print(f"   forward launch: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us at {C} rows", flush=True)


if __name__ == "__main__":
    t0 = time.time()
    for H in (128, 64):
        for mask in (0, 1):
            for inverse in (False, True):
                one_layer(mask, inverse, H, 1024 if H == 128 else 900)
    whole_flow()
    print(f"done in {time.time() - t0:.1f} s")
