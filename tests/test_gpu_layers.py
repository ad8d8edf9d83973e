"""GPU parity tests of the conditioner GEMMs, RealNVP, Planar, Radial and the model-level
interface against the golden fixtures produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from tests.helpers import RTOL_BF16, RTOL_FP32, T, golden, parse_masks, rel_err, sub_sd

pytestmark = pytest.mark.gpu


def _mods():
    from normalizingflow_b200 import _bf16, _ops, flows, models
    return _ops, _bf16, flows, models


@pytest.mark.parametrize("M,K,N,act", [(1, 1, 1, 0), (7, 32, 24, 1), (300, 33, 130, 1), (1000, 128, 736, 0),
                                       (257, 800, 100, 1)])
def test_linear_f32_matches_torch(M, K, N, act):
    ops, _, _, _ = _mods()
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    if act:
        ref = torch.tanh(ref)
    y = ops.linear_f32(x.cuda(), w.cuda(), b.cuda(), act)
    assert rel_err(y, ref) <= RTOL_FP32


@pytest.mark.parametrize("M,K,N,act,f32", [(128, 64, 128, 1, False), (1, 8, 16, 0, True), (300, 32, 128, 1, False),
                                           (1000, 128, 736, 0, True), (513, 800, 800, 1, False),
                                           (129, 104, 100, 1, False), (4096, 128, 128, 1, False)])
def test_linear_bf16_tcgen05_matches_torch(M, K, N, act, f32):
    """tcgen05 GEMM vs an fp64 matmul of the SAME bf16-rounded operands: only accumulation
    order, MUFU.TANH and the bf16 output rounding may differ."""
    _, bf, _, _ = _mods()
    g = torch.Generator().manual_seed(M + K + N)
    x = (torch.randn(M, K, generator=g)).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, generator=g)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    if act:
        ref = torch.tanh(ref)
    y = bf.linear_bf16(x.cuda(), w.cuda(), b.cuda(), act, f32)
    torch.cuda.synchronize()
    y = y[:, :N].float()
    tol = 2e-3 if (f32 and not act) else 1e-2
    assert rel_err(y, ref) <= tol, rel_err(y, ref)


@pytest.mark.parametrize("M,K,N,act", [(512, 32, 128, 1), (4099, 128, 736, 0), (1000, 128, 128, 1), (777, 800, 800, 1),
                                        (2048, 36, 50, 0), (640, 4, 1, 1), (130000, 32, 128, 1)])
def test_linear_tf32x3_is_fp32_accurate(M, K, N, act):
    """csrc/linear_tf32.cu: 3xTF32 on tcgen05 (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) against an
    fp64 product — fp32-class error, i.e. no worse than the CUDA-core fp32 kernel it replaces — on ragged
    row counts, K blocks with a zero-filled tail, single-column outputs and several N tiles."""
    ops = _mods()[0]
    from normalizingflow_b200 import _lib
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    assert ops.TF32X3
    before = _lib.launch_count()
    y = ops.linear_f32(x, w, b, act)
    assert _lib.launch_count() == before + 1
    ref = x.double() @ w.double().t() + b.double()
    if act:
        ref = torch.tanh(ref)
    err = float((y.double() - ref).abs().max() / ref.abs().max().clamp_min(1.0))
    ops.TF32X3 = False
    try:
        y_cc = ops.linear_f32(x, w, b, act)          # the CUDA-core fp32 kernel on the same inputs
    finally:
        ops.TF32X3 = True
    err_cc = float((y_cc.double() - ref).abs().max() / ref.abs().max().clamp_min(1.0))
    # measured on B200: 0.5-6.5e-6 of the output scale, 1.2-4x the CUDA-core kernel's own error on the same
    # inputs (the tensor core's fp32 accumulation truncates when it aligns addends, which is why the k-steps
    # are spread over four accumulators) — inside the 1e-5 parity class
    print(f"tf32x3 {M}x{K}x{N} act={act}: max err {err:.2e} (CUDA-core fp32 kernel {err_cc:.2e})")
    assert err <= max(5e-6, 2.5 * err_cc), (err, err_cc)


def test_fcnn_fp32_and_bf16_vs_reference():
    ops, bf, flows, _ = _mods()
    g = golden("nsfcl_d64.npz")
    sd = sub_sd(g, "m1.sd.psi.")
    net = flows.FCNN(32, 23 * 32, 24)
    net.load_state_dict(sd)
    net = net.cuda()
    x = T(g["m1.x"]).reshape(-1, 32, 2)[:, :, 1].contiguous()
    ref = T(g["m1.params"]).reshape(x.shape[0], -1)
    with torch.no_grad():
        y32 = net(x.cuda())
        net.precision = "bf16"
        y16 = net(x.cuda())
    assert rel_err(y32, ref) <= RTOL_FP32
    assert rel_err(y16, ref) <= RTOL_BF16


def test_realnvp_matches_reference():
    _, _, flows, _ = _mods()
    g = golden("realnvp.npz")
    for tag in ("d2", "d64", "d6"):
        d, H = int(g[tag + ".d"]), int(g[tag + ".H"])
        layer = flows.RealNVP(d, hidden_dim=H)
        layer.load_state_dict(sub_sd(g, tag + ".sd."))
        layer = layer.cuda()
        with torch.no_grad():
            z, ld = layer.forward(T(g[tag + ".x"]).cuda())
            x, ldi = layer.inverse(T(g[tag + ".zin"]).cuda())
        assert rel_err(z, g[tag + ".z"]) <= RTOL_FP32 and rel_err(ld, g[tag + ".ld"]) <= RTOL_FP32
        assert rel_err(x, g[tag + ".x_inv"]) <= RTOL_FP32 and rel_err(ldi, g[tag + ".ld_inv"]) <= RTOL_FP32


def test_planar_radial_match_reference():
    _, _, flows, _ = _mods()
    g = golden("planar_radial.npz")
    for tag, d in (("d128", 128), ("d5", 5)):
        x = T(g[tag + ".x"]).cuda()
        pl = flows.Planar(d)
        pl.load_state_dict({k: T(g[f"{tag}.planar.{k}"]) for k in ("w", "u", "b")})
        rd = flows.Radial(d)
        rd.load_state_dict({k: T(g[f"{tag}.radial.{k}"]) for k in ("x0", "log_alpha", "beta")})
        pl, rd = pl.cuda(), rd.cuda()
        with torch.no_grad():
            z, ld = pl.forward(x)
            zr, ldr = rd.forward(x)
        assert rel_err(z, g[tag + ".planar.z"]) <= RTOL_FP32 and rel_err(ld, g[tag + ".planar.ld"]) <= RTOL_FP32
        assert ldr.shape == (1,)                                       # Q9
        assert rel_err(zr, g[tag + ".radial.z"]) <= RTOL_FP32 and rel_err(ldr, g[tag + ".radial.ld"]) <= RTOL_FP32
        with pytest.raises(NotImplementedError):
            pl.inverse(x)


def test_nsf_cl_layers_whole_layer():
    _, _, flows, _ = _mods()
    for name in ("nsfcl_d64.npz", "nsfcl_lj38.npz"):
        g = golden(name)
        size, dim, K, B, H = int(g["size"]), int(g["dim"]), int(g["K"]), float(g["B"]), int(g["H"])
        for mi, mask in enumerate(parse_masks(g)):
            p = f"m{mi}."
            layer = flows.NSF_CL(size, dim=dim, K=K, B=B, hidden_dim=H, mask=mask)
            layer.load_state_dict(sub_sd(g, p + "sd."))
            layer = layer.cuda()
            with torch.no_grad():
                z, ld = layer.forward(T(g[p + "x"]).cuda())
                xi, ldi = layer.inverse(T(g[p + "zin"]).cuda())
            # conditioner in fp32 differs from the reference's addmm by round-off only
            assert rel_err(z, g[p + "z"]) <= 2e-5 and rel_err(ld, g[p + "ld"]) <= 5e-5, (name, mask)
            assert rel_err(xi, g[p + "x_inv"]) <= 2e-5 and rel_err(ldi, g[p + "ld_inv"]) <= 5e-5, (name, mask)


def test_models_match_reference():
    _, _, flows, models = _mods()
    g = golden("models.npz")
    dev = torch.device("cuda")
    fl = [flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=16, mask=[i % 2]) for i in range(8)]
    m = models.NormalizingFlowModel(models.GaussianPrior(64, device=dev), fl, device=dev)
    m.load_state_dict(sub_sd(g, "nsf.sd."))
    m = m.to(dev)
    z, plp, ld = m.forward(T(g["nsf.x"]).cuda())
    assert ld.dtype == torch.float32 and ld.shape == (64,)
    for a, k in ((z, "nsf.z"), (plp, "nsf.prior_lp"), (ld, "nsf.ld")):
        assert rel_err(a.detach(), g[k]) <= 5e-5, k
    ev = m.evaluate(T(g["nsf.x"]).cuda())
    assert not ev.requires_grad and rel_err(ev, g["nsf.evaluate"]) <= 5e-5
    x, ldi = m.inverse(T(g["nsf.zin"]).cuda())
    assert rel_err(x.detach(), g["nsf.x_inv"]) <= 5e-5 and rel_err(ldi.detach(), g["nsf.ld_inv"]) <= 5e-5
    xs, lpx, zs = m.sample(10)
    assert xs.shape == (10, 64) and lpx.shape == (10,) and zs.shape == (10, 64) and not xs.requires_grad
    # bf16 conditioner: 1e-2 class
    for f in fl:
        f.psi.precision = "bf16"
    with torch.no_grad():
        z16, plp16, ld16 = m.forward(T(g["nsf.x"]).cuda())
    # 16-bit conditioner GEMMs (fp16 operands in the fused layer kernel): the north star's 1e-2, chain of 8 layers
    assert rel_err(z16, g["nsf.z"]) <= 1e-2 and rel_err(ld16, g["nsf.ld"]) <= 1e-2, (rel_err(z16, g["nsf.z"]), rel_err(ld16, g["nsf.ld"]))

    fl = [flows.RealNVP(2, hidden_dim=12) for _ in range(8)]
    m = models.NormalizingFlowModel(models.GaussianPrior(2, device=dev), fl, device=dev)
    m.load_state_dict(sub_sd(g, "rnvp.sd."))
    m = m.to(dev)
    with torch.no_grad():
        z, plp, ld = m.forward(T(g["rnvp.x"]).cuda())
        x, ldi = m.inverse(T(g["rnvp.x"]).cuda())
    assert rel_err(z, g["rnvp.z"]) <= 2e-5 and rel_err(ld, g["rnvp.ld"]) <= 2e-5 and rel_err(plp, g["rnvp.prior_lp"]) <= 2e-5
    assert rel_err(x, g["rnvp.x_inv"]) <= 2e-5 and rel_err(ldi, g["rnvp.ld_inv"]) <= 2e-5

    fl = [flows.Planar(16) for _ in range(6)]
    m = models.NormalizingFlowModel(models.GaussianPrior(16, device=dev), fl, device=dev)
    m.load_state_dict(sub_sd(g, "planar.sd."))
    m = m.to(dev)
    with torch.no_grad():
        z, plp, ld = m.forward(T(g["planar.x"]).cuda())          # six layers, ONE fused launch
    assert rel_err(z, g["planar.z"]) <= 2e-5 and rel_err(ld, g["planar.ld"]) <= 2e-5
    assert rel_err(plp, g["planar.prior_lp"]) <= 2e-5


@pytest.mark.parametrize("H", [128, 24, 100])
def test_fused_layer_kernel_matches_unfused_and_reference(H):
    """nsf_fused.cu (MLP on tcgen05 with fp16 operands + spline epilogue, one launch) vs (a) the unfused
    bf16 path on the same weights (bf16 class) and (b) the fp32 oracle (z within 3e-3, log-det — a sum of
    32 terms — within 1e-2), including a ragged batch whose last tile is padded."""
    from oracle import nf_oracle as O
    _, _, flows, _ = _mods()
    from normalizingflow_b200 import _fused
    torch.manual_seed(H)
    for mask in ([0], [1]):
        layer = flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=H, mask=mask)
        layer.psi.precision = "bf16"
        layer = layer.cuda()
        assert _fused.eligible(layer)
        N = 128 * 5 + 37
        x = (torch.randn(N, 64, generator=torch.Generator().manual_seed(3)) * 1.5)
        sd = {k: v.detach().cpu() for k, v in layer.state_dict().items()}
        for inv in (False, True):
            with torch.no_grad():
                layer.fused = True
                zf, lf = layer._transform(x.cuda(), inv)
                layer.fused = False
                zu, lu = layer._transform(x.cuda(), inv)
            ro, rl, _, _ = O.nsf_cl(x, sd, 32, 2, mask, 8, 3.0, inv, prefix="psi.")
            assert zf.shape == (N, 64) and lf.shape == (N,)
            # the fused kernel's operands are fp16 (11-bit significands), the unfused path's bf16: the two differ
            # by the bf16 class, and the fused kernel is ~8x closer to the fp32 oracle
            assert rel_err(zf, zu) <= 1e-2 and rel_err(lf, lu) <= 5e-2, (mask, inv, rel_err(zf, zu), rel_err(lf, lu))
            assert rel_err(zf, ro) <= 3e-3 and rel_err(lf, rl) <= 1e-2, (mask, inv, rel_err(zf, ro), rel_err(lf, rl))
            # conditioning columns pass through bit-exactly, in the reference's column order (Q5)
            assert torch.equal(zf.cpu().reshape(N, 32, 2)[:, :, 0], x.reshape(N, 32, 2)[:, :, mask[0]])
            acc = torch.full((N,), 1.5, device="cuda")
            with torch.no_grad():
                layer.fused = True
                _, la = layer._transform(x.cuda(), inv, acc)
            assert rel_err(la, lf + 1.5) <= 1e-6


def test_host_streaming_entry_points_equal_device_calls():
    _, _, flows, models = _mods()
    dev = torch.device("cuda")
    torch.manual_seed(1)
    fl = [flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=16, mask=[i % 2]) for i in range(3)]
    m = models.NormalizingFlowModel(models.GaussianPrior(64, device=dev), fl, device=dev).to(dev)
    x = torch.randn(1000, 64).pin_memory()
    lp = m.evaluate_host(x, chunk_rows=256)
    xs, lpx = m.inverse_host(x, chunk_rows=300)
    torch.cuda.synchronize()
    ref = m.evaluate(x.cuda())
    with torch.no_grad():
        rx, rld = m.inverse(x.cuda())
        rlpx = m.prior.log_prob(x.cuda()) - rld
    assert torch.equal(lp, ref.cpu()) and torch.equal(xs, rx.cpu()) and torch.equal(lpx, rlpx.cpu())
    # wait=False: back-to-back calls overlap; host_sync() makes every result visible
    lp2 = m.evaluate_host(x, chunk_rows=256, wait=False)
    xs2, lpx2 = m.inverse_host(x, chunk_rows=256, wait=False)
    lp3 = m.evaluate_host(xs2.clone().pin_memory() if False else x, chunk_rows=256, wait=False)
    m.host_sync()
    assert torch.equal(lp2, ref.cpu()) and torch.equal(xs2, rx.cpu()) and torch.equal(lpx2, rlpx.cpu())
    assert torch.equal(lp3, ref.cpu())


def test_multivariate_normal_prior_is_routed_to_the_logprob_kernel():
    """The prior the reference builds, MultivariateNormal(0, vars*I) (applications/src/setup.py:25-30),
    passed straight into NormalizingFlowModel: log_prob runs on nfk_gauss_logprob (launch counted) and
    equals torch's; a full-covariance prior keeps torch's own log_prob."""
    _, _, flows, models = _mods()
    from normalizingflow_b200 import _lib
    dev = torch.device("cuda")
    torch.manual_seed(0)
    fl = [flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=16, mask=[i % 2]) for i in range(2)]
    x = torch.randn(300, 64, device=dev)
    for var in (1.0, 0.37):
        prior = torch.distributions.MultivariateNormal(torch.zeros(64, device=dev), var * torch.eye(64, device=dev))
        m = models.NormalizingFlowModel(prior, fl, device=dev).to(dev)
        why = []
        assert models._scaled_identity_var(prior, why) is not None, why
        assert abs(m._prior_var() - var) < 1e-6
        with torch.no_grad():
            n0 = _lib.launch_count()
            z, plp, ld = m.forward(x)
            n_fwd = _lib.launch_count() - n0
            ev = m.evaluate(x)
            xs, lpx, zs = m.sample(64)
        ref = prior.log_prob(z)
        assert rel_err(plp, ref) <= 2e-6
        assert rel_err(ev, ref + ld) <= 2e-6
        assert rel_err(lpx, prior.log_prob(zs) - m.inverse(zs)[1]) <= 2e-6
        m2 = models.NormalizingFlowModel(models.GaussianPrior(64, var, device=dev), fl, device=dev).to(dev)
        with torch.no_grad():
            n0 = _lib.launch_count()
            m2.forward(x)
            assert _lib.launch_count() - n0 == n_fwd      # same kernels as with the native GaussianPrior
    cov = torch.eye(64, device=dev) + 0.1
    full = torch.distributions.MultivariateNormal(torch.zeros(64, device=dev), cov)
    m = models.NormalizingFlowModel(full, fl, device=dev).to(dev)
    assert m._prior_var() is None
    with torch.no_grad():
        z, plp, _ = m.forward(x)
    assert rel_err(plp, full.log_prob(z)) <= 1e-6


def test_host_api_results_are_ready_when_the_call_returns():
    """evaluate_host / inverse_host with wait=True return host tensors whose device->host copies have
    landed (no torch.cuda.synchronize() by the caller); differently chunked calls keep their own pipes
    and host_sync() waits for all of them."""
    _, _, flows, models = _mods()
    dev = torch.device("cuda")
    torch.manual_seed(1)
    fl = [flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=32, mask=[i % 2]) for i in range(4)]
    for f in fl:
        f.psi.precision = "bf16"
    m = models.NormalizingFlowModel(models.GaussianPrior(64, device=dev), fl, device=dev).to(dev)
    hx = torch.randn(5000, 64, generator=torch.Generator().manual_seed(2)).pin_memory()
    with torch.no_grad():
        ref = m.evaluate(hx.to(dev)).cpu()
        rx, rld = m.inverse(hx.to(dev))
        rlp = (m.prior.log_prob(hx.to(dev)) - rld).cpu()
    out = torch.full((5000,), float("nan")).pin_memory()
    got = m.evaluate_host(hx, out=out, chunk_rows=1024)           # read immediately, no explicit sync
    assert torch.equal(got, ref)
    x2, lp2 = m.inverse_host(hx, chunk_rows=2048)
    assert torch.equal(x2, rx.cpu()) and rel_err(lp2, rlp) <= 1e-6
    o1 = torch.full((5000,), float("nan")).pin_memory()
    o2 = torch.full((5000,), float("nan")).pin_memory()
    m.evaluate_host(hx, out=o1, chunk_rows=512, wait=False)
    m.evaluate_host(hx, out=o2, chunk_rows=4096, wait=False)      # another pipe (different chunk shape)
    m.host_sync()
    assert torch.equal(o1, ref) and torch.equal(o2, ref)


def test_bf16_fcnn_backward_runs_on_the_tensor_core_kernels():
    """Training gradients of a generic bf16 FCNN (RealNVP s/t nets, NSF_AR conditioners, K != 8 layers):
    _wide.Mlp3WideFn (dgrad with fused tanh backward + nfk_wgrad_ws) vs the library cross-check path
    and vs fp32 autograd."""
    _, _, flows, _ = _mods()
    from normalizingflow_b200 import _bf16, _wide
    torch.manual_seed(4)
    for (n_in, H, n_out, N) in ((32, 200, 96, 1000), (1, 100, 1, 513), (76, 160, 874, 300)):
        net = flows.FCNN(n_in, n_out, H, precision="bf16").cuda()
        assert _wide.mlp3_grad_ok(net)
        x = torch.randn(N, n_in, device="cuda", requires_grad=True)
        gy = torch.randn(N, n_out, device="cuda")
        l0, l2, l4 = net.network[0], net.network[2], net.network[4]
        ps = [l0.weight, l0.bias, l2.weight, l2.bias, l4.weight, l4.bias]
        y = net(x)
        assert y.grad_fn is not None and "Mlp3WideFn" in type(y.grad_fn).__name__
        g_new = torch.autograd.grad(y, [x] + ps, gy)
        y_old = _bf16.MLP3Bf16Fn.apply(x, *ps, net)
        g_old = torch.autograd.grad(y_old, [x] + ps, gy)
        net.precision = "fp32"
        y32 = net(x)
        g32 = torch.autograd.grad(y32, [x] + ps, gy)
        net.precision = "bf16"
        assert rel_err(y, y32.detach().double().cpu()) <= 1e-2
        for a, b, c in zip(g_new, g_old, g32):
            scale = float(c.abs().max()) + 1e-12
            assert float((a - c).abs().max()) <= 2e-2 * scale, (n_in, H, n_out, float((a - c).abs().max()) / scale)
            assert float((a - b).abs().max()) <= 2e-2 * scale


@pytest.mark.parametrize("L", [1, 5, 32, 40])
@pytest.mark.parametrize("N", [1, 17, 1000])
def test_planar_stack_gram_form_matches_the_sequential_reference(L, N):
    """d = 128 stacks run in the Gram-matrix form on the tensor cores (csrc/planar_mma.cu; 32 layers per launch,
    fp16 hi + lo operands): against the oracle's layer-by-layer chain (nf/flows_1.py:42-60) in fp32 and fp64, and
    against the per-layer register kernel."""
    ops, _, _, _ = _mods()
    from oracle import nf_oracle as O
    gen = torch.Generator().manual_seed(100 * L + N)
    d = 128
    w = (torch.rand(L, d, generator=gen) * 2 - 1) / d ** 0.5 * 3.0          # 3x the default init scale
    u = (torch.rand(L, d, generator=gen) * 2 - 1) / d ** 0.5 * 3.0
    b = (torch.rand(L, generator=gen) * 2 - 1)
    x = torch.randn(N, d, generator=gen) * 1.5
    z32, ld32 = x, torch.zeros(N)
    z64, ld64 = x.double(), torch.zeros(N, dtype=torch.float64)
    for l in range(L):
        z32, l32 = O.planar(z32, w[l], u[l], b[l:l + 1])
        ld32 = ld32 + l32
        z64, l64 = O.planar(z64, w[l].double(), u[l].double(), b[l:l + 1].double())
        ld64 = ld64 + l64
    z, ld = ops.planar_stack(x.cuda(), w.cuda(), u.cuda(), b.cuda())
    assert rel_err(z, z64) <= 1e-5 and rel_err(ld, ld64) <= 1e-5, (rel_err(z, z64), rel_err(ld, ld64))
    assert rel_err(z, z32) <= 1e-5 and rel_err(ld, ld32) <= 1e-5, (rel_err(z, z32), rel_err(ld, ld32))
    acc = torch.full((N,), 0.5, device="cuda")
    z2, ld2 = ops.planar_stack(x.cuda(), w.cuda(), u.cuda(), b.cuda(), logdet=acc)
    assert torch.equal(z2, z) and rel_err(ld2, ld64 + 0.5) <= 1e-5
    ops.PLANAR_MMA = False
    try:
        zo, ldo = ops.planar_stack(x.cuda(), w.cuda(), u.cuda(), b.cuda())
    finally:
        ops.PLANAR_MMA = True
    assert rel_err(z, zo.cpu()) <= 1e-5 and rel_err(ld, ldo.cpu()) <= 1e-5


@pytest.mark.parametrize("per_sample", [True, False], ids=["per_sample", "batch_global"])
@pytest.mark.parametrize("d,L,N", [(128, 32, 1000), (64, 3, 17), (32, 1, 1), (256, 5, 300)])
def test_radial_stack_matches_the_layer_by_layer_reference(per_sample, d, L, N):
    """Fused runs of radial layers (csrc/radial_stack.cu) against the oracle's layer-by-layer chain
    (nf/flows_1.py:85-97: batch-global Frobenius norm, quirk Q9, and the per-sample variant) and against the
    per-layer kernels; through NormalizingFlowModel, which builds the fused run."""
    _, _, flows, models = _mods()
    from oracle import nf_oracle as O
    dev = torch.device("cuda")
    torch.manual_seed(7 * d + L)
    layers = [flows.Radial(d, per_sample=per_sample) for _ in range(L)]
    m = models.NormalizingFlowModel(models.GaussianPrior(d, device=dev), layers, device=dev).to(dev)
    x = torch.randn(N, d, generator=torch.Generator().manual_seed(N)) * 1.3
    z64 = x.double()
    ld64 = torch.zeros(N if per_sample else 1, dtype=torch.float64)
    for l in layers:
        z64, l64 = O.radial(z64, l.x0.detach().cpu().double(), l.log_alpha.detach().cpu().double(),
                            l.beta.detach().cpu().double(), per_sample=per_sample)
        ld64 = ld64 + l64
    with torch.no_grad():
        plan = m._forward_plan()
        assert (len(plan) == 1 and type(plan[0]).__name__ == "RadialStack") or L == 1
        z, plp, ld = m.forward(x.to(dev))
        m.fuse_planar = False
        z1, _, ld1 = m.forward(x.to(dev))                       # per-layer kernels
    assert rel_err(z, z64) <= 1e-5, rel_err(z, z64)
    ref_ld = ld64 if per_sample else ld64.expand(N)
    assert rel_err(ld, ref_ld) <= 2e-5, rel_err(ld, ref_ld)
    assert rel_err(z, z1.cpu()) <= 1e-5 and rel_err(ld, ld1.cpu()) <= 2e-5
