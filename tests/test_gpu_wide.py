"""GPU parity tests of the wide conditioner path (csrc/gemm_ws.cu: persistent warp-specialised
tcgen05 GEMM over image-layout operands) against fp64 matmuls of the same bf16-rounded operands,
and of a whole NSF_CL layer with the class-default hidden width 800 (nf/flows.py:216) against
the oracle."""
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _wide():
    from normalizingflow_b200 import _wide
    return _wide


@pytest.fixture(autouse=True, params=[0, 1], ids=["single_cta", "cta_pair"])
def pair_mode(request):
    """Every test runs with the wide-path GEMMs forced to single-CTA tiles and to CTA pairs
    (tcgen05 cta_group::2; needs >= 2 M tiles, smaller problems fall back to single)."""
    from normalizingflow_b200 import _lib
    _lib.lib.nfk_set_gemm_ws_pair_mode(request.param)
    yield request.param
    _lib.lib.nfk_set_gemm_ws_pair_mode(-1)


def _layer(w, b, kb, fmt=0):
    W = _wide()
    n_out, k_in = w.shape
    tiles = W.plan_tiles(W.blocks(n_out))
    w_img, bp = W.weight_image(w.cuda(), b.cuda(), kb, tiles, fmt)
    return dict(w=w_img, b=bp, KB=kb, kmma_last=(k_in - 64 * (kb - 1) + 15) // 16, tiles=tiles, n_out=n_out, fmt=fmt)


FMT_DTYPE = {0: torch.bfloat16, 1: torch.float16}


@pytest.mark.parametrize("M,K,N,act,f32", [(128, 64, 64, 1, False), (1, 32, 800, 1, False), (300, 800, 800, 1, False),
                                           (1000, 800, 736, 0, True), (129, 76, 874, 0, True),
                                           (20000, 800, 1748, 0, True), (40000, 832, 832, 1, False),
                                           (257, 16, 100, 0, True)])
@pytest.mark.parametrize("fmt", [0, 1], ids=["bf16", "fp16"])
def test_gemm_ws_matches_fp64(M, K, N, act, f32, fmt):
    W = _wide()
    dt = FMT_DTYPE[fmt]
    from normalizingflow_b200._lib import i32_array
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    kb = W.blocks(K)
    a_img = W.pack_input(x.cuda(), K, 1, [0], kb, fmt)
    # the packed image round-trips to the input rounded to the image format
    back = W.image_to_rows(a_img, M, K).float().cpu()
    assert torch.equal(back, x.to(dt).float())
    # the device packer (nfk_pack_w_img) and the host packer build the same weight image
    tiles0 = W.plan_tiles(W.blocks(N))
    assert torch.equal(W.pack_weight(w.cuda(), kb, tiles0, fmt=fmt).cpu(), W.weight_image(w, b, kb, tiles0, fmt)[0])
    lay = _layer(w, b, kb, fmt)
    lay["tiles_c"] = i32_array(lay["tiles"])
    y = W.gemm(a_img, lay, M, act, f32)
    torch.cuda.synchronize()
    ref = torch.nn.functional.linear(x.to(dt).double(), w.to(dt).double(), b.double())
    if act:
        ref = torch.tanh(ref)
    if f32:
        got = y.float().cpu()
    else:
        got = W.image_to_rows(y, M, N).float().cpu()
        # padding columns of the image are exact zeros (they are the next layer's K padding)
        full = W.image_to_rows(y, M, sum(lay["tiles"]) * 64).float().cpu()
        assert torch.count_nonzero(full[:, N:]) == 0
    tol = 2e-3 if (f32 and not act) else (1e-2 if fmt == 0 else 2e-3)
    assert rel_err(got, ref) <= tol, rel_err(got, ref)


def test_wide_mlp3_matches_16bit_chain():
    """FCNN 32 -> 800 -> 800 -> 736 through the three wide GEMMs vs an fp64 evaluation that rounds
    the operands to the inference image format (fp16) at the same places."""
    W = _wide()
    from normalizingflow_b200 import flows
    torch.manual_seed(0)
    net = flows.FCNN(32, 736, 800, precision="bf16").cuda()
    x = torch.randn(3000, 64, generator=torch.Generator().manual_seed(5))
    y = W.mlp3(net, x.cuda(), 32, 2, [1]).cpu()
    bf = lambda t: t.to(FMT_DTYPE[W.INFER_FMT]).double()
    l0, l2, l4 = net.network[0], net.network[2], net.network[4]
    xc = bf(x.reshape(-1, 32, 2)[:, :, 1])
    h1 = bf(torch.tanh(xc @ bf(l0.weight.cpu()).T + l0.bias.double().cpu()).float())
    h2 = bf(torch.tanh(h1 @ bf(l2.weight.cpu()).T + l2.bias.double().cpu()).float())
    ref = h2 @ bf(l4.weight.cpu()).T + l4.bias.double().cpu()
    assert rel_err(y, ref) <= 2e-3, rel_err(y, ref)
    # and against the plain fp32 MLP: the north star's 1e-2 class for the 16-bit conditioner GEMMs
    f32 = lambda t: t.double().cpu()
    xc32 = x.reshape(-1, 32, 2)[:, :, 1].double()
    r32 = torch.tanh(torch.tanh(xc32 @ f32(l0.weight).T + f32(l0.bias)) @ f32(l2.weight).T + f32(l2.bias)) \
        @ f32(l4.weight).T + f32(l4.bias)
    assert rel_err(y, r32) <= 2e-3, rel_err(y, r32)


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("mask", [[0], [1]])
@pytest.mark.parametrize("arith", ["fast", "hybrid", "exact"])
def test_rqs_epilogue_matches_unfused(inverse, mask, arith):
    """GEMM-with-spline-epilogue vs the same bf16 conditioner followed by the stand-alone spline
    kernel: same parameters up to accumulation order, so outputs agree far inside the 1e-2 class;
    also checks the tail tile (N not a multiple of 128) and log-det accumulation."""
    from normalizingflow_b200 import flows
    torch.manual_seed(2)
    lay = flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=200, mask=mask, arith=arith)
    lay.psi.precision = "bf16"
    lay = lay.cuda()
    with torch.no_grad():
        lay.psi.network[4].weight.mul_(4.0)          # non-uniform bins
    x = torch.randn(1000, 64, generator=torch.Generator().manual_seed(4)).cuda()
    x[0, :4] = torch.tensor([3.0, -3.0, 3.5, -7.0])  # tail bound and identity tails
    ld0 = torch.randn(1000, generator=torch.Generator().manual_seed(5)).cuda()
    with torch.no_grad():
        lay.fused = True
        o1, l1 = lay._transform(x, inverse, ld0.clone())
        lay.fused = False
        o2, l2 = lay._transform(x, inverse, ld0.clone())
    assert rel_err(o1.cpu(), o2.cpu().double()) <= 2e-4
    assert rel_err(l1.cpu(), l2.cpu().double()) <= 1e-3


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("mask", [[0], [1], [2], [0, 1], [1, 2], [0, 2]])
def test_rqs_epilogue_lj38_geometry(inverse, mask):
    """LJ-38 geometry (size 38, dim 3, all six masks of applications/src/setup.py:60-61): the
    spline epilogue's column bookkeeping (quirk Q5 output order) vs the stand-alone kernel."""
    from normalizingflow_b200 import flows
    torch.manual_seed(3)
    lay = flows.NSF_CL(38, dim=3, K=8, B=4.0, hidden_dim=160, mask=mask, arith="hybrid")
    lay.psi.precision = "bf16"
    lay = lay.cuda()
    with torch.no_grad():
        lay.psi.network[4].weight.mul_(4.0)
    x = (1.5 * torch.randn(700, 114, generator=torch.Generator().manual_seed(6))).cuda()
    with torch.no_grad():
        lay.fused = True
        o1, l1 = lay._transform(x, inverse)
        lay.fused = False
        o2, l2 = lay._transform(x, inverse)
    assert rel_err(o1.cpu(), o2.cpu().double()) <= 2e-4
    assert rel_err(l1.cpu(), l2.cpu().double()) <= 1e-3


@pytest.mark.parametrize("inverse", [False, True])
def test_nsf_layer_h800_vs_oracle(inverse):
    """Whole NSF_CL layer at the class-default hidden width: wide 16-bit conditioner (fp16 operand images on
    the inference forward) with the spline epilogue vs the fp32 oracle: the north star's 1e-2 for BOTH z
    and log_det (a sum of 32 per-feature terms)."""
    from normalizingflow_b200 import flows
    from oracle import nf_oracle as O
    torch.manual_seed(1)
    lay = flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=800, mask=[1])
    sd = {k: v.detach().clone() for k, v in lay.state_dict().items()}
    lay.psi.precision = "bf16"
    lay = lay.cuda()
    x = torch.randn(2500, 64, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        out, ld = (lay.inverse if inverse else lay.forward)(x.cuda())
    ro, rld = O.nsf_cl(x, sd, 32, 2, [1], 8, 3.0, inverse)[:2]
    assert rel_err(out.cpu(), ro) <= 1e-2, rel_err(out.cpu(), ro)
    assert rel_err(ld.cpu(), rld) <= 1e-2, rel_err(ld.cpu(), rld)


@pytest.mark.parametrize("N", [0, 1, 127, 129, 300])
def test_wide_layer_edge_batches_and_nan_tails(N):
    """Empty, single-row and ragged batches (tail tile of the 128-row M tiling), NaN / out-of-range
    inputs pass through as identity with zero log-det (quirk Q6) on the spline-epilogue path."""
    from normalizingflow_b200 import flows
    torch.manual_seed(4)
    lay = flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=192, mask=[0])
    lay.psi.precision = "bf16"
    lay = lay.cuda()
    x = torch.randn(N, 64, generator=torch.Generator().manual_seed(9)).cuda()
    if N:
        x[0, 1] = float("nan")
        x[0, 3] = 1e30
        x[0, 5] = -3.0000002
    with torch.no_grad():
        lay.fused = True
        o1, l1 = lay.forward(x)
        lay.fused = False
        o2, l2 = lay.forward(x)
    assert o1.shape == (N, 64) and l1.shape == (N,)
    if N:
        assert torch.isnan(o1[0, 1]) and float(o1[0, 3]) == float(x[0, 3]) and float(o1[0, 5]) == float(x[0, 5])
        ok = ~torch.isnan(o2)
        assert rel_err(o1[ok].cpu(), o2[ok].cpu().double()) <= 2e-4
        assert rel_err(l1.cpu(), l2.cpu().double()) <= 1e-3


@pytest.mark.parametrize("M,P,Q,pad", [(128, 64, 64, False), (1000, 200, 96, False), (5000, 800, 800, False),
                                        (3000, 736, 200, True), (700, 160, 38, False), (40000, 128, 128, False)])
def test_wgrad_ws_matches_fp64(M, P, Q, pad):
    """Batch-contraction GEMM straight from two bf16 images (MN-major tcgen05 operands, split-K with
    fp32 atomics) vs an fp64 matmul of the same bf16-rounded operands; image column sums (bias
    gradients) vs a plain sum."""
    W = _wide()
    g = torch.Generator().manual_seed(M + P + Q)
    p_cols = ((P + 22) // 23) * 24 if pad else P                     # padded: 24 columns per feature, 24th zero
    a = torch.randn(M, p_cols, generator=g)
    if pad:
        a.reshape(M, -1, 24)[:, :, 23] = 0
    b = torch.randn(M, Q, generator=g)
    a_img = W.pack_input(a.cuda(), p_cols, 1, [0], W.blocks(p_cols))
    b_img = W.pack_input(b.cuda(), Q, 1, [0], W.blocks(Q))
    c = W.wgrad(a_img, b_img, M, P, Q, pad_p=pad).cpu()
    ab = a.to(torch.bfloat16).double()
    ref = ab.t() @ b.to(torch.bfloat16).double()
    if pad:
        ref = ref.reshape(-1, 24, Q)[:, :23].reshape(-1, Q)[:P]
    scale = float(ref.abs().max())
    assert float((c - ref).abs().max()) <= 1e-4 * scale, float((c - ref).abs().max()) / scale
    cs = W.image_colsum(a_img, p_cols).cpu()
    assert float((cs - ab.sum(0)).abs().max()) <= 1e-4 * float(ab.sum(0).abs().max() + 1)
