"""CPU-only tests of the host side: the C-ABI library loads and exports everything include/nfk.h
declares, the layer classes keep the reference's constructor signatures / state-dict keys /
error behaviour, the HMC driver reproduces the reference's chain, weight packing for the fused
kernel, and the multi-rank helpers (gloo, world size 2)."""
import os
import re

import numpy as np
import pytest
import torch

from tests.helpers import T, golden, sub_sd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from normalizingflow_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "nfk.h")).read()
    declared = set(re.findall(r"\b(nfk_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    missing = [n for n in sorted(declared) if not hasattr(_lib.lib, n)]
    assert not missing, missing
    # every declared function has a ctypes signature and vice versa
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.MISSING == []
    assert _lib.lib.nfk_abi_version() == 2
    assert _lib.lib.nfk_nsf_fused_rows_per_tile() == 128


def test_argument_errors_without_touching_a_gpu():
    """shape checks run on the host before any launch: NFK_EINVAL -> ValueError (nf/utils.py:64-71)"""
    import ctypes
    from normalizingflow_b200 import _lib
    bad_mask = (ctypes.c_int32 * 1)(5)
    rc = _lib.lib.nfk_rqs_coupling(None, None, None, None, None, 4, 32, 2, bad_mask, 1, 8, 3.0, 0, 0, 1, None)
    assert rc == _lib.NFK_EINVAL and b"mask" in _lib.lib.nfk_last_error()
    ok_mask = (ctypes.c_int32 * 1)(1)
    rc = _lib.lib.nfk_rqs_coupling(None, None, None, None, None, 4, 32, 2, ok_mask, 1, 2000, 3.0, 0, 0, 1, None)
    assert rc == _lib.NFK_EINVAL
    assert _lib.lib.nfk_rqs_coupling(None, None, None, None, None, 0, 32, 2, ok_mask, 1, 8, 3.0, 0, 0, 1, None) == 0
    with pytest.raises(ValueError):
        _lib.check(_lib.NFK_EINVAL, "x")
    with pytest.raises(RuntimeError):
        _lib.check(_lib.NFK_ECUDA, "x")


def test_state_dict_keys_and_ctor_signatures_match_reference():
    import inspect
    from nf.flows import FCNN, NSF_CL, Planar, Radial, RealNVP
    from nf.models import NormalizingFlowModel
    g = golden("nsfcl_d64.npz")
    ref_keys = sorted(k[len("m0.sd."):] for k in g.files if k.startswith("m0.sd."))
    layer = NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=24, mask=[0])
    assert sorted(layer.state_dict().keys()) == ref_keys
    layer.load_state_dict(sub_sd(g, "m0.sd."))                        # reference checkpoint loads
    assert isinstance(layer.mask, torch.Tensor) and layer.mask.dtype == torch.int64      # Q14
    assert "mask" not in layer.state_dict()
    sig = inspect.signature(NSF_CL.__init__).parameters
    assert [sig[k].default for k in ("dim", "K", "B", "hidden_dim", "device", "mask")] == [3, 32, 3, 800, "cpu", [1]]
    g = golden("realnvp.npz")
    ref_keys = sorted(k[len("d6.sd."):] for k in g.files if k.startswith("d6.sd."))
    assert sorted(RealNVP(6, hidden_dim=40).state_dict().keys()) == ref_keys
    assert inspect.signature(RealNVP.__init__).parameters["hidden_dim"].default == 800
    assert sorted(Planar(5).state_dict().keys()) == ["b", "u", "w"]
    assert sorted(Radial(5).state_dict().keys()) == ["beta", "log_alpha", "x0"]
    g = golden("models.npz")
    ref_keys = sorted(k[len("nsf.sd."):] for k in g.files if k.startswith("nsf.sd."))
    m = NormalizingFlowModel(None, [NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=16, mask=[i % 2]) for i in range(8)])
    assert sorted(m.state_dict().keys()) == ref_keys
    assert list(FCNN(3, 4, 5).state_dict().keys()) == [f"network.{i}.{p}" for i in (0, 2, 4) for p in ("weight", "bias")]


def test_no_cpu_path_and_reference_error_behaviour():
    from nf.flows import NSF_CL, Planar, Radial
    from nf.utils import unconstrained_RQS
    layer = NSF_CL(4, dim=3, K=8, B=3.0, hidden_dim=8, mask=[1])
    with pytest.raises(RuntimeError, match="no CPU path"):
        layer.forward(torch.zeros(2, 12))
    with pytest.raises(NotImplementedError):
        Planar(4).inverse(torch.zeros(1, 4))                            # flows_1.py:62-63
    with pytest.raises(NotImplementedError):
        Radial(4).inverse(torch.zeros(1, 4))
    with pytest.raises(NotImplementedError):
        Planar(4, nonlinearity=torch.nn.functional.elu)
    with pytest.raises(ValueError, match="Minimal bin width"):           # utils.py:68-69
        unconstrained_RQS(torch.zeros(1), torch.zeros(1, 2000), torch.zeros(1, 2000), torch.zeros(1, 1999))


class _HarmonicSim:
    """the duck-typed simulation oracle/gen_golden.py drove the reference HMC with"""

    def __init__(self, nparticles, dim, k=1.0):
        self.nparticles, self.dim, self.k = nparticles, dim, k
        self.position = torch.linspace(-1, 1, nparticles * dim)
        self.velocity = torch.zeros(nparticles * dim)

    def get_position(self):
        return self.position

    def get_potential(self):
        return 0.5 * self.k * torch.sum(self.position ** 2)

    def set_position(self, p):
        self.position = p.flatten().clone()

    def set_velocity(self, v):
        self.velocity = v.flatten().clone()

    def integration_step(self, path_len, dt):
        q, v = self.position, self.velocity
        f = -self.k * q
        for _ in range(path_len):
            v = v + 0.5 * dt * f
            q = q + dt * v
            f = -self.k * q
            v = v + 0.5 * dt * f
        self.position, self.velocity = q, v
        return q, self.get_potential()


def test_hmc_driver_reproduces_reference_chain():
    """velocity law, potential-only Metropolis rule (Q12) and bookkeeping of nf/hmc.py:43-65"""
    from nf.hmc import HMC
    g = golden("hmc.npz")
    sim = _HarmonicSim(int(g["nparticles"]), int(g["dim"]), k=float(g["k"]))
    torch.manual_seed(int(g["seed"]))
    h = HMC(sim, path_len=int(g["path_len"]), dt=float(g["dt"]), dim=int(g["dim"]), beta=float(g["beta"]))
    pos, pot, logp, acc = h.hmc(epochs=int(g["epochs"]))
    assert pos.shape == g["positions"].shape
    np.testing.assert_allclose(pos.numpy(), g["positions"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(pot.numpy(), g["potentials"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(logp.numpy(), g["last_logp"], rtol=1e-6, atol=1e-6)
    assert acc == pytest.approx(float(g["accept"]))


def test_fused_weight_image_is_the_sw128_layout():
    from normalizingflow_b200._fused import _swizzle_image
    rows, kb = 24, 2
    m = torch.arange(rows * 64 * kb, dtype=torch.float32).reshape(rows, 64 * kb).to(torch.bfloat16)
    img = _swizzle_image(m)
    assert img.shape == (kb, rows, 8, 8)
    flat = img.reshape(kb, rows * 64)
    for b in range(kb):
        for r in (0, 1, 7, 8, 13, 23):
            for c in range(8):
                src = m[r, b * 64 + c * 8: b * 64 + c * 8 + 8]
                dst = flat[b, r * 64 + ((c ^ (r & 7)) * 8): r * 64 + ((c ^ (r & 7)) * 8) + 8]
                assert torch.equal(src, dst)


def test_reference_checkpoint_loads_unchanged(tmp_path):
    """A .pth written by the reference's train.py:39-40 loads into the drop-in model with no missing
    or unexpected keys; optimizer / scheduler state round-trips too."""
    import os as _os
    from normalizingflow_b200 import checkpoint, flows, models
    fl = [flows.NSF_CL(4, dim=2, K=8, B=3.0, hidden_dim=8, mask=[i % 2]) for i in range(2)] + [flows.RealNVP(8, hidden_dim=8)]
    m = models.NormalizingFlowModel(models.GaussianPrior(8, device="cpu"), fl)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=10, gamma=0.5)
    path = _os.path.join(ROOT, "tests", "golden", "ref_checkpoint.pth")
    epoch, losses, res = checkpoint.load_checkpoint(path, m, opt, sched)
    assert epoch == 3 and len(losses) == 3
    assert list(res.missing_keys) == [] and list(res.unexpected_keys) == []
    blob = torch.load(path, map_location="cpu", weights_only=False)
    for k, v in blob["model"].items():
        assert torch.equal(m.state_dict()[k], v), k
    assert opt.state_dict()["state"][0]["step"] == 3 and sched.last_epoch == 3
    # and the writer produces the same container
    out = tmp_path / "mine.pth"
    checkpoint.save_checkpoint(out, m, opt, sched, epoch=4, losses=[1.0])
    again = torch.load(out, map_location="cpu", weights_only=False)
    assert set(again) == set(checkpoint.KEYS) and set(again["model"]) == set(blob["model"])


def test_wide_path_tile_plan_and_operand_images():
    """Host logic of the wide conditioner path (_wide.py): N-tile plan, weight / activation images
    (the SWIZZLE_128B shared-memory layout stored in HBM) and argument checks of its entry points."""
    import ctypes
    from normalizingflow_b200 import _lib, _wide
    for nblk in range(1, 40):
        tiles = _wide.plan_tiles(nblk)
        assert sum(tiles) == nblk and max(tiles) <= 4 and min(tiles) >= 1
        assert len(tiles) == (nblk + 3) // 4 and max(tiles) - min(tiles) <= 1
    assert _wide.plan_tiles(13) == [4, 3, 3, 3]                       # hidden 800 -> 832 padded columns
    # weight image: tile t, K block kb, row r, chunk slot j holds W[row0_t + r, kb*64 + (j ^ r%8)*8 ...]
    g = torch.Generator().manual_seed(0)
    w = torch.randn(100, 70, generator=g)
    b = torch.randn(100, generator=g)
    tiles = _wide.plan_tiles(_wide.blocks(100))                        # 2 blocks -> one tile
    img, bp = _wide.weight_image(w, b, _wide.blocks(70), tiles)
    assert img.numel() == 128 * 128 and bp.shape == (128,)
    assert torch.equal(bp[:100], b) and float(bp[100:].abs().sum()) == 0.0
    im = img.reshape(2, 128, 8, 8)
    wb = torch.zeros(128, 128, dtype=torch.bfloat16)
    wb[:100, :70] = w.to(torch.bfloat16)
    for kb in range(2):
        for r in (0, 5, 63, 99, 127):
            for j in range(8):
                c = j ^ (r & 7)
                assert torch.equal(im[kb, r, j], wb[r, kb * 64 + c * 8: kb * 64 + c * 8 + 8])
    # image_to_rows inverts the activation image layout
    rows = torch.randn(300, 192, generator=g).to(torch.bfloat16)
    pad = torch.zeros(384, 192, dtype=torch.bfloat16)
    pad[:300] = rows
    a_img = torch.stack([_wide._swizzle_image(pad[m * 128:(m + 1) * 128]) for m in range(3)]).reshape(3, 3, 128, 64)
    assert torch.equal(_wide.image_to_rows(a_img, 300, 192), rows)
    # argument errors are caught on the host (NFK_EINVAL), before any launch
    L = _lib.lib
    tiles_c = (ctypes.c_int32 * 1)(5)
    assert L.nfk_gemm_ws(None, None, None, None, 128, 1, 4, tiles_c, 1, 1, 0, 0, 0, None, 0, None) == _lib.NFK_EINVAL
    tiles_c = (ctypes.c_int32 * 1)(4)
    assert L.nfk_gemm_ws(None, None, None, None, 128, 1, 7, tiles_c, 1, 1, 0, 0, 0, None, 0, None) == _lib.NFK_EINVAL
    assert L.nfk_gemm_ws(None, None, None, None, 128, 1, 4, tiles_c, 1, 2, 0, 0, 0, None, 0, None) == _lib.NFK_EINVAL
    assert L.nfk_gemm_ws(None, None, None, None, 0, 1, 4, tiles_c, 1, 1, 0, 0, 0, None, 1, None) == 0
    # image format: only NFK_IMG_BF16 / NFK_IMG_F16; the tanh-backward epilogue reads bf16 images only
    assert L.nfk_gemm_ws(None, None, None, None, 128, 1, 4, tiles_c, 1, 1, 0, 0, 0, None, 2, None) == _lib.NFK_EINVAL
    assert L.nfk_gemm_ws(None, None, None, None, 128, 1, 4, tiles_c, 1, 2, 0, 0, 0, tiles_c, 1, None) == _lib.NFK_EINVAL
    # fp16 weight image on the host packer saturates instead of overflowing to inf
    img16, _ = _wide.weight_image(torch.tensor([[1e6, -1e6, 0.5]]), None, 1, [1], _wide.F16)
    assert img16.dtype == torch.float16 and float(img16.float().abs().max()) == 65504.0
    mask = (ctypes.c_int32 * 1)(3)
    rc = L.nfk_gemm_ws_rqs(None, None, None, None, None, None, 128, 13, 2, 32, 2, mask, 1, 3.0, 0, 0, 2, 0, None, None, None)
    assert rc == _lib.NFK_EINVAL and b"mask" in L.nfk_last_error()
    mask = (ctypes.c_int32 * 1)(1)
    rc = L.nfk_gemm_ws_rqs(None, None, None, None, None, None, 128, 13, 2, 200, 2, mask, 1, 3.0, 0, 0, 2, 1, None, None, None)
    assert rc == _lib.NFK_EINVAL and b"exceed" in L.nfk_last_error()
    assert L.nfk_gemm_ws_rqs_bwd(None, None, None, None, None, None, 1.0, None, None, 0, 13, 2, 32, 2, mask, 1, 3.0,
                                 0, None) == 0
    assert L.nfk_gemm_ws_rows_per_tile() == 128


def test_shard_rows_partitions_the_batch():
    from normalizingflow_b200.dist import shard_rows
    for n, w in ((1 << 20, 8), (10, 3), (7, 8), (0, 2)):
        spans = [shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from normalizingflow_b200 import dist as nd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = nd.init_from_env("gloo")
    torch.manual_seed(100 + rank)                       # ranks start from DIFFERENT weights
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    nd.broadcast_parameters(net, src=0)
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(64, 6, generator=gen)               # the GLOBAL batch; each rank takes its rows
    a, b = nd.shard_rows(64, r, w)
    loss = net(x[a:b]).pow(2).mean()
    loss.backward()
    nbytes = nd.allreduce_gradients(net.parameters(), average=True)
    # the persistent flat bucket gives the same averaged gradients with one in-place all-reduce
    flat_a = torch.cat([p.grad.flatten() for p in net.parameters()]).clone()
    bucket = nd.GradBucket(net.parameters())
    bucket.zero()
    net(x[a:b]).pow(2).mean().backward()
    assert bucket.allreduce() == nbytes
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in net.parameters())
    assert torch.allclose(bucket.flat, flat_a, rtol=1e-6, atol=1e-7)
    gm = nd.global_mean(net(x[a:b]).detach().flatten())
    flat = torch.cat([p.grad.flatten() for p in net.parameters()])
    wts = torch.cat([p.detach().flatten() for p in net.parameters()])
    if rank == 0:
        torch.save({"grad": flat, "w": wts, "bytes": nbytes, "gm": gm}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_matches_single_process_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "r0.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(100)                              # rank 0's weights were broadcast
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    x = torch.randn(64, 6, generator=torch.Generator().manual_seed(7))
    net(x).pow(2).mean().backward()                     # equal shard sizes: mean of means == global mean
    ref = torch.cat([p.grad.flatten() for p in net.parameters()])
    assert torch.allclose(got["w"], torch.cat([p.detach().flatten() for p in net.parameters()]))
    assert torch.allclose(got["grad"], ref, rtol=1e-5, atol=1e-6)
    assert got["bytes"] == ref.numel() * 4
    assert torch.allclose(got["gm"], net(x).detach().mean(), rtol=1e-5, atol=1e-6)


def test_grad_bucket_views_survive_zero_grad_set_to_none():
    """GradBucket keeps every .grad a view into one flat buffer; an optimizer.zero_grad(set_to_none=True)
    in between is repaired by bucket.zero(), and autograd accumulates into the views."""
    from normalizingflow_b200 import dist as nd
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Tanh(), torch.nn.Linear(3, 2))
    bucket = nd.GradBucket(net.parameters())
    assert bucket.nbytes == 4 * sum(p.numel() for p in net.parameters())
    x = torch.randn(5, 4)
    net(x).pow(2).sum().backward()
    ref = torch.cat([p.grad.flatten() for p in net.parameters()]).clone()
    assert torch.equal(bucket.flat, ref) and float(ref.abs().sum()) > 0
    torch.optim.SGD(net.parameters(), lr=0.1).zero_grad(set_to_none=True)
    assert all(p.grad is None for p in net.parameters())
    bucket.zero()
    assert all(p.grad is not None for p in net.parameters()) and float(bucket.flat.abs().sum()) == 0
    net(x).pow(2).sum().backward()
    assert torch.equal(bucket.flat, ref)
    assert bucket.allreduce() == 0            # not distributed: nothing to exchange


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the native one) prints ONE JSON line
    with the contract's keys and runs without a GPU."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-rows", "256"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rqs_flow_samples_per_sec_fwd_inv_logdet" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_backward_kernel_parameter_row_order_interleaves_width_and_height_logits():
    """csrc/nsf_fused_bwd.cu reads a feature's width / height logit pair from adjacent accumulator columns (packed
    fp32x2 arithmetic): the host permutes the 24 padded parameter rows of W3 / b3 accordingly."""
    from normalizingflow_b200 import _fused
    order = _fused._BWD_ROW_ORDER
    assert sorted(order) == list(range(24))
    assert [order[2 * j] for j in range(8)] == list(range(8))               # width logits (nf/utils.py:27-33: W first)
    assert [order[2 * j + 1] for j in range(8)] == list(range(8, 16))       # height logits
    assert order[16:] == list(range(16, 24))                                # derivative logits + the padding row


def test_flow_eligibility_for_the_one_launch_gradient_path_is_decided_on_the_host():
    """_fused.flow_eligible: every layer an NSF_CL(32, dim 2, K 8, hidden <= 128) with a 16-bit conditioner on the fused
    path, under an isotropic zero-mean Gaussian prior -- no kernel is launched to find out."""
    import torch
    from normalizingflow_b200 import _fused, flows, models
    if not _fused._lib.have("nfk_nsf_pairs_fused_bwd"):
        return

    def make(H, prec, K=8):
        fl = [flows.NSF_CL(32, dim=2, K=K, B=3.0, hidden_dim=H, mask=[i % 2]) for i in range(2)]
        for f in fl:
            f.psi.precision = prec
        return models.NormalizingFlowModel(models.GaussianPrior(64), fl)
    assert _fused.flow_eligible(make(128, "bf16")) and _fused.flow_eligible(make(16, "bf16"))
    assert not _fused.flow_eligible(make(200, "bf16"))          # wide conditioner: the multi-launch path
    assert not _fused.flow_eligible(make(128, "fp32"))          # parity mode: autograd through the fp32 kernels
    assert not _fused.flow_eligible(make(128, "bf16", K=4))
    m = make(128, "bf16")
    m.flows[1].fused = False
    assert not _fused.flow_eligible(m)
    mvn = torch.distributions.MultivariateNormal(torch.zeros(64), 0.25 * torch.eye(64))
    m2 = models.NormalizingFlowModel(mvn, list(make(64, "bf16").flows))
    assert _fused.flow_eligible(m2)                             # what the reference builds (applications/src/setup.py:25-30)
    m3 = models.NormalizingFlowModel(torch.distributions.MultivariateNormal(torch.ones(64), torch.eye(64)), list(make(64, "bf16").flows))
    assert not _fused.flow_eligible(m3)
