#!/usr/bin/env python
"""Benchmark of the RQS-flow hot path (BASELINE.json metric: RQS-flow samples/sec, one forward
+ one inverse pass, each with log-det, d=64, 8 layers, K=8 bins, batch 2^20 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); the batch is sharded by rows,
every rank evaluates 2^20 rows (weak scaling), no data-path collective.  Prints ONE JSON line
on rank 0.  `--impl reference` times the CPU restatement of the reference (oracle/) on the
host cores (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms (cpu_baseline, --impl reference)
# run on rank 0 only and must see all host cores, so undo that before torch loads its OpenMP runtime
if os.environ.get("RANK", "0") == "0" and os.environ.get("OMP_NUM_THREADS") == "1":
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rqs_flow_samples_per_sec_fwd_inv_logdet"
UNIT = "samples/s"
D, SIZE, DIM, LAYERS, KBINS, TAIL = 64, 32, 2, 8, 8, 3.0
ROW_BYTES_PER_LAYER = 32 * 23 * 4 + 64 * 4 + 64 * 4 + 8      # 3,464 B


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="rows per GPU")
    ap.add_argument("--hidden", type=int, default=128, help="conditioner hidden width")
    ap.add_argument("--arith", default="auto", choices=["auto", "hybrid", "exact", "fast"],
                    help="spline arithmetic (include/nfk.h); auto = fast with the bf16 conditioner, hybrid with fp32")
    ap.add_argument("--conditioner", default="auto", choices=["auto", "bf16", "fp32", "fp32x3"],
                    help="bf16 = 16-bit tensor-core operands; fp32 = 3xTF32 GEMMs + stand-alone spline; "
                         "fp32x3 = split-operand fused layer kernel (fp32-class)")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunk-rows", type=int, default=262144, help="row chunk of the host-buffer (e2e) pipeline")
    ap.add_argument("--e2e-wait", type=int, default=0, help="1: order the current stream after every host-API call")
    ap.add_argument("--no-fused", action="store_true", help="run the layer as separate GEMM + spline kernels")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch rows per GPU; strong: --batch rows in total, split over the ranks")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block (N > 1)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the H=800 / strict-fp32 secondary blocks")
    ap.add_argument("--no-train", action="store_true", help="skip the training block")
    ap.add_argument("--no-hmc", action="store_true", help="skip the flow-preconditioned HMC block")
    ap.add_argument("--train-rows", type=int, default=65536, help="rows per GPU of the training block")
    return ap.parse_args()


def workload_name(a):
    return f"cfg2: RQS coupling flow d=64, 8 x NSF_CL(size=32,dim=2,K=8,B=3,H={a.hidden}), batch {a.batch}/GPU"


def specs():
    return [dict(type="NSF_CL", size=SIZE, dim=DIM, K=KBINS, B=TAIL, mask=[i % 2]) for i in range(LAYERS)]


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle (torch-CPU restatement of the reference) on the host cores
# ------------------------------------------------------------------------------------------
def make_state_dict(hidden):
    """Reference-initialised weights (nn.Linear default init under manual_seed(0)), built
    without touching the GPU so both arms share them."""
    import torch.nn as nn
    torch.manual_seed(0)
    sd = {}
    for i in range(LAYERS):
        net = nn.Sequential(nn.Linear(SIZE, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh(),
                            nn.Linear(hidden, 23 * SIZE))
        for k, v in net.state_dict().items():
            sd[f"flows.{i}.psi.network.{k}"] = v.detach().clone()
    return sd


def cpu_pass(sd, rows, repeats=1):
    from oracle import nf_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    x = torch.randn(rows, D, generator=torch.Generator().manual_seed(1))
    z = torch.randn(rows, D, generator=torch.Generator().manual_seed(2))
    sp = specs()
    best = float("inf")
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.flow_fwd_inv_pass(sp, sd, x, z)
            best = min(best, time.perf_counter() - t0)
    return best


def cpu_baseline(a, sd, budget_s=12.0):
    rows = a.cpu_rows
    if rows <= 0:
        probe_rows = 2048
        t = cpu_pass(sd, probe_rows)
        rows = int(max(2048, min(65536, probe_rows * budget_s / max(t, 1e-3))))
        rows = (rows // 1024) * 1024
    t = cpu_pass(sd, rows)
    return {"value": rows / t, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{rows} rows of the same workload (one fwd+inv pass of oracle/nf_oracle.py, torch CPU fp32, "
                      f"{torch.get_num_threads()} threads), {t:.2f} s"}


def parity_check(a, sd, model, hx, hz, dev, rows=2048):
    """The timed configuration checked against the oracle (CPU restatement of the reference, fp32) on
    the first rows of the timed inputs: max |a-b| / max(1, |b|) of z, x, the log-dets and the prior
    log-prob.  The oracle is the checker here, nothing of it is timed."""
    from oracle import nf_oracle as O
    x, z = hx[:rows].clone(), hz[:rows].clone()
    with torch.no_grad():
        (rz, rplp, rld), (rx, rldi) = O.flow_fwd_inv_pass(specs(), sd, x, z)
        gz, gplp, gld = model.forward(x.to(dev))
        gx, gldi = model.inverse(z.to(dev))

    def rel(g, r):
        return float(((g.double().cpu() - r.double()).abs() / r.double().abs().clamp_min(1.0)).max())
    # the gate proper is per layer on identical inputs: feed every layer the oracle's own
    # intermediate activation and compare that one layer
    # ... and, for the tolerance the north star states on the conditioner GEMMs themselves, the spline
    # parameters (output of the conditioner MLP as the layer's own GEMM kernels compute it, un-fused)
    # against the oracle's fp32 MLP on the same conditioning columns
    per_z, per_ld, per_p, cur = 0.0, 0.0, 0.0, x
    sp = specs()
    bins_total = bins_vs_ref = bins_vs_own = 0
    from normalizingflow_b200 import _fused
    with torch.no_grad():
        for i, layer in enumerate(model.flows):
            ro, rl = O.apply_layer(sp[i], sd, i, cur, False)
            lower = cur.view(rows, SIZE, DIM)[:, :, sp[i]["mask"][0]].contiguous()
            rp = O.fcnn(lower, sd, f"flows.{i}.psi.")
            if layer.fused and _fused.eligible(layer):
                # the fused kernel's debug instantiation returns the parameters each element saw and the bin it
                # used: compare the parameters with the fp32 conditioner, the bins with (a) the oracle's search on
                # THOSE parameters (must be identical) and (b) the reference's bins (differ only where an input
                # lies closer to a knot than the conditioner's parameter error)
                go, gl, gp, gb = _fused.run_debug(layer, cur.to(dev), False)
                _, _, own_b = O.nsf_cl_transform(cur, gp.cpu(), SIZE, DIM, sp[i]["mask"], KBINS, TAIL, False)
                _, _, ref_b = O.nsf_cl_transform(cur, rp.reshape(rows, SIZE, 23), SIZE, DIM, sp[i]["mask"], KBINS, TAIL, False)
                bins_total += gb.numel()
                bins_vs_own += int((gb.cpu().long() != own_b).sum())
                bins_vs_ref += int((gb.cpu().long() != ref_b).sum())
                gp = gp.reshape(rows, -1)
            else:
                go, gl = layer.forward(cur.to(dev))
                gp = layer.psi(lower.to(dev))
            per_z, per_ld = max(per_z, rel(go, ro)), max(per_ld, rel(gl, rl))
            per_p = max(per_p, rel(gp, rp))
            cur = ro
    bins = None
    if bins_total:
        bins = {"elements": bins_total, "differ_from_oracle_search_on_kernels_own_params": bins_vs_own,
                "differ_from_reference_bins": bins_vs_ref}
    return {"rows": rows, "against": "oracle/nf_oracle.py (fp32, host)",
            "per_layer_identical_inputs": {"z": per_z, "log_det": per_ld, "conditioner_output_spline_params": per_p},
            "bins": bins,
            "chain_of_8_layers": {"z": rel(gz, rz), "log_det_fwd": rel(gld, rld), "prior_logprob": rel(gplp, rplp),
                                  "x": rel(gx, rx), "log_det_inv": rel(gldi, rldi)},
            "class": "1e-5 per layer (fp32-class conditioner)" if a.conditioner in ("fp32", "fp32x3") else
            "16-bit tensor-core conditioner GEMMs (north star: 1e-2 class): fp16 operands in the fused layer kernel "
            "(hidden <= 128) and in the wide path's operand images; conditioner_output_spline_params is measured on "
            "the layer's un-fused GEMM kernels; log_det is a sum of 32 per-feature terms per layer"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sd = make_state_dict(a.hidden)
    rows = a.cpu_rows if a.cpu_rows > 0 else 8192
    for _ in range(a.warmup):
        cpu_pass(sd, min(rows, 1024))
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cpu_pass(sd, rows)
    dt = (time.perf_counter() - t0) / max(1, a.steps)
    val = rows / dt
    cores = os.cpu_count() or 1
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "hidden": a.hidden,
                       "note": "CPU arm: each step is a bounded sample of the workload"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{rows} rows per step, torch CPU fp32, {torch.get_num_threads()} threads"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------
from tools.benchlib import (ClockSampler, barrier as _barrier, conditioner_flops, hbm_roofline,      # noqa: E402
                            host_copy_ceiling, max_over_ranks, peaks, tensor_roofline)

LABELS = {"nsf_pairs_fused": "nsf_pairs_fused_kernel (conditioner GEMMs + RQS epilogue, one launch per layer pass)",
          "nsf_pairs_fused3x": "nsf_fused3x_kernel (split-operand fp32-class conditioner GEMMs + RQS epilogue, one launch per layer pass)",
          "rqs_coupling": "rqs_coupling_pairs (stand-alone spline transform)",
          "gemm_ws_l1": "gemm_ws_kernel<image> (input x hidden conditioner GEMM)",
          "gemm_ws_l2": "gemm_ws_kernel<image> (hidden x hidden conditioner GEMM, persistent tcgen05)",
          "gemm_ws_l3": "gemm_ws_kernel<fp32 rows> (last conditioner GEMM)",
          "gemm_ws_rqs": "gemm_ws_kernel<RQS epilogue> (last conditioner GEMM + spline transform)",
          "linear_tf32x3": "linear_tf32x3_kernel (3xTF32 conditioner GEMM)"}


def build_flow(hidden, cond, arith, fused, dev, sd):
    from normalizingflow_b200.flows import NSF_CL
    from normalizingflow_b200.models import GaussianPrior, NormalizingFlowModel
    flows = [NSF_CL(SIZE, dim=DIM, K=KBINS, B=TAIL, hidden_dim=hidden, mask=[i % 2], arith=arith) for i in range(LAYERS)]
    for f in flows:
        f.psi.precision = cond
        f.fused = fused
    model = NormalizingFlowModel(GaussianPrior(D, device=dev), flows, device=dev)
    model.load_state_dict(sd)
    return model.to(dev)


def time_device(model, x, z, steps, warmup, dev, world, local, rank, with_clocks=True, graphed=False):
    """(ms per step [max over ranks], per-kernel CUDA-event sums, libnfk launches, clocks) of `steps` fwd+inv passes.
    graphed: the step is captured once with graphs.GraphedCallable (public API) and replayed as ONE CUDA graph -- the
    per-kernel event timer is then empty, the launch count is what one replay contains."""
    from normalizingflow_b200 import _lib, _ops

    def eager_step():
        with torch.no_grad():
            a = model.forward(x)
            b = model.inverse(z)
        return a + b
    step, per_replay = eager_step, None
    if graphed:
        from normalizingflow_b200.graphs import GraphedCallable
        l0 = _lib.launch_count()
        eager_step()
        per_replay = _lib.launch_count() - l0
        g = GraphedCallable(lambda xx, zz: eager_step(), x, z)
        step = g.graph.replay
    for _ in range(max(3, warmup)):
        step()
    _barrier(world)
    clocks = ClockSampler(local).start() if (rank == 0 and with_clocks) else None
    timer = _ops.KernelTimer()
    _ops.KERNEL_TIMER = timer
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(world)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    _barrier(world)
    ms = e0.elapsed_time(e1)
    launches = (_lib.launch_count() - launches0) if per_replay is None else per_replay * steps
    _ops.KERNEL_TIMER = None
    ksum = timer.summary()
    clk = clocks.stop() if clocks is not None else None
    ms = max_over_ranks([ms], dev, world)[0]
    return ms / steps, ksum, launches, clk, ms


def roofline_of(ksum, N, hidden, total_ms, steps):
    """Roofline object of the dominant kernel (largest CUDA-event sum on the launching stream).  Transform /
    fused-layer kernels are quoted against the ALGORITHMIC bytes of the unfused transform (3,464 B
    per row per layer pass; a fused kernel that keeps the parameters on chip can read above 1.0); conditioner
    GEMMs against their algorithmic flops and the sustained bf16 tensor peak."""
    kname, (n_l, k_ms) = max(ksum.items(), key=lambda kv: kv[1][1]) if ksum else ("none", (0, 0.0))
    avg_ms = k_ms / max(1, n_l)
    H = hidden
    tensor_flops = {"gemm_ws_l1": 2.0 * N * SIZE * H, "gemm_ws_l2": 2.0 * N * H * H,
                    "gemm_ws_l3": 2.0 * N * H * 23 * SIZE, "gemm_ws_rqs": 2.0 * N * H * 23 * SIZE,
                    "linear_tf32x3": None}
    extra = {"avg_launch_ms": avg_ms, "launches_timed": n_l, "share_of_step": (k_ms / total_ms) if total_ms else None,
             "kernels_ms_per_step": {k: v[1] / steps for k, v in ksum.items()}}
    if tensor_flops.get(kname):
        r = tensor_roofline(LABELS.get(kname, kname), tensor_flops[kname], avg_ms, extra)
        r["algorithmic_flops_per_launch"] = tensor_flops[kname]
    else:
        r = hbm_roofline(LABELS.get(kname, kname), ROW_BYTES_PER_LAYER * N, avg_ms, extra)
        r["algorithmic_bytes_per_launch"] = ROW_BYTES_PER_LAYER * N
        by_design = {"nsf_pairs_fused": 64 * 4 * 2 + 8, "nsf_pairs_fused3x": 64 * 4 * 2 + 8, "rqs_coupling": ROW_BYTES_PER_LAYER}.get(kname)
        r["hbm_bytes_per_launch_by_design"] = by_design * N if by_design else None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            r["traffic"] = json.load(open(tp)).get(kname + "_bytes_per_launch")
        except Exception:
            pass
    # whole-step view: algorithmic bytes of the 16 layer passes and the conditioner flops against both peaks
    hbm, sus, _, _ = peaks()
    step_ms = total_ms / steps
    r["whole_step"] = {"algorithmic_GBps": 16 * ROW_BYTES_PER_LAYER * N / (step_ms * 1e-3) / 1e9,
                       "frac_of_hbm_peak": 16 * ROW_BYTES_PER_LAYER * N / (step_ms * 1e-3) / 1e9 / hbm,
                       "conditioner_TFLOPs": 16 * conditioner_flops(SIZE, DIM, 1, H) * N / (step_ms * 1e-3) / 1e12,
                       "frac_of_sustained_bf16_peak": 16 * conditioner_flops(SIZE, DIM, 1, H) * N / (step_ms * 1e-3) / 1e12 / sus}
    return r


def dtype_string(cond, hidden, fused):
    if cond == "fp32":
        return "f32 (3xTF32 tensor-core conditioner GEMMs, fp32 transforms)"
    if cond == "fp32x3":
        return "f32 transforms / split fp16 hi+lo operands (3 MMAs per product, fp32-class) conditioner GEMMs"
    if fused and hidden <= 128:
        return "f32 transforms / fp16-operand conditioner GEMMs (fp32 accumulate)"
    return "f32 transforms / fp16-operand conditioner GEMMs (fp32 accumulate), operand images in HBM"


def time_e2e(model, hx, hz, steps, chunk_rows, wait, dev, world, sample_instead=False):
    """End to end through the public host-buffer API: pinned host inputs, results back in pinned host memory."""
    N = hx.shape[0]
    out_lp = torch.empty(N, dtype=torch.float32).pin_memory()
    out_x = torch.empty(N, D, dtype=torch.float32).pin_memory()
    out_lpx = torch.empty(N, dtype=torch.float32).pin_memory()
    out_z = torch.empty(N, D, dtype=torch.float32).pin_memory() if sample_instead else None

    def e2e_step():
        model.evaluate_host(hx, out=out_lp, chunk_rows=chunk_rows, wait=wait)                       # log p(x)
        if sample_instead:
            model.sample_host(N, out_x=out_x, out_log_px=out_lpx, out_z=out_z, chunk_rows=chunk_rows, wait=wait)
        else:
            model.inverse_host(hz, out_x=out_x, out_log_px=out_lpx, chunk_rows=chunk_rows, wait=wait)   # sampling
    for _ in range(2):
        e2e_step()
    model.host_sync()
    _barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        e2e_step()
    model.host_sync()                      # every result of every step is in host memory
    e1.record()
    _barrier(world)
    wall = (time.perf_counter() - t0) * 1e3
    ems = max(e0.elapsed_time(e1), wall)
    ems = max_over_ranks([ems], dev, world)[0] / steps
    h2d = N * D * 4 if sample_instead else 2 * N * D * 4
    d2h = (2 * N * D * 4 + 2 * N * 4) if sample_instead else (N * D * 4 + 2 * N * 4)
    return {"value": world * N / (ems * 1e-3), "unit": UNIT, "ms_per_step": ems, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h}


def train_block(a, sd, dev, world, rank):
    """Short data-parallel training run of the cfg-2 flow (forward-KL step, applications/src/train.py:22-29): every
    rank owns its rows, gradients live in ONE flat bucket that is all-reduced in place over NCCL between backward()
    and the Adam step.  Reports the step and all-reduce times, and the two correctness checks of the sharded
    training path: replicas stay bit-identical, and the averaged sharded gradient equals the gradient of ONE
    process on the concatenated batch."""
    import torch.distributed as dist
    from normalizingflow_b200 import _lib
    from normalizingflow_b200.dist import GradBucket, allreduce_gradients, broadcast_parameters
    rows = a.train_rows
    model = build_flow(a.hidden, "bf16", "hybrid", True, dev, sd)
    broadcast_parameters(model)
    params = list(model.parameters())
    x = torch.randn(rows, D, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank))
    opt = torch.optim.Adam(params, lr=1e-4)
    bucket = GradBucket(params)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step(timed=False):
        zz, plp, ld = model(x)
        loss = -torch.mean(plp + ld)
        bucket.zero()
        loss.backward()
        if timed:
            ev[1].record()
        nb = bucket.allreduce()
        if timed:
            ev[2].record()
        opt.step()
        return nb
    for _ in range(4):
        step()
    _barrier(world)
    l0 = _lib.launch_count()
    t_ar, nbytes, steps = 0.0, 0, 5
    ev[0].record()
    for _ in range(steps):
        nbytes = step(True)
        ev[3].record()
        torch.cuda.synchronize()
        t_ar += ev[1].elapsed_time(ev[2])
    ms = ev[0].elapsed_time(ev[3]) / steps
    launches = (_lib.launch_count() - l0) / steps
    # the all-reduce alone: ranks aligned by a barrier, ten back-to-back reductions of the same bucket (inside a
    # step the measured interval also contains the wait for the slowest rank to reach the collective)
    ar_only = 0.0
    if world > 1:
        _barrier(world)
        ev[1].record()
        for _ in range(10):
            bucket.allreduce()
        ev[2].record()
        torch.cuda.synchronize()
        ar_only = ev[1].elapsed_time(ev[2]) / 10
    ms, ar, ar_only = max_over_ranks([ms, t_ar / steps, ar_only], dev, world)
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    ident, rel = True, None
    if world > 1:
        lo, hi = flat.clone(), flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ident = bool(torch.equal(lo, hi))
        xs = torch.randn(8192, D, device=dev, generator=torch.Generator(device=dev).manual_seed(900 + rank))
        zz, plp, ld = model(xs)
        opt.zero_grad(set_to_none=True)
        (-torch.mean(plp + ld)).backward()
        allreduce_gradients(params)
        g_dist = torch.cat([p.grad.reshape(-1) for p in params]).clone()
        gathered = [torch.empty_like(xs) for _ in range(world)]
        dist.all_gather(gathered, xs)
        if rank == 0:
            zz, plp, ld = model(torch.cat(gathered))
            opt.zero_grad(set_to_none=True)
            (-torch.mean(plp + ld)).backward()
            g_one = torch.cat([p.grad.reshape(-1) for p in params])
            rel = float((g_dist - g_one).abs().max() / g_one.abs().max().clamp_min(1e-30))
    out = {"what": f"forward-KL training step of the cfg-2 flow (H={a.hidden}, bf16 tensor-core forward+backward), {rows} rows/GPU, "
                   "Adam, one in-place NCCL all-reduce of the flat fp32 gradient bucket per step",
           "ms_per_step": ms, "train_samples_per_s": world * rows / (ms * 1e-3),
           "allreduce_ms_in_step_incl_rank_skew": ar, "allreduce_ms_alone": ar_only,
           "allreduce_GBps_alone": (nbytes / (ar_only * 1e-3) / 1e9) if ar_only else None,
           "allreduce_bytes": nbytes, "grad_elements": int(flat.numel()), "libnfk_launches_per_step": launches,
           "replicas_bit_identical_after_steps": ident, "sharded_vs_single_process_gradient_rel_err": rel}
    del model, opt, bucket, x
    torch.cuda.empty_cache()
    return out


def hmc_block(a, sd, dev, world, rank):
    """BASELINE config 5 (flow-preconditioned HMC, nf/hmc.py:34-41): 65,536 chains in total, split over the ranks, through
    the cfg-2 flow.  One leapfrog trajectory of path_len steps = path_len + 1 log-prob + gradient evaluations, replayed
    as one CUDA graph; every evaluation is 8 fused forward launches, the log-prob reduction and 8 one-launch layer
    backwards (csrc/nsf_fused_bwd.cu).  Reports the time per evaluation and the force parity against the multi-launch
    bf16 path and against autograd through the fp32 parity path."""
    from normalizingflow_b200 import _fused, _lib, _wide
    from normalizingflow_b200.dist import shard_rows
    from normalizingflow_b200.hmc import FlowSimulation
    chains, path_len, dt = 65536, 10, 0.01
    r0, r1 = shard_rows(chains, rank, world)
    C = r1 - r0
    model = build_flow(a.hidden, "bf16", "fast", True, dev, sd)
    q0 = 0.7 * torch.randn(C, D, device=dev, generator=torch.Generator(device=dev).manual_seed(50 + rank))
    sim = FlowSimulation(model, n_chains=C, nparticles=SIZE, dim=DIM, init_pos=q0)
    sim.set_velocity(torch.randn(C, D, device=dev, generator=torch.Generator(device=dev).manual_seed(60 + rank)))
    fused_path = _fused.flow_logp_and_grad(model, q0[:256]) is not None

    def traj():
        sim.integration_step(path_len=path_len, dt=dt)
    for _ in range(3):
        traj()
    _barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps, n0 = 5, sim.grad_evals
    e0.record()
    for _ in range(reps):
        traj()
    e1.record()
    torch.cuda.synchronize()
    evals = sim.grad_evals - n0
    (ms_eval,) = max_over_ranks([e0.elapsed_time(e1) / evals], dev, world)
    # launches and kernel time of ONE evaluation, eagerly
    sim.potential_and_force(q0)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0.record()
    U, F = sim.potential_and_force(q0)
    e1.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - l0
    out = {"what": f"flow-preconditioned HMC (BASELINE config 5): {chains} chains in total ({C} per GPU), d = 64, 8 x NSF_CL(H={a.hidden}), "
                   f"leapfrog path_len {path_len}, trajectory replayed as one CUDA graph",
           "one_launch_per_layer_path": bool(fused_path), "ms_per_logprob_grad_eval": ms_eval,
           "value": chains / (ms_eval * 1e-3), "unit": "chain log-prob+grad evaluations/s", "scaling": "strong",
           "libnfk_launches_per_eval": launches, "eager_single_eval_ms": e0.elapsed_time(e1)}
    # SURVEY 8(d) accounting of one evaluation (parameters counted as if they travelled: forward 8 x 3,464 B, backward
    # 8 x (params 2,944 + x 256 + grad_z 256 + grad_x 256 + grad_params 2,944 + 8) per chain)
    nbytes = (8 * 3464 + 8 * (2944 + 256 + 256 + 256 + 2944 + 8)) * C
    peak = peaks()[0]
    out["roofline"] = {"bound": "hbm", "kernel": "one log-prob + grad evaluation through the flow (81,024 B per chain, SURVEY 8(d))",
                       "achieved": nbytes / (ms_eval * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                       "frac": nbytes / (ms_eval * 1e-3) / 1e9 / peak,
                       "note": "the fused kernels keep the spline parameters and their gradients on chip: real traffic is "
                               "about 1.3 KB per chain and layer, and the evaluation is bound by the adjoint arithmetic "
                               "(issue slots / MUFU), not by HBM"}
    # the replayed graph (launches really overlapping, tile by tile) must reproduce the launch-by-launch trajectory bit for
    # bit: (a) separate kick / drift launches, tile flags inside evaluations vs none; (b) the trajectory as ONE chain
    # (leapfrog folded into the last backward launch, evaluations hanging on each other) vs the fold launch by launch
    def small_traj(use_graph, chain, fold, chain_evals):
        _fused.TILE_CHAIN, _fused.CHAIN_EVALS = ("always" if chain else False), chain_evals
        try:
            s2 = FlowSimulation(model, n_chains=min(C, 32768), nparticles=SIZE, dim=DIM, init_pos=q0[:32768])
            s2.use_graph, s2.fused_leapfrog = use_graph, fold
            s2.set_velocity(torch.ones_like(s2.position) * 0.3)
            qa, ua = s2.integration_step(path_len=3, dt=dt)
            qb, ub = s2.integration_step(path_len=3, dt=dt)
            torch.cuda.synchronize()
            return qa.clone(), ua.clone(), qb.clone(), ub.clone()
        finally:
            _fused.TILE_CHAIN, _fused.CHAIN_EVALS = True, True
    same = lambda a_, b_: bool(all(torch.equal(x_, y_) for x_, y_ in zip(a_, b_)))          # noqa: E731
    out["graph_replay_with_tile_flags_equals_launch_by_launch_bitwise"] = same(small_traj(False, False, False, False),
                                                                               small_traj(True, True, False, False))
    out["trajectory_as_one_chain_equals_launch_by_launch_bitwise"] = same(small_traj(False, True, True, False),
                                                                          small_traj(True, True, True, True))
    out["leapfrog_folded_into_last_backward_launch"] = bool(sim.fused_leapfrog and sim._fused_leapfrog_ok(sim.position))
    if rank == 0:
        sub = q0[:2048]
        lp_f, g_f = _fused.flow_logp_and_grad(model, sub) if fused_path else _wide.flow_logp_and_grad(model, sub)
        ref = build_flow(a.hidden, "fp32", "hybrid", False, dev, sd)          # fp32 parity path, torch autograd
        for p_ in ref.parameters():
            p_.requires_grad_(False)
        xs = sub.clone().requires_grad_(True)
        zz, plp, ld = ref(xs)
        (g_r,) = torch.autograd.grad((plp + ld).sum(), xs)
        lp_r = (plp + ld).detach()
        err = (g_f - g_r).abs() / (1.0 + g_r.abs())
        out["parity"] = {"rows": int(sub.shape[0]), "against": "torch autograd through the fp32 parity path of the same weights",
                         "force_median_rel": float(err.median()), "force_p99_rel": float(err.flatten().kthvalue(int(0.99 * err.numel())).values),
                         "force_frac_gt_3e-2": float((err > 3e-2).float().mean()),
                         "logp_max_rel": float(((lp_f - lp_r).abs() / lp_r.abs().clamp_min(1.0)).max()),
                         "note": "16-bit conditioner class (1e-2); d log|det| / dx jumps across knots (the spline is C1) and the "
                                 "16-bit conditioner moves the knots by ~1e-3, so a few elements next to a knot see the neighbouring "
                                 "bin's second derivative: the bulk is gated, not the maximum (tests/test_gpu_hmc.py)"}
        del ref
    del model, sim
    torch.cuda.empty_cache()
    return out


def run_native(a):
    import torch.distributed as dist
    from normalizingflow_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from normalizingflow_b200.dist import bind_to_gpu_numa_node, shard_rows
    numa = bind_to_gpu_numa_node(local)          # before any pinned allocation (first touch = local node)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cond = a.conditioner
    if cond == "auto":
        cond = "bf16" if _lib.have("nfk_linear_bf16") else "fp32"
    a.conditioner = cond
    if a.arith == "auto":
        # 16-bit GEMMs already move the spline parameters by ~1e-3: the FAST spline arithmetic (bins still the
        # exact search on the kernel's own parameters) is inside that class; fp32-class conditioners get HYBRID
        a.arith = "fast" if cond == "bf16" else "hybrid"
    sd = make_state_dict(a.hidden)
    model = build_flow(a.hidden, cond, a.arith, not a.no_fused, dev, sd)

    strong = a.scaling == "strong"
    N = a.batch
    if strong:                                    # the global batch of a.batch rows is split over the ranks
        r0, r1 = shard_rows(a.batch, rank, world)
        N = r1 - r0
    # each rank owns its rows of the global batch: seeded per rank
    gx = torch.Generator().manual_seed(1 + 1000 * rank)
    gz = torch.Generator().manual_seed(2 + 1000 * rank)
    hx = torch.randn(N, D, generator=gx).pin_memory()
    hz = torch.randn(N, D, generator=gz).pin_memory()
    x = hx.to(dev)
    z = hz.to(dev)

    ms_per_step, ksum, launches, clk, total_ms = time_device(model, x, z, a.steps, a.warmup, dev, world, local, rank)
    n_global = a.batch if strong else world * N
    value = n_global / (ms_per_step * 1e-3)

    # ---- end to end through the public API with HOST buffers
    e2e = e2e_sample = ceiling = None
    if not a.no_e2e:
        e2e = time_e2e(model, hx, hz, a.steps, a.chunk_rows, bool(a.e2e_wait), dev, world)
        e2e["value"] = n_global / (e2e["ms_per_step"] * 1e-3)
        e2e["api"] = (f"NormalizingFlowModel.evaluate_host(x) -> log p(x) [N] (z stays on the device, as evaluate()) + "
                      f"inverse_host(z) -> (x, log_px): pinned host buffers, {a.chunk_rows}-row chunks, H2D / kernels / D2H "
                      "on three streams, the kernels of a chunk replayed as one CUDA graph")
        e2e_sample = time_e2e(model, hx, hz, max(2, a.steps // 2), a.chunk_rows, bool(a.e2e_wait), dev, world, sample_instead=True)
        e2e_sample["value"] = n_global / (e2e_sample["ms_per_step"] * 1e-3)
        e2e_sample["api"] = ("evaluate_host(x) + sample_host(n) -> (x, log_px, z): latents drawn on the device as the "
                             "reference's sample(n) does (nf/models.py:31-35), no H2D for the sampling half")
        # what the host memory system of THIS box gives `world` ranks moving the same bytes with no kernels at all
        ceiling = host_copy_ceiling(dev, world, e2e["h2d_bytes_per_step"], e2e["d2h_bytes_per_step"], a.chunk_rows * D * 4)
        ceiling["e2e_ms_per_step"] = e2e["ms_per_step"]
        ceiling["e2e_frac_of_copy_ceiling"] = ceiling["seconds_per_step_bytes"] * 1e3 / e2e["ms_per_step"]
        ceiling["device_ms_per_step"] = ms_per_step
        ceiling["note"] = ("an e2e step cannot be faster than max(device time, copy-ceiling time); "
                           "e2e_frac_of_best_possible relates it to that bound")
        ceiling["e2e_frac_of_best_possible"] = max(ms_per_step, ceiling["seconds_per_step_bytes"] * 1e3) / e2e["ms_per_step"]

    # ---- strong scaling of the 2^20-row global batch (north star: "split the batch across the 8 GPUs")
    strong_blk = None
    if not strong and world > 1 and not a.no_strong:
        r0, r1 = shard_rows(a.batch, rank, world)
        xs_, zs_ = x[: r1 - r0], z[: r1 - r0]
        # the per-GPU share is small (2^20 / N rows: ~90 us per layer launch at N = 8), so the step is replayed as one
        # CUDA graph (graphs.GraphedCallable), as the host-buffer pipeline does for its chunks
        s_ms, _, _, _, s_tot = time_device(model, xs_, zs_, a.steps, a.warmup, dev, world, local, rank, with_clocks=False,
                                           graphed=True)
        _, s_ksum, _, _, _ = time_device(model, xs_, zs_, 3, 1, dev, world, local, rank, with_clocks=False)
        kn, (nl, kms) = max(s_ksum.items(), key=lambda kv: kv[1][1])
        strong_blk = {"scaling": "strong", "global_batch": a.batch, "rows_per_gpu": r1 - r0, "ms_per_step": s_ms,
                      "value": a.batch / (s_ms * 1e-3), "unit": UNIT,
                      "speedup_vs_one_gpu_same_run": (a.batch / (s_ms * 1e-3)) / (value / world),
                      "efficiency": (a.batch / (s_ms * 1e-3)) / value,
                      "one_gpu_value_same_run": value / world,
                      "dominant_kernel_avg_launch_ms": kms / max(1, nl),
                      "how": "one CUDA-graph replay per step (16 layer launches + log-prob reduction), consecutive layer "
                             "kernels chained by programmatic dependent launch",
                      "note": "every rank runs its contiguous 1/N of the 2^20-row batch, no data-path collective; "
                              "one_gpu_value is this run's weak-scaling per-GPU rate on the full 2^20 rows"}
        if not a.no_e2e:
            se = time_e2e(model, hx[: r1 - r0], hz[: r1 - r0], a.steps, min(a.chunk_rows, max(32768, (r1 - r0) // 2)),
                          bool(a.e2e_wait), dev, world)
            strong_blk["e2e"] = {"value": a.batch / (se["ms_per_step"] * 1e-3), "ms_per_step": se["ms_per_step"]}

    # ---- secondary configurations of the same workload (class-default hidden width; strict fp32 parity)
    secondary = []
    if not a.no_secondary and a.hidden == 128 and cond == "bf16" and not a.no_fused:
        specs2 = [("cfg2_h800_class_default", 800, "bf16", "fast", "tensor-pipe-bound regime (42.09 TFLOP per 2^20-row step)")]
        if _lib.have("nfk_nsf_pairs_fused2"):
            specs2.append(("cfg2_h128_fp32_class_fused", 128, "fp32x3", "hybrid",
                           "strict parity: split-operand (fp32-class) conditioner GEMMs + HYBRID spline (exact bin search) in ONE kernel per layer pass"))
        specs2.append(("cfg2_h128_strict_fp32_unfused", 128, "fp32", "hybrid",
                       "strict parity, unfused: 3xTF32 conditioner GEMMs + stand-alone HYBRID spline kernel"))
        for name, H2, c2, ar2, note in specs2:
            sd2 = sd if H2 == a.hidden else make_state_dict(H2)
            m2 = build_flow(H2, c2, ar2, True, dev, sd2)
            st2 = max(2, min(5, a.steps // 4))
            ms2, ks2, l2, clk2, tot2 = time_device(m2, x, z, st2, 3, dev, world, local, rank)
            blk = {"name": name, "hidden": H2, "conditioner": c2, "arith": ar2, "note": note, "steps": st2,
                   "ms_per_step": ms2, "value": world * N / (ms2 * 1e-3), "unit": UNIT, "dtype": dtype_string(c2, H2, True),
                   "roofline": roofline_of(ks2, N, H2, tot2, st2), "clocks": clk2, "gpu_launches": l2}
            if rank == 0 and world == 1 and not a.no_cpu_baseline:
                a2 = argparse.Namespace(**vars(a))
                a2.conditioner, a2.hidden = c2, H2
                blk["parity"] = parity_check(a2, sd2, m2, hx, hz, dev, rows=1024)
            secondary.append(blk)
            del m2
            torch.cuda.empty_cache()

    # ---- training with the NCCL gradient all-reduce (exercised under the driver at every N)
    train = None
    if not a.no_train and cond == "bf16":
        train = train_block(a, sd, dev, world, rank)

    # ---- flow-preconditioned HMC (config 5): log-prob + gradient through the flow
    hmc = None
    if not a.no_hmc and cond == "bf16" and a.hidden <= 128 and not a.no_fused:
        hmc = hmc_block(a, sd, dev, world, rank)

    if rank == 0:
        roofline = roofline_of(ksum, N, a.hidden, total_ms, a.steps)
        fused_kernel = bool(not a.no_fused and cond in ("bf16", "fp32x3") and a.hidden <= 128)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": max(3, a.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": dtype_string(cond, a.hidden, not a.no_fused), "data": "synthetic",
                "config": {"workload": workload_name(a), "hidden": a.hidden, "arith": a.arith,
                           "conditioner": {"bf16": "16-bit tensor-core operands (fp16 on the inference forward, bf16 for gradients)",
                                           "fp32": "fp32 (3xTF32)", "fp32x3": "fp32-class (split fp16 hi+lo operands)"}[cond],
                           "fused_layer_kernel": fused_kernel,
                           "spline_epilogue_on_last_gemm": bool(not a.no_fused and cond == "bf16" and a.hidden > 128),
                           "global_batch": n_global, "rows_per_gpu": N, "parallelism": f"batch-sharded x{world}",
                           "l2": "inputs larger than L2 (x 268 MB per 2^20 rows, 126 MB L2)", "numa": numa},
                "roofline": roofline, "clocks": clk, "gpu_launches": launches, "e2e": e2e,
                "e2e_sample_host": e2e_sample, "host_copy_ceiling": ceiling, "strong": strong_blk,
                "secondary": secondary, "train": train, "hmc": hmc}
        if not a.no_cpu_baseline and world == 1:       # reported at N=1 only
            line["cpu_baseline"] = cpu_baseline(a, sd)
            line["parity"] = parity_check(a, sd, model, hx, hz, dev)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)


if __name__ == "__main__":
    main()
