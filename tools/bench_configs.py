#!/usr/bin/env python
"""The other BASELINE.json configurations (1, 3, 4, 5) -- the parity cases of bench.py's headline workload -- measured
with CUDA events on 1, 2, 4 or 8 GPUs of one box.  One JSON line per configuration, each with `value`, the roofline
object of SURVEY.md 8(d) (algorithmic bytes or flops / time / measured peak) and the SM clocks sampled during the
timed region.

    python tools/bench_configs.py [--quick] [--only 3 4 5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
        tools/bench_configs.py

Multi-GPU: one process per GPU; every configuration keeps its GLOBAL batch (strong scaling: 2^20 rows for configs 3
and 4, 65,536 chains for config 5) and each rank owns a contiguous 1/N of it; no data-path collective, the training
steps of config 4 add the flat-bucket NCCL gradient all-reduce.  Times are the MAX over ranks.
"""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200 import _lib                                                                  # noqa: E402
from normalizingflow_b200.dist import GradBucket, broadcast_parameters, shard_rows                     # noqa: E402
from normalizingflow_b200.flows import NSF_AR, NSF_CL, Planar, Radial, RealNVP                         # noqa: E402
from normalizingflow_b200.hmc import HMC, FlowSimulation                                               # noqa: E402
from normalizingflow_b200.models import GaussianPrior, NormalizingFlowModel                            # noqa: E402
from tools.benchlib import (ClockSampler, barrier, conditioner_flops, hbm_roofline, max_over_ranks,    # noqa: E402
                            peaks, rqs_layer_bytes, tensor_roofline, world_info)

WORLD, RANK, LOCAL = world_info()
torch.cuda.set_device(LOCAL)
dev = torch.device("cuda", LOCAL)


def timeit(fn, iters=5, warm=2):
    """ms per call, max over ranks, bracketed by barriers"""
    for _ in range(warm):
        fn()
    barrier(WORLD)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    barrier(WORLD)
    return max_over_ranks([e0.elapsed_time(e1) / iters], dev, WORLD)[0]


def my_rows(n_global):
    a, b = shard_rows(n_global, RANK, WORLD)
    return b - a


def double_well_logp(x):      # U(x,y) = x^4/4 - 3x^2 + x + y^2/2  (absent in the reference)
    return -(0.25 * x[:, 0] ** 4 - 3 * x[:, 0] ** 2 + x[:, 0] + 0.5 * x[:, 1] ** 2)


def cfg1(H):
    """replicas only: a 4,096-row batch of a 2-D flow does not shard usefully; every rank runs the same step"""
    torch.manual_seed(0)
    m = NormalizingFlowModel(GaussianPrior(2, device=dev), [RealNVP(2, hidden_dim=H) for _ in range(8)], device=dev).to(dev)
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)

    def train_step():                                   # reverse KL (applications/src/setup.py:90-94)
        z = m.prior.sample((4096,))
        x, ld = m.inverse(z)
        loss = (m.prior.log_prob(z) - ld - double_well_logp(x)).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    t_train = timeit(train_step)
    t_sample = timeit(lambda: m.sample(4096))
    from normalizingflow_b200.graphs import GraphedCallable, GraphedTrainStep
    opt_g = torch.optim.Adam(m.parameters(), lr=1e-4, capturable=True)

    def loss_fn():
        z = m.prior.sample((4096,))
        x, ld = m.inverse(z)
        return (m.prior.log_prob(z) - ld - double_well_logp(x)).mean()
    gstep = GraphedTrainStep(loss_fn, opt_g)
    t_train_g = timeit(gstep)
    gsample = GraphedCallable(lambda z: m.inverse(z)[0], m.prior.sample((4096,)))
    zz = m.prior.sample((4096,))
    t_sample_g = timeit(lambda: gsample(zz))
    # 4 conditioners per module, 1 -> H -> H -> 1; forward + backward (dgrad + wgrad) = 3 x forward flops
    flops = 3 * 8 * 4 * 2.0 * (H + H * H + H) * 4096
    return {"config": f"1: 8 x RealNVP(2, H={H}), batch 4096, reverse-KL on a 2-D double well (replicas only: does not shard)",
            "metric": "train_samples_per_s", "value": 4096 / t_train_g * 1e3, "unit": "samples/s", "scaling": "replicas",
            "train_step_ms": t_train, "sample_ms": t_sample, "graphed_train_step_ms": t_train_g, "graphed_sample_ms": t_sample_g,
            "roofline": tensor_roofline("whole graphed training step (launch-latency-bound at this size)", flops, t_train_g)}


def cfg3(n_global, H=800, precision="bf16"):
    torch.manual_seed(0)
    cyc = [[0], [1], [2], [0, 1], [1, 2], [0, 2]]
    masks = (cyc * 2)[:8]
    fl = [NSF_CL(38, dim=3, K=8, B=4.0, hidden_dim=H, mask=mk) for mk in masks]
    for f in fl:
        f.psi.precision = precision
    m = NormalizingFlowModel(GaussianPrior(114, device=dev), fl, device=dev).to(dev)
    n = my_rows(n_global)
    x = torch.randn(n, 114, device=dev, generator=torch.Generator(device=dev).manual_seed(1 + RANK))
    t_eval = timeit(lambda: m.evaluate(x), iters=3, warm=2)
    t_samp = timeit(lambda: m.sample(n), iters=3, warm=1)
    flops = sum(conditioner_flops(38, 3, len(mk), H) for mk in masks) * n_global
    nbytes = sum(rqs_layer_bytes(38, 3, len(mk)) for mk in masks) * n_global
    return {"config": f"3: LJ-38 d=114, 8 x NSF_CL(38, dim=3, K=8, B=4, H={H}, 16-bit conditioner), global batch {n_global}",
            "metric": "evaluate_samples_per_s", "value": n_global / t_eval * 1e3, "unit": "samples/s", "scaling": "strong",
            "rows_per_gpu": n, "evaluate_ms": t_eval, "sample_ms": t_samp, "sample_samples_per_s": n_global / t_samp * 1e3,
            "roofline": tensor_roofline("whole evaluate() call: conditioner GEMMs (tensor-bound at H=800)", flops / WORLD, t_eval),
            "roofline_hbm_view": hbm_roofline("whole evaluate() call against the unfused transforms' algorithmic bytes (52,808 B/sample)",
                                              nbytes / WORLD, t_eval)}


def cfg4(n_global):
    torch.manual_seed(0)
    out = {}
    n = my_rows(n_global)
    x = torch.randn(n, 128, device=dev, generator=torch.Generator(device=dev).manual_seed(1 + RANK))
    for name, mk in (("planar", lambda: Planar(128)), ("radial_batch_global", lambda: Radial(128)),
                     ("radial_per_sample", lambda: Radial(128, per_sample=True))):
        torch.manual_seed(0)
        m = NormalizingFlowModel(GaussianPrior(128, device=dev), [mk() for _ in range(32)], device=dev).to(dev)
        broadcast_parameters(m)
        params = list(m.parameters())
        opt = torch.optim.Adam(params, lr=1e-4)
        bucket = GradBucket(params)
        t_eval = timeit(lambda: m.evaluate(x), iters=5)

        def train_step():
            z, plp, ld = m.forward(x)
            loss = -(plp + ld).mean()
            bucket.zero()
            loss.backward()
            bucket.allreduce()
            opt.step()
        t_train = timeit(train_step, iters=3)
        per_layer = 32 * (2 * 128 * 4 + 8) * n            # 1,032 B per row per layer
        fused_floor = (2 * 128 * 4 + 4) * n
        passes = {"planar": 1.0, "radial_per_sample": 1.0, "radial_batch_global": 32.5}[name]
        out[name] = {"evaluate_ms": t_eval, "train_step_ms": t_train, "evaluate_samples_per_s": n_global / t_eval * 1e3,
                     "train_samples_per_s": n_global / t_train * 1e3, "grad_allreduce_bytes": bucket.nbytes if WORLD > 1 else 0,
                     "roofline": hbm_roofline(f"{name}: fused 32-layer run, against the per-layer algorithmic bytes "
                                              "(> 1.0 means the fusion removed HBM passes)", per_layer, t_eval,
                                              {"frac_of_own_minimum_traffic": passes * (2 * 128 * 4) * n / (t_eval * 1e-3) / 1e9 / peaks()[0],
                                               "minimum_passes_over_the_batch": passes,
                                               "stack_fused_floor_bytes": fused_floor})}
    out.update({"config": f"4: 32 x Planar(128) / 32 x Radial(128) (reference batch-global norm and per-sample), global batch "
                          f"{n_global}, density eval + forward-KL train step (Adam, NCCL gradient all-reduce when N > 1)",
                "metric": "planar_evaluate_samples_per_s", "value": out["planar"]["evaluate_samples_per_s"], "unit": "samples/s",
                "scaling": "strong", "rows_per_gpu": n, "roofline": out["planar"]["roofline"]})
    return out


def cfg5(c_global, H, precision):
    torch.manual_seed(0)
    fl = [NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=H, mask=[i % 2]) for i in range(8)]
    for f in fl:
        f.psi.precision = precision
        if precision == "bf16":
            f.arith = "fast"       # as bench.py: inside the 16-bit conditioner's class, bins still the exact search
    m = NormalizingFlowModel(GaussianPrior(64, device=dev), fl, device=dev).to(dev)
    C = my_rows(c_global)
    sim = FlowSimulation(m, n_chains=C, nparticles=32, dim=2, generator=torch.Generator(device=dev).manual_seed(5 + RANK))
    h = HMC(sim, path_len=10, dt=0.05, dim=2, beta=1.0)
    h.hmc(epochs=3)                      # warm-up: lazy kernel loading, allocator growth, graph capture
    barrier(WORLD)
    n0 = sim.grad_evals
    t0 = time.perf_counter()
    pos, pot, logp, acc = h.hmc(epochs=8)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    evals = sim.grad_evals - n0
    ms_eval = max_over_ranks([dt / evals * 1e3], dev, WORLD)[0]
    # forward 8 x 3,464 + backward 8 x (params 2,944 + x 256 + grad_z 256 + grad_x 256 + grad_params 2,944 + 8)
    nbytes = (8 * 3464 + 8 * (2944 + 256 + 256 + 256 + 2944 + 8)) * C
    return {"config": f"5: flow-preconditioned HMC, {c_global} chains, d=64, 8 x NSF_CL(H={H}, {precision}), path_len 10, dt 0.05",
            "metric": "chain_logprob_grad_evals_per_s", "value": c_global / ms_eval * 1e3, "unit": "chain evaluations/s",
            "scaling": "strong", "chains_per_gpu": C, "ms_per_logprob_grad_eval": ms_eval, "accept_rate": acc,
            "roofline": hbm_roofline("one log-prob + grad evaluation through the flow (about 81 KB per chain, SURVEY 8(d))",
                                     nbytes, ms_eval)}


def cfg_train(N, H, precision):
    """forward-KL training step (applications/src/train.py:22-29) of the cfg-2 flow."""
    torch.manual_seed(0)
    fl = [NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=H, mask=[i % 2]) for i in range(8)]
    for f in fl:
        f.psi.precision = precision
    m = NormalizingFlowModel(GaussianPrior(64, device=dev), fl, device=dev).to(dev)
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    x = torch.randn(N, 64, device=dev)

    def step():
        z, plp, ld = m(x)
        loss = -torch.mean(plp + ld)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    t = timeit(step, iters=3, warm=2)
    flops = 3 * 8 * conditioner_flops(32, 2, 1, H) * N
    return {"config": f"A11: forward-KL train step, 8 x NSF_CL(32, dim 2, K 8, H={H}, {precision} conditioner), batch {N}",
            "metric": "train_samples_per_s", "value": N / t * 1e3, "unit": "samples/s", "train_step_ms": t,
            "roofline": tensor_roofline("whole step: forward + dgrad + wgrad conditioner GEMMs", flops, t)}


def cfg_targets(N):
    """the prior / target next to the flow in the shipped LJ experiment (applications/input/LJ.yaml)."""
    from normalizingflow_b200 import systems
    torch.manual_seed(0)
    centers = (torch.rand(32, 3) - 0.5) * 3.0
    ec = systems.EinsteinCrystal(centers, dim=3, boxlength=3.4, alpha=1000)
    lj = systems.LJ(boxlength=3.4, cutoff=1.6, shift=True)
    x = ec.sample(N)
    t_ec = timeit(lambda: ec.log_prob(x))
    t_lj = timeit(lambda: lj.potential(x.reshape(N, 32, 3)))
    return {"config": f"N2: EinsteinCrystal.log_prob + LJ.potential, 32 particles (d=96), batch {N}",
            "metric": "einstein_logprob_samples_per_s", "value": N / t_ec * 1e3, "unit": "samples/s",
            "einstein_logprob_ms": t_ec, "lj_potential_ms": t_lj, "lj_pairs_per_s": N * 32 * 32 / t_lj * 1e3,
            "roofline": hbm_roofline("einstein_logprob_kernel", (96 * 4 + 4) * N, t_ec)}


def cfg_nsf_ar(N, precision):
    """the shipped LJ experiment (applications/input/LJ.yaml): 2 x NSF_AR(96, K = 32, hidden 354)."""
    torch.manual_seed(0)
    fl = [NSF_AR(96, K=32, B=3.0, hidden_dim=354) for _ in range(2)]
    for f in fl:
        for l in f.layers:
            l.precision = precision
    m = NormalizingFlowModel(GaussianPrior(96, device=dev), fl, device=dev).to(dev)
    x = torch.randn(N, 96, device=dev)
    t_eval = timeit(lambda: m.evaluate(x), iters=3, warm=1)
    flops = 2 * sum(2.0 * (2 * i * 354 + 354 * 354 + 354 * 95) for i in range(1, 96)) * N
    return {"config": f"N1: LJ.yaml flow, 2 x NSF_AR(96, K=32, H=354, {precision} conditioners), batch {N}",
            "metric": "evaluate_samples_per_s", "value": N / t_eval * 1e3, "unit": "samples/s", "evaluate_ms": t_eval,
            "roofline": tensor_roofline("whole evaluate(): 190 conditioner MLPs as grouped GEMMs", flops, t_eval)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", nargs="*", default=None, help="subset of: 1 3 4 5 5fp32 train ar targets")
    a = ap.parse_args()
    if WORLD > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n3 = (1 << 18) if a.quick else (1 << 20)
    n4 = (1 << 18) if a.quick else (1 << 20)
    jobs = [("1", lambda: cfg1(100)), ("1", lambda: cfg1(800)), ("3", lambda: cfg3(n3)), ("4", lambda: cfg4(n4)),
            ("5", lambda: cfg5(65536, 128, "bf16")), ("5", lambda: cfg5(65536, 800, "bf16"))]
    if WORLD == 1:
        jobs += [("5fp32", lambda: cfg5(65536, 128, "fp32")), ("ar", lambda: cfg_nsf_ar(65536, "bf16")),
                 ("train", lambda: cfg_train(262144, 128, "bf16")), ("train", lambda: cfg_train(262144, 800, "bf16")),
                 ("targets", lambda: cfg_targets(1 << 20))]
    for tag, fn in jobs:
        if a.only is not None and tag not in a.only:
            continue
        clocks = ClockSampler(LOCAL).start() if RANK == 0 else None
        l0 = _lib.launch_count()
        r = fn()
        r["libnfk_launches"] = _lib.launch_count() - l0
        r["n_gpus"] = WORLD
        if clocks is not None:
            r["clocks"] = clocks.stop()
            print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
    if WORLD > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
