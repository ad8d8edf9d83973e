#!/usr/bin/env python
"""Timings of the other BASELINE.json configurations (1, 3, 4, 5) on one B200 — the parity cases of
bench.py's headline workload, measured with CUDA events.  Prints one JSON object per config.

    python tools/bench_configs.py [--quick]
"""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200 import _lib
from normalizingflow_b200.flows import NSF_AR, NSF_CL, Planar, Radial, RealNVP
from normalizingflow_b200.hmc import HMC, FlowSimulation
from normalizingflow_b200.models import GaussianPrior, NormalizingFlowModel

dev = torch.device("cuda:0")


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def double_well_logp(x):      # U(x,y) = x^4/4 - 3x^2 + x + y^2/2  (absent in the reference)
    return -(0.25 * x[:, 0] ** 4 - 3 * x[:, 0] ** 2 + x[:, 0] + 0.5 * x[:, 1] ** 2)


def cfg1(H):
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
    # the same step / sampling call captured once as a CUDA graph (launch-latency-bound at this size)
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
    return {"config": f"1: 8 x RealNVP(2, H={H}), batch 4096, reverse-KL on a 2-D double well",
            "train_step_ms": t_train, "sample_ms": t_sample, "train_samples_per_s": 4096 / t_train * 1e3,
            "graphed_train_step_ms": t_train_g, "graphed_sample_ms": t_sample_g,
            "graphed_train_samples_per_s": 4096 / t_train_g * 1e3}


def cfg3(N, H=800, precision="bf16"):
    torch.manual_seed(0)
    cyc = [[0], [1], [2], [0, 1], [1, 2], [0, 2]]
    masks = (cyc * 2)[:8]
    fl = [NSF_CL(38, dim=3, K=8, B=4.0, hidden_dim=H, mask=mk) for mk in masks]
    for f in fl:
        f.psi.precision = precision
    m = NormalizingFlowModel(GaussianPrior(114, device=dev), fl, device=dev).to(dev)
    x = torch.randn(N, 114, device=dev)
    t_eval = timeit(lambda: m.evaluate(x), iters=3, warm=1)
    t_samp = timeit(lambda: m.sample(N), iters=3, warm=1)
    return {"config": f"3: LJ-38 d=114, 8 x NSF_CL(38, dim=3, K=8, B=4, H={H}, {precision} conditioner), batch {N}",
            "evaluate_ms": t_eval, "sample_ms": t_samp, "evaluate_samples_per_s": N / t_eval * 1e3,
            "sample_samples_per_s": N / t_samp * 1e3}


def cfg4(N):
    torch.manual_seed(0)
    out = {}
    for name, mk in (("planar", lambda: Planar(128)), ("radial", lambda: Radial(128, per_sample=True))):
        m = NormalizingFlowModel(GaussianPrior(128, device=dev), [mk() for _ in range(32)], device=dev).to(dev)
        x = torch.randn(N, 128, device=dev)
        opt = torch.optim.Adam(m.parameters(), lr=1e-4)
        t_eval = timeit(lambda: m.evaluate(x), iters=5)

        def train_step():
            z, plp, ld = m.forward(x)
            loss = -(plp + ld).mean()
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        t_train = timeit(train_step, iters=3)
        out[name] = {"evaluate_ms": t_eval, "train_step_ms": t_train, "evaluate_samples_per_s": N / t_eval * 1e3,
                     "train_samples_per_s": N / t_train * 1e3,
                     "evaluate_GBps_vs_per_layer_bytes": 32 * (2 * 128 * 4 + 8) * N / t_eval / 1e6}
    out["config"] = f"4: 32 x Planar(128) / 32 x Radial(128, per_sample), batch {N}, density eval + forward-KL train step"
    return out


def cfg5(C, H, precision):
    torch.manual_seed(0)
    fl = [NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=H, mask=[i % 2]) for i in range(8)]
    for f in fl:
        f.psi.precision = precision
    m = NormalizingFlowModel(GaussianPrior(64, device=dev), fl, device=dev).to(dev)
    sim = FlowSimulation(m, n_chains=C, nparticles=32, dim=2)
    h = HMC(sim, path_len=10, dt=0.05, dim=2, beta=1.0)
    h.hmc(epochs=3)                      # warm-up: lazy kernel loading, allocator growth, graph capture
    torch.cuda.synchronize()
    n0 = sim.grad_evals
    t0 = time.perf_counter()
    pos, pot, logp, acc = h.hmc(epochs=8)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    evals = sim.grad_evals - n0
    return {"config": f"5: flow-preconditioned HMC, {C} chains, d=64, 8 x NSF_CL(H={H}, {precision}), path_len 10, dt 0.05",
            "ms_per_logprob_grad_eval": dt / evals * 1e3, "chain_grad_evals_per_s": C * evals / dt, "accept_rate": acc}


def cfg_train(N, H, precision):
    """forward-KL training step (applications/src/train.py:22-29) of the cfg-2 flow:
    loss = -mean(prior_logprob + log_det), backward through every layer, Adam."""
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
    return {"config": f"A11: forward-KL train step, 8 x NSF_CL(32, dim 2, K 8, H={H}, {precision} conditioner), batch {N}",
            "train_step_ms": t, "train_samples_per_s": N / t * 1e3}


def cfg_targets(N):
    """the prior / target next to the flow in the shipped LJ experiment
    (applications/input/LJ.yaml: 32 particles, Einstein-crystal prior alpha 1000, LJ target)."""
    from normalizingflow_b200 import systems
    torch.manual_seed(0)
    centers = (torch.rand(32, 3) - 0.5) * 3.0
    ec = systems.EinsteinCrystal(centers, dim=3, boxlength=3.4, alpha=1000)
    lj = systems.LJ(boxlength=3.4, cutoff=1.6, shift=True)
    x = ec.sample(N)
    t_ec = timeit(lambda: ec.log_prob(x))
    t_lj = timeit(lambda: lj.potential(x.reshape(N, 32, 3)))
    return {"config": f"N2: EinsteinCrystal.log_prob + LJ.potential, 32 particles (d=96), batch {N}",
            "einstein_logprob_ms": t_ec, "einstein_GBps": N * 96 * 4 / t_ec / 1e6, "lj_potential_ms": t_lj,
            "lj_pairs_per_s": N * 32 * 32 / t_lj * 1e3}


def cfg_nsf_ar(N, precision):
    """the shipped LJ experiment (applications/input/LJ.yaml): 2 x NSF_AR(dim = 32*3,
    K = 32 splines, hidden 354); B = 3."""
    torch.manual_seed(0)
    fl = [NSF_AR(96, K=32, B=3.0, hidden_dim=354) for _ in range(2)]
    for f in fl:
        for l in f.layers:
            l.precision = precision
    m = NormalizingFlowModel(GaussianPrior(96, device=dev), fl, device=dev).to(dev)
    x = torch.randn(N, 96, device=dev)
    t_eval = timeit(lambda: m.evaluate(x), iters=3, warm=1)
    t_samp = timeit(lambda: m.sample(N), iters=2, warm=1)
    return {"config": f"N1: LJ.yaml flow, 2 x NSF_AR(96, K=32, H=354, {precision} conditioners), batch {N}",
            "evaluate_ms": t_eval, "sample_ms": t_samp, "evaluate_samples_per_s": N / t_eval * 1e3,
            "sample_samples_per_s": N / t_samp * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    n3 = 131072 if a.quick else 1 << 19
    n4 = 1 << 18 if a.quick else 1 << 20
    for fn in (lambda: cfg1(100), lambda: cfg1(800), lambda: cfg3(n3), lambda: cfg4(n4),
               lambda: cfg5(65536, 128, "fp32"), lambda: cfg5(65536, 128, "bf16"),
               lambda: cfg_nsf_ar(65536, "fp32"), lambda: cfg_nsf_ar(65536, "bf16"),
               lambda: cfg_train(262144, 128, "fp32"), lambda: cfg_train(262144, 128, "bf16"),
               lambda: cfg_train(262144, 800, "bf16"), lambda: cfg_targets(1 << 20)):
        l0 = _lib.launch_count()
        r = fn()
        r["libnfk_launches"] = _lib.launch_count() - l0
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
