#!/usr/bin/env python
"""Data-parallel training step on N GPUs of one box (BASELINE configs 2 and 4): every rank
owns its rows of the global batch, runs forward-KL loss + backward on the libnfk kernels, the gradients
are summed with ONE flat-bucket NCCL all-reduce over NVLink (normalizingflow_b200/dist.py) between
backward() and the optimizer step (applications/src/train.py:27-28), Adam steps.  Weak scaling: per-GPU
rows fixed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29512 tools/bench_train_dist.py [--steps 5]

Checks, printed with the numbers: (a) after the all-reduce every rank holds bit-identical gradients and,
after the step, bit-identical parameters; (b) the averaged gradient of the N-rank job equals the gradient
of ONE process evaluated on the concatenated global batch (rank 0 recomputes it when the batch fits)."""
import argparse, json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200 import _lib
from normalizingflow_b200.dist import GradBucket, allreduce_gradients, broadcast_parameters
from normalizingflow_b200.flows import NSF_CL, Planar, Radial
from normalizingflow_b200.models import GaussianPrior, NormalizingFlowModel


def build(kind, dev):
    torch.manual_seed(0)
    if kind == "planar":
        d, fl = 128, [Planar(128) for _ in range(32)]
    elif kind == "radial":
        d, fl = 128, [Radial(128, per_sample=True) for _ in range(32)]
    else:
        H = int(kind.split("_h")[1])
        d, fl = 64, [NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=H, mask=[i % 2]) for i in range(8)]
        for f in fl:
            f.psi.precision = "bf16"
    return d, NormalizingFlowModel(GaussianPrior(d, device=dev), fl, device=dev).to(dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--kinds", nargs="*", default=["planar", "radial", "nsf_h128", "nsf_h800"])
    ap.add_argument("--check-rows", type=int, default=16384, help="rows per rank of the gradient-equality check")
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    for kind in a.kinds:
        d, m = build(kind, dev)
        broadcast_parameters(m)
        params = list(m.parameters())
        n_grad = sum(p.numel() for p in params)
        rows = (1 << 20) if kind in ("planar", "radial") else 262144
        x = torch.randn(rows, d, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank))
        opt = torch.optim.Adam(params, lr=1e-4)
        bucket = GradBucket(params)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

        def step(timed=False):
            z, plp, ld = m(x)
            loss = -torch.mean(plp + ld)
            bucket.zero()
            loss.backward()
            if timed:
                ev[1].record()
            nbytes = bucket.allreduce()
            if timed:
                ev[2].record()
            opt.step()
            return nbytes

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = _lib.launch_count()
        t_ar = 0.0
        ev[0].record()
        for _ in range(a.steps):
            nbytes = step(True)
            ev[3].record()
            torch.cuda.synchronize()
            t_ar += ev[1].elapsed_time(ev[2])
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[3]) / a.steps
        t = torch.tensor([ms, t_ar / a.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # (a) replicas stay bit-identical
        flat = torch.cat([p.detach().reshape(-1) for p in params])
        ident = True
        if world > 1:
            lo, hi = flat.clone(), flat.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            ident = bool(torch.equal(lo, hi))
        # (b) averaged sharded gradient == gradient of one process on the concatenated batch
        rel = None
        if world > 1:
            xs = torch.randn(a.check_rows, d, device=dev, generator=torch.Generator(device=dev).manual_seed(900 + rank))
            z, plp, ld = m(xs)
            opt.zero_grad(set_to_none=True)
            (-torch.mean(plp + ld)).backward()
            allreduce_gradients(params)
            g_dist = torch.cat([p.grad.reshape(-1) for p in params]).clone()
            gathered = [torch.empty_like(xs) for _ in range(world)]
            dist.all_gather(gathered, xs)
            if rank == 0:
                xa = torch.cat(gathered)
                z, plp, ld = m(xa)
                opt.zero_grad(set_to_none=True)
                (-torch.mean(plp + ld)).backward()
                g_one = torch.cat([p.grad.reshape(-1) for p in params])
                rel = float((g_dist - g_one).abs().max() / g_one.abs().max().clamp_min(1e-30))
        if rank == 0:
            print(json.dumps({"config": f"train {kind}: forward-KL step, {rows} rows/GPU, Adam, NCCL grad all-reduce",
                              "n_gpus": world, "ms_per_step": float(t[0]), "allreduce_ms": float(t[1]),
                              "allreduce_bytes": nbytes, "grad_elements": n_grad,
                              "train_samples_per_s": world * rows / float(t[0]) * 1e3,
                              "replicas_bit_identical_after_steps": ident,
                              "sharded_vs_single_process_gradient_rel_err": rel,
                              "libnfk_launches_per_step": (_lib.launch_count() - l0) / a.steps}), flush=True)
        del m, opt, x
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
