#!/usr/bin/env python
"""BASELINE config 3: Lennard-Jones 38-particle cluster (d = 114), 8 x NSF_CL(38, dim 3, K 8, B 4, H 800,
masks cycling [0],[1],[2],[0,1],[1,2],[0,2]) — sampling and log-prob evaluation of a GLOBAL batch of 2^20 rows
sharded by rows over the GPUs of one box (strong scaling: 131,072 rows per GPU on 8), no data-path collective;
one scalar all-reduce for the reported mean log-prob.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
        tools/bench_cfg3_dist.py
"""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200 import _lib
from normalizingflow_b200.dist import broadcast_parameters, global_mean, shard_rows, sharded_sample
from normalizingflow_b200.flows import NSF_CL
from normalizingflow_b200.models import GaussianPrior, NormalizingFlowModel


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    masks = ([[0], [1], [2], [0, 1], [1, 2], [0, 2]] * 2)[:8]
    fl = [NSF_CL(38, dim=3, K=8, B=4.0, hidden_dim=800, mask=mk) for mk in masks]
    for f in fl:
        f.psi.precision = "bf16"
    m = NormalizingFlowModel(GaussianPrior(114, device=dev), fl, device=dev).to(dev)
    broadcast_parameters(m)
    N = 1 << 20
    a, b = shard_rows(N, rank, world)
    x = torch.randn(b - a, 114, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank))
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(2):
        lp = m.evaluate(x)
        sharded_sample(m, N, rank, world, seed=1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.launch_count()
    iters = 3
    e[0].record()
    for _ in range(iters):
        lp = m.evaluate(x)
    e[1].record()
    for _ in range(iters):
        xs, lpx, z, _ = sharded_sample(m, N, rank, world, seed=1)
    e[2].record()
    torch.cuda.synchronize()
    t = torch.tensor([e[0].elapsed_time(e[1]) / iters, e[1].elapsed_time(e[2]) / iters], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    mean_lp = float(global_mean(lp))
    mean_lpx = float(global_mean(lpx))
    if rank == 0:
        print(json.dumps({"config": "3: LJ-38 d=114, 8 x NSF_CL(38, dim 3, K 8, B 4, H 800, bf16 conditioner), global batch 1048576 sharded by rows",
                          "n_gpus": world, "rows_per_gpu": b - a, "evaluate_ms": float(t[0]), "sample_ms": float(t[1]),
                          "evaluate_samples_per_s": N / float(t[0]) * 1e3, "sample_samples_per_s": N / float(t[1]) * 1e3,
                          "mean_log_prob_of_inputs": mean_lp, "mean_log_prob_of_samples": mean_lpx,
                          "libnfk_launches_per_pass": (_lib.launch_count() - l0) / (2 * iters)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
