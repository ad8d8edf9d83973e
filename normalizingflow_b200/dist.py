"""Multi-GPU plumbing: one process per GPU (torchrun), batch sharded by rows.

Sampling, density evaluation and HMC need no data-path collective — rows are independent
(nf/models.py:16-18), weights are replicated.  Training adds ONE flat-bucket all-reduce of the
gradients per step (NCCL over NVLink on a GPU box, gloo in the CPU tests), to be called between
``loss.backward()`` and ``optimizer.step()`` (the reference's applications/src/train.py:27-28).
Bug-compatible ``Radial`` all-reduces one float per layer inside its op (quirk Q9).
"""
from __future__ import annotations

import os
from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the default group when
    WORLD_SIZE > 1 (nccl when CUDA is present, else gloo)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def bind_to_gpu_numa_node(local_rank: int) -> str:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that pinned host
    buffers allocated afterwards (first touch) and the copy-issuing thread are local to the GPU's
    PCIe root: with 8 ranks streaming host batches, cross-socket traffic is what caps the
    end-to-end rate.  Best effort; always returns a one-line description: what was bound, or
    ``"unbound: <why>"`` when nothing was changed (single-node VM, topology unreadable, ...)."""
    try:
        import pynvml
    except Exception as e:                                      # noqa: BLE001
        return f"unbound: pynvml not importable ({type(e).__name__})"
    try:
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(_physical_index(local_rank))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]                                  # 0000:1b:00.0
    except Exception as e:                                      # noqa: BLE001
        return f"unbound: NVML query failed ({type(e).__name__}: {e})"
    try:
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
    except Exception as e:                                      # noqa: BLE001
        return f"unbound: /sys/bus/pci/devices/{bdf}/numa_node unreadable ({type(e).__name__})"
    try:
        n_nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
    except OSError:
        n_nodes = 0
    if node < 0:
        return f"unbound: GPU {local_rank} ({bdf}) reports numa_node={node} (no affinity exposed; host has {n_nodes} NUMA node(s))"
    try:
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = os.sched_getaffinity(0)
        cpus = sorted(set(cpus) & allowed)
        if not cpus:
            return f"unbound: no allowed CPU on NUMA node {node} of GPU {local_rank} ({bdf})"
        if n_nodes <= 1:
            return f"unbound: single NUMA node host (GPU {local_rank} {bdf} -> node {node}, {len(cpus)} CPUs): nothing to bind"
        os.sched_setaffinity(0, cpus)
        return f"GPU {local_rank} ({bdf}) -> NUMA node {node}, {len(cpus)} CPUs"
    except Exception as e:                                      # noqa: BLE001 - best effort by design
        return f"unbound: {type(e).__name__}: {e}"


def _physical_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        ids = [v for v in vis.split(",") if v.strip() != ""]
        if local_rank < len(ids) and ids[local_rank].strip().isdigit():
            return int(ids[local_rank])
    return local_rank


def _parse_cpulist(text: str):
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def shard_rows(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range [start, stop) of ``rank``; sizes differ by at most one row."""
    base, rem = divmod(n_total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
    from . import _lib                      # .data writes do not bump Parameter._version
    _lib.invalidate_caches()


def allreduce_gradients(params: Iterable[torch.nn.Parameter], average: bool = True, group=None) -> int:
    """Sum (or average) the gradients of ``params`` over the ranks with ONE all-reduce on a flat
    fp32 bucket.  Returns the bucket size in bytes (0 when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    plist = [p for p in params if p.grad is not None]
    if not plist:
        return 0
    flat = torch.cat([p.grad.reshape(-1).float() for p in plist])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= world
    off = 0
    for p in plist:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat.numel() * 4


class GradBucket:
    """Persistent flat fp32 gradient bucket: every parameter's ``.grad`` is a view into ONE buffer, so the
    per-step gradient exchange is a single in-place all-reduce with no gather / scatter copies around it
    (``allreduce_gradients`` builds the bucket with ``torch.cat`` and copies the result back: at 10 M
    gradient elements that is ~1.3 ms per step on a B200 around a 0.2 ms NVLink all-reduce).

        bucket = GradBucket(model.parameters())
        for batch in data:
            bucket.zero()                  # instead of optimizer.zero_grad()
            loss(batch).backward()         # autograd accumulates into the views
            bucket.allreduce()             # one NCCL all-reduce (sum, then 1/world)
            optimizer.step()
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise ValueError("GradBucket holds fp32 gradients (the reference trains in fp32)")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero(self) -> None:
        self.flat.zero_()
        for p, off in zip(self.params, self._offsets()):
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                p.grad = self.flat[off:off + p.numel()].view_as(p)     # someone called zero_grad(set_to_none=True)

    def _offsets(self):
        off = 0
        for p in self.params:
            yield off
            off += p.numel()

    def allreduce(self, average: bool = True) -> int:
        if not (dist.is_available() and dist.is_initialized()):
            return 0
        world = dist.get_world_size(self.group)
        if world == 1:
            return 0
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        if average:
            self.flat.mul_(1.0 / world)
        return self.nbytes


def global_mean(t: torch.Tensor, group=None) -> torch.Tensor:
    """Mean of a per-row quantity over all ranks (e.g. the reported mean log-prob)."""
    s = torch.stack([t.float().sum(), torch.tensor(float(t.numel()), device=t.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, group=group)
    return s[0] / s[1]


def sharded_sample(model, n_total: int, rank: int, world: int, seed: int = 0):
    """Each rank draws and pushes forward its own rows of a global sample of ``n_total`` rows.
    The latent rows are generated from a per-rank generator so the union over ranks does not
    depend on the order in which ranks run."""
    start, stop = shard_rows(n_total, rank, world)
    prior = model.prior
    if hasattr(prior, "generator"):
        g = torch.Generator(device=prior.device)
        g.manual_seed(seed * 1_000_003 + rank)
        prior.generator = g
    x, log_px, z = model.sample(stop - start)
    return x, log_px, z, (start, stop)
