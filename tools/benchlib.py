"""Shared measurement helpers of bench.py and tools/bench_configs.py: measured peaks, the nvidia-smi clock
sampler (B200_PROFILING.md: clocks DURING the timed region), CUDA-event timing with max-over-ranks, the
algorithmic byte / flop figures of SURVEY.md 8(d) and the N-rank host-copy ceiling probe."""
import json
import os
import subprocess
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------
# peaks
# ------------------------------------------------------------------------------------------------
def peaks():
    """(hbm GB/s, bf16 sustained TFLOP/s, bf16 burst TFLOP/s, source string)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return (float(j["hbm_gbs"]), float(j.get("bf16_tflops_sustained") or 1500.0),
                float(j.get("bf16_tflops") or 1600.0), "measured (MEASURED_PEAKS.json)")
    return 6650.0, 1500.0, 1600.0, "fallback (B200_PROFILING.md)"


# SURVEY.md 8(d) ---------------------------------------------------------------------------------
def rqs_layer_bytes(size, dim, n_mask, K=8):
    """algorithmic bytes per row per layer pass of an RQS coupling: params + x + z + log-det rmw"""
    f_t = size * (dim - n_mask)
    return f_t * (3 * K - 1) * 4 + 2 * size * dim * 4 + 8


def conditioner_flops(size, dim, n_mask, H, K=8):
    """flop per row per layer pass of the conditioner MLP: 2 (F_c H + H^2 + H (3K-1) F_t)"""
    f_c, f_t = size * n_mask, size * (dim - n_mask)
    return 2.0 * (f_c * H + H * H + H * (3 * K - 1) * f_t)


def hbm_roofline(kernel, alg_bytes, ms, extra=None):
    hbm, _, _, src = peaks()
    ach = alg_bytes / (ms * 1e-3) / 1e9 if ms else None
    r = {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": hbm, "unit": "GB/s",
         "frac": (ach / hbm) if ach else None, "traffic": None, "peak_source": src + " hbm_gbs",
         "algorithmic_bytes": alg_bytes, "ms": ms}
    if extra:
        r.update(extra)
    return r


def tensor_roofline(kernel, flops, ms, extra=None):
    _, sus, burst, src = peaks()
    ach = flops / (ms * 1e-3) / 1e12 if ms else None
    r = {"bound": "tensor", "kernel": kernel, "achieved": ach, "peak": sus, "unit": "TFLOP/s",
         "frac": (ach / sus) if ach else None, "traffic": None,
         "peak_source": src + " bf16_tflops_sustained (kernel timed inside a long step)", "algorithmic_flops": flops, "ms": ms}
    if extra:
        r.update(extra)
    return r


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [t.strip() for t in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# timing
# ------------------------------------------------------------------------------------------------
def world_info():
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")))


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(values, dev, world):
    t = torch.tensor(list(values), dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def time_steps(fn, steps, warmup, dev, world):
    """ms per call of fn(): `warmup` untimed calls, then `steps` calls between CUDA events, bracketed by a
    barrier + synchronize on both sides, MAX over ranks."""
    for _ in range(warmup):
        fn()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier(world)
    return max_over_ranks([e0.elapsed_time(e1) / steps], dev, world)[0]


# ------------------------------------------------------------------------------------------------
# host-copy ceiling: what the box's host memory system gives N ranks copying concurrently
# ------------------------------------------------------------------------------------------------
def host_copy_ceiling(dev, world, h2d_bytes, d2h_bytes, chunk_bytes, iters=3):
    """Every rank concurrently streams `h2d_bytes` host->device and `d2h_bytes` device->host per iteration
    from / into its own pinned buffers in `chunk_bytes` pieces on two streams (the traffic pattern of the
    e2e step without any kernel).  Returns per-rank and aggregate GB/s (max-over-ranks time) -- the ceiling
    an end-to-end step with these byte counts cannot beat on this box."""
    n_in, n_out = max(1, h2d_bytes // 4), max(1, d2h_bytes // 4)
    ck = max(1, chunk_bytes // 4)
    h_in = torch.empty(n_in, dtype=torch.float32).pin_memory()
    h_in.zero_()
    h_out = torch.empty(n_out, dtype=torch.float32).pin_memory()
    d_in = torch.empty(min(ck, n_in) * 2, dtype=torch.float32, device=dev)
    d_out = torch.zeros(min(ck, n_out) * 2, dtype=torch.float32, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def once():
        k = 0
        with torch.cuda.stream(s_in):
            for a in range(0, n_in, ck):
                b = min(n_in, a + ck)
                o = (k & 1) * min(ck, n_in)
                d_in[o:o + b - a].copy_(h_in[a:b], non_blocking=True)
                k += 1
        k = 0
        with torch.cuda.stream(s_out):
            for a in range(0, n_out, ck):
                b = min(n_out, a + ck)
                o = (k & 1) * min(ck, n_out)
                h_out[a:b].copy_(d_out[o:o + b - a], non_blocking=True)
                k += 1

    once()
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(iters):
        once()
    s_in.synchronize()
    s_out.synchronize()
    dt = (time.perf_counter() - t0) / iters
    dt = max_over_ranks([dt], dev, world)[0]
    barrier(world)
    per_rank = (h2d_bytes + d2h_bytes) / dt / 1e9
    return {"seconds_per_step_bytes": dt, "gbps_per_rank": per_rank, "gbps_aggregate": per_rank * world,
            "h2d_bytes": h2d_bytes, "d2h_bytes": d2h_bytes, "chunk_bytes": chunk_bytes, "ranks": world,
            "how": "all ranks concurrently, pinned buffers, H2D and D2H on two streams, no kernels; max over ranks"}
