#!/usr/bin/env python
"""Turn ncu outputs into the small summaries committed under profiles/.

    python tools/summarize_ncu.py launches <launches.csv> <out.csv>     # gpu__time_duration.sum launch list
    python tools/summarize_ncu.py report <prof.ncu-rep> <out.json>      # --set full capture (needs `ncu` on PATH)
"""
import collections
import csv
import io
import json
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.sum", "lts__t_bytes.sum", "smsp__cycles_active.avg",
           # pipe utilisation and the stall mix (fused layer kernel analysis, profiles/fused_analysis_r01.md)
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.min.pct_of_peak_sustained_active", "smsp__issue_active.max.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches(src, dst):
    """launch list (one row per launch and metric) -> per-kernel launches, average duration, share of the
    captured time and, when the capture carries them, DRAM bytes per launch"""
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", "")) * scale.get(r[mu], 1.0)
        a = agg.setdefault(r[kn], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        if r[mn] == "gpu__time_duration.sum":
            a["n"] += 1
            a["us"] += v
        elif r[mn] == "dram__bytes_read.sum":
            a["rd"] += v
        elif r[mn] == "dram__bytes_write.sum":
            a["wr"] += v
    tot = sum(a["us"] for a in agg.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "avg_us", "total_us", "share_of_captured_time", "dram_read_bytes_per_launch",
                    "dram_write_bytes_per_launch"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            n = max(1, a["n"])
            w.writerow([k[:120], a["n"], f"{a['us'] / n:.1f}", f"{a['us']:.1f}", f"{a['us'] / tot:.4f}",
                        f"{a['rd'] / n:.0f}", f"{a['wr'] / n:.0f}"])
    print(f"{dst}: {len(agg)} kernels, {sum(a['n'] for a in agg.values())} launches, {tot / 1e3:.2f} ms captured")


def report(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:120]}
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                d[m] = {"value": r[i], "unit": units[i]}
        out.append(d)
    json.dump(out, open(dst, "w"), indent=1)
    print(f"{dst}: {len(out)} launches")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
