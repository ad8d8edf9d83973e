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
           "sm__inst_executed_pipe_xu.sum", "lts__t_bytes.sum", "smsp__cycles_active.avg"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] == "ns" else v * (1e3 if r[mu] == "ms" else 1.0)      # -> us
        a = agg.setdefault(r[kn], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "avg_us", "total_us", "share_of_captured_time"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k[:120], a[0], f"{a[1] / a[0]:.1f}", f"{a[1]:.1f}", f"{a[1] / tot:.4f}"])
    print(f"{dst}: {len(agg)} kernels, {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms captured")


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
