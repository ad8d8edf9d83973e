#!/usr/bin/env python
"""Run bench.py several times with different extra flags in ONE box lease and print the key numbers
side by side (A/B comparisons must share a box: PCIe and host load vary between leases).

    python tools/bench_sweep.py "--e2e-wait 1" "--e2e-wait 0" ...
"""
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for flags in sys.argv[1:]:
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--no-cpu-baseline"] + flags.split(),
                         capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print(f"{flags:40s} value {d['value'] / 1e6:8.2f} M/s  {d['ms_per_step']:7.3f} ms  | e2e {e.get('value', 0) / 1e6:8.2f} M/s "
              f"{e.get('ms_per_step', 0):7.3f} ms | frac {d['roofline']['frac']:.3f}", flush=True)
    except Exception as ex:                                  # noqa: BLE001
        print(flags, "FAILED", ex, out.stderr[-400:], flush=True)
