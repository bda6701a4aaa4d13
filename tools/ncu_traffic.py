#!/usr/bin/env python
"""Writes profiles/ncu_traffic.json: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels
captured with `ncu --set full` under tools/ncu_bench.sh, keyed by the bench workload ("METHOD:pairs") and kernel family.
bench.py reads it to fill `roofline.traffic` for the dominant kernel.

  tools/ncu_traffic.py FAST_VGICP:32 knn_cov=gpurun_out/ncu_X_knn_cov.ncu-rep lsq_eval=gpurun_out/ncu_X_lsq_eval.ncu-rep
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram_bytes(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        v, u = d[k]
        tot += float(v.replace(",", "")) * UNIT[u]
    dur_v, dur_u = d["gpu__time_duration.sum"]
    dur_us = float(dur_v.replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(dur_u, 1.0)
    return tot, dur_us, d.get("Kernel Name", ("?",))[0]


def main():
    key = sys.argv[1]
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        table = json.load(open(path))
    except Exception:
        table = {}
    entry = table.setdefault(key, {})
    for arg in sys.argv[2:]:
        fam, rep = arg.split("=")
        b, us, name = dram_bytes(rep)
        entry[fam] = {"dram_bytes_per_launch": b, "kernel": name.split("(")[0], "duration_us_under_ncu": us,
                      "source": f"ncu --set full --clock-control none, one launch of the bench step ({os.path.basename(rep)})"}
    json.dump(table, open(path, "w"), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
