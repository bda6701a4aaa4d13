#!/usr/bin/env python
"""Step 3 of tools/reference_dump/README.md: turns the output of dump_reference.cpp (real pclomp / fast_gicp / PCL, run on a ROS
box on the inputs of tools/reference_dump/export_inputs.py) into tests/golden/reference_dump_v1.npz, the fixture
tests/test_reference_dump.py compares the oracle with.   python tests/golden/import_reference_dump.py /tmp/b2r_dump"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main(d, path=None):
    cases = [l.split() for l in open(os.path.join(d, "cases.txt")) if l.strip()]
    dump = [l.split() for l in open(os.path.join(d, "reference_dump.txt")) if l.strip() and not l.startswith("#")]
    versions = open(os.path.join(d, "reference_dump.txt")).readline().strip()
    out = {"versions": np.array(versions)}
    clouds = {}
    align_rows, align_meta = [], []
    for row in dump:
        if row[0] == "align" and row[2] != "skipped":
            cid = int(row[1])
            case = next(c for c in cases if c[0] == "align" and int(c[1]) == cid)
            for f in (case[3], case[4]):
                clouds.setdefault(f, np.fromfile(os.path.join(d, f), dtype=np.float32).reshape(-1, 4))
            align_meta.append((cid, case[2], case[3], case[4], float(case[5]), case[6]))
            align_rows.append([float(row[3]), float(row[4])] + [float(x) for x in row[5:21]] + [float(x) for x in case[7:23]])
        elif row[0] == "filter":
            k = int(row[1])
            case = next(c for c in cases if c[0] == "filter" and int(c[1]) == k)
            clouds.setdefault(case[2], np.fromfile(os.path.join(d, case[2]), dtype=np.float32).reshape(-1, 4))
            for name in ("dist", "vg", "radius", "sor"):
                out[f"filter{k}_{name}"] = np.fromfile(os.path.join(d, f"{name}_{k}.bin"), dtype=np.float32).reshape(-1, 4)
            out[f"filter{k}_input"] = np.array(case[2])
            out[f"filter{k}_params"] = np.array([float(x) for x in case[3:10]])
    out["align_rows"] = np.array(align_rows, dtype=np.float64)  # converged, fitness, T[16] column-major, guess[16] column-major
    out["align_meta"] = np.array(align_meta, dtype=object)
    for f, c in clouds.items():
        out["cloud:" + f] = c
    path = path or os.path.join(HERE, "reference_dump_v1.npz")
    np.savez_compressed(path, **out)
    print(f"{len(align_rows)} alignments, {len(clouds)} clouds -> {path}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/tmp/b2r_dump")
