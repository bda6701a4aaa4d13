#!/usr/bin/env python
"""Step 1 of tools/reference_dump/README.md: the seeded inputs and the case list for dump_reference.cpp."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mrg_slam_b200 import synth  # noqa: E402
from tests import oraclelib as O  # noqa: E402


def main(out):
    os.makedirs(out, exist_ok=True)
    clouds, lines = {}, []

    def cloud(name, pts):
        if name not in clouds:
            clouds[name] = pts
            np.ascontiguousarray(pts, dtype=np.float32).tofile(os.path.join(out, f"cloud_{name}.bin"))
        return f"cloud_{name}.bin"

    def pre(sensor, idx):
        c = O.distance_filter(synth.scan(sensor, idx), 0.1, 35.0)
        c, _ = O.voxelgrid(c, 0.1, 1)
        return c[O.radius_outlier(c, 0.5, 2)]

    # alignments: the VLP-16 pair of the test-suite (three guesses) and the first 16 pairs of the loop-closure batch
    a, b = pre(synth.VLP16, 3), pre(synth.VLP16, 4)
    gt = np.linalg.inv(synth.pose(3)) @ synth.pose(4)
    near = gt.copy(); near[0, 3] -= 0.25; near[1, 3] += 0.1
    cid = 0
    for method, res, nn in (("FAST_VGICP", 1.0, "-"), ("FAST_GICP", 1.0, "-"), ("NDT_OMP", 1.0, "DIRECT7"), ("NDT_OMP", 0.5, "DIRECT1"),
                            ("NDT_OMP", 1.0, "KDTREE"), ("SMALL_GICP", 1.0, "-"), ("GICP", 1.0, "-")):
        for g in (np.eye(4), near):
            lines.append(f"align {cid} {method} {cloud('vlp16_3', a)} {cloud('vlp16_4', b)} {res} {nn} " +
                         " ".join(repr(float(x)) for x in np.asarray(g, dtype=np.float32).T.reshape(16)))
            cid += 1
    n_targets, n_cand = 256, 16
    poses = [synth.pose(bench.FIRST_SCAN + i) for i in range(n_targets + n_cand)]
    pairs, guesses = bench.batch_pairs(n_targets, n_cand, poses)
    pool = bench.oracle_pool(O, sorted({c for i in range(16) for c in pairs[i]}))
    for i in range(16):
        ti, ci = pairs[i]
        for method in ("FAST_VGICP", "FAST_GICP", "NDT_OMP"):
            lines.append(f"align {cid} {method} {cloud(f'kf_{ti}', pool[ti])} {cloud(f'kf_{ci}', pool[ci])} 1.0 DIRECT7 " +
                         " ".join(repr(float(x)) for x in np.asarray(guesses[i], dtype=np.float32).T.reshape(16)))
            cid += 1
    # filters: raw scans through distance filter / VoxelGrid 0.1 / RADIUS (0.5, 2) / STATISTICAL (30, 1.2)
    for k, (sensor, idx) in enumerate(((synth.VLP16, 9), (synth.HDL64, 5))):
        lines.append(f"filter {k} {cloud(f'raw_{k}', synth.scan(sensor, idx))} 0.1 35.0 0.1 0.5 2 30 1.2")
    open(os.path.join(out, "cases.txt"), "w").write("\n".join(lines) + "\n")
    print(f"{len(clouds)} clouds, {len(lines)} cases -> {out}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/tmp/b2r_dump")
