"""Small driver for ncu captures (run under gpurun): one prefilter + one batch of FAST_VGICP / NDT / GICP aligns.

  ncu --set full --clock-control none --import-source on -k regex:knn_cov -c 1 -o gpurun_out/prof python tests/prof_driver.py
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from mrg_slam_b200 import lib as B  # noqa: E402
from mrg_slam_b200 import synth  # noqa: E402


def main():
    method = sys.argv[1] if len(sys.argv) > 1 else "FAST_VGICP"
    pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    reg = B.Registration(B.default_config(getattr(B, method)))
    scans = [reg.prefilter(synth.scan(synth.HDL64, 100 + i)) for i in range(pairs + 1)]
    for _ in range(reps):
        clouds = [B.Cloud(reg, s) for s in scans]
        res = reg.align_batch(clouds[1:], clouds[:-1], [np.eye(4)] * pairs, with_fitness=True)
        for c in clouds:
            c.close()
    print(method, "pairs", pairs, "converged", sum(r.converged for r in res), "timings", reg.last_timings(), "launches", reg.kernel_launches())


if __name__ == "__main__":
    main()
