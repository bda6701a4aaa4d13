"""Small workload for compute-sanitizer (run under gpurun): every method once on small clouds, batch + single + fitness + filters.

  compute-sanitizer --tool memcheck  python tests/sanitize_driver.py
  compute-sanitizer --tool racecheck python tests/sanitize_driver.py
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from mrg_slam_b200 import lib as B  # noqa: E402
from mrg_slam_b200 import synth  # noqa: E402


def main():
    reg0 = B.Registration(B.default_config(B.FAST_VGICP))
    raw = [synth.scan(synth.VLP16, 10 + i) for i in range(3)]
    clouds_np = []
    for r in raw:
        c = reg0.distance_filter(r, 0.5, 30.0)
        c, _ = reg0.voxelgrid(c, 0.4, 1)
        c = reg0.radius_outlier(c, 0.9, 2)
        clouds_np.append(c)
    reg0.statistical_outlier(clouds_np[0], 20, 1.0)
    print("points", [len(c) for c in clouds_np])
    far = np.eye(4)
    far[0, 3] = 1.5  # far guess: many source points without a voxel / correspondence (exercises the hit rings)
    for method in (B.FAST_VGICP, B.NDT_OMP, B.FAST_GICP, B.SMALL_GICP):
        reg = B.Registration(B.default_config(method))
        cl = [B.Cloud(reg, c) for c in clouds_np]
        res = reg.align_batch([cl[1], cl[2], cl[2]], [cl[0], cl[0], cl[1]], [np.eye(4), far, np.eye(4)], with_fitness=True)
        print(method, [(r.converged, r.iterations, round(r.fitness, 4)) for r in res])
        for c in cl:
            c.close()
        reg.close()
    reg0.close()


if __name__ == "__main__":
    main()
