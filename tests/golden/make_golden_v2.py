"""Generates tests/golden/golden_v2.npz: the rows added after v1 — SMALL_GICP, the map cloud
(MapCloudGenerator + ApproximateMeanVoxelGrid) and the odometry state machine — frozen from the oracle on seeded
synthetic scans (same caveat as make_golden.py: the reference ships no vectors; these freeze the checker).

    python tests/golden/make_golden_v2.py     # rewrites golden_v2.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mrg_slam_b200 import synth  # noqa: E402
from mrg_slam_b200.odometry import ScanMatchingOdometry  # noqa: E402
from tests import oraclelib as O  # noqa: E402
from tests.golden.make_golden import guesses, sha  # noqa: E402

MAP_FIRST, MAP_STEP, MAP_COUNT = 40, 3, 4
ODO_FIRST, ODO_COUNT = 20, 7


def prefiltered(sensor, idx):
    c = O.distance_filter(synth.scan(sensor, idx), 0.1, 35.0)
    c, _ = O.voxelgrid(c, 0.1, 1)
    return c[O.radius_outlier(c, 0.5, 2)]


def map_inputs():
    clouds = [prefiltered(synth.VLP16, MAP_FIRST + MAP_STEP * i) for i in range(MAP_COUNT)]
    p0 = np.linalg.inv(synth.pose(MAP_FIRST))
    poses = [p0 @ synth.pose(MAP_FIRST + MAP_STEP * i) for i in range(MAP_COUNT)]
    return clouds, poses


def sorted_map(points, keys):
    order = np.lexsort((keys[:, 0], keys[:, 1], keys[:, 2]))
    return points[order]


def build():
    out = {"generator_version": np.array(2)}
    A, B = prefiltered(synth.VLP16, 3), prefiltered(synth.VLP16, 4)
    gt = np.linalg.inv(synth.pose(3)) @ synth.pose(4)
    r = O.Registration(O.default_params(O.SMALL_GICP))
    r.setInputTarget(A); r.setInputSource(B)
    Ts, conv, its, fits = [], [], [], []
    for g in guesses(gt):
        res = r.align(g)
        Ts.append(np.array(list(res.T), dtype=np.float32)); conv.append(res.converged); its.append(res.iterations); fits.append(r.getFitnessScore())
    out.update(sgicp_T=np.stack(Ts), sgicp_conv=np.array(conv), sgicp_iters=np.array(its), sgicp_fitness=np.array(fits))
    far = gt.copy(); far[:3, 3] += [0.8, -0.6, 0.1]
    err, H, b, corr, _ = r.linearize(far)
    out.update(sgicp_lin_err=np.array(err), sgicp_lin_H=H, sgicp_lin_b=b, sgicp_corr_sha=np.array(sha(corr)))
    clouds, poses = map_inputs()
    for tag, (res_, mp, far_) in {"fine": (0.05, 1, -1.0), "coarse": (0.5, 2, 25.0)}.items():
        pts, keys = O.map_cloud(clouds, poses, [1, 0, 0, 0], res_, mp, far_, True)
        out.update({f"map_{tag}_n": np.array(len(pts)), f"map_{tag}_sha": np.array(sha(sorted_map(pts, keys)))})
    odo = ScanMatchingOdometry(O.Registration(O.default_params(O.FAST_VGICP)))
    traj = [odo.matching(0.1 * i, prefiltered(synth.VLP16, ODO_FIRST + i)) for i in range(ODO_COUNT)]
    out.update(odo_traj=np.stack(traj).astype(np.float32), odo_switches=np.array(odo.keyframe_switches))
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v2.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path), "bytes")
