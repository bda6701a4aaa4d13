import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """The CPU-side shared libraries (oracle, synth) are built on demand; libb2r.so must already exist or build."""
    for d in ("oracle", "mrg_slam_b200/synth"):
        subprocess.run(["make", "-C", os.path.join(ROOT, d)], check=True, capture_output=True)
    if not os.path.exists(os.path.join(ROOT, "mrg_slam_b200", "libb2r.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "mrg_slam_b200/csrc"), "-j8"], check=True, capture_output=True)


def oracle_prefilter(cloud):
    """distance 0.1-35 m, VoxelGrid 0.1, RADIUS (0.5, 2): the YAML prefilter chain, on the oracle."""
    from tests import oraclelib as O

    c = O.distance_filter(cloud, 0.1, 35.0)
    c, _ = O.voxelgrid(c, 0.1, 1)
    return c[O.radius_outlier(c, 0.5, 2)]


@pytest.fixture(scope="session")
def vlp16_pair():
    """Two consecutive prefiltered VLP-16 scans (~15k points each) + ground-truth relative pose (source -> target)."""
    from mrg_slam_b200 import synth

    a = oracle_prefilter(synth.scan(synth.VLP16, 3))
    b = oracle_prefilter(synth.scan(synth.VLP16, 4))
    gt = np.linalg.inv(synth.pose(3)) @ synth.pose(4)
    return a, b, gt


@pytest.fixture(scope="session")
def small_pair():
    """A small pair (VLP-16, voxelised at 0.4 m: ~3k points) for the slower oracle-side checks."""
    from mrg_slam_b200 import synth
    from tests import oraclelib as O

    def prep(c):
        c = O.distance_filter(c, 0.5, 30.0)
        c, _ = O.voxelgrid(c, 0.4, 1)
        return c

    a, b = prep(synth.scan(synth.VLP16, 10)), prep(synth.scan(synth.VLP16, 11))
    gt = np.linalg.inv(synth.pose(10)) @ synth.pose(11)
    return a, b, gt


def pose_error(Ta, Tb):
    d = np.linalg.inv(Ta) @ Tb
    return float(np.linalg.norm(d[:3, 3])), float(np.arccos(np.clip((np.trace(d[:3, :3]) - 1) / 2, -1, 1)))
