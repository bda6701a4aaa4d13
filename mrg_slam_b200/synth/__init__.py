"""Deterministic synthetic LiDAR scans (VLP-16 / HDL-64 KITTI-shape / OS1-128), see synth.cpp.

Layout follows the KITTI velodyne records the reference's replay drivers stream
(/root/reference python_scripts/kitti_singlerobot_processor.py:164-185): float32 x,y,z,intensity.
"""
import ctypes
import os

import numpy as np

VLP16, HDL64, OS1_128, OS1_128_1M = 0, 1, 2, 3
DEFAULT_SEED = 0x5EED0000

_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libb2r_synth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = ctypes.CDLL(path)
        lib.b2r_synth_num_rays.restype = ctypes.c_int
        lib.b2r_synth_num_rays.argtypes = [ctypes.c_int]
        lib.b2r_synth_scan.restype = ctypes.c_int
        lib.b2r_synth_scan.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p]
        lib.b2r_synth_pose.restype = None
        lib.b2r_synth_pose.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p]
        _lib = lib
    return _lib


def scan(sensor: int, scan_idx: int, seed: int = DEFAULT_SEED) -> np.ndarray:
    """Returns an (n, 4) float32 array x,y,z,intensity in the sensor frame."""
    lib = _load()
    cap = lib.b2r_synth_num_rays(sensor)
    buf = np.empty((cap, 4), dtype=np.float32)
    n = lib.b2r_synth_scan(sensor, seed, scan_idx, buf.ctypes.data)
    return np.ascontiguousarray(buf[:n])


def pose(scan_idx: int, seed: int = DEFAULT_SEED) -> np.ndarray:
    """Ground-truth sensor pose (4x4 float64, world <- sensor) of a scan."""
    lib = _load()
    T = np.empty((4, 4), dtype=np.float64)
    lib.b2r_synth_pose(seed, scan_idx, T.ctypes.data)
    return T
