"""The device state machine of PCL's BFGS GICP (mrg_slam_b200/csrc/gicp_pcl_sm.hpp, plain host/device C++) compiled for the HOST
and driven by the oracle's own correspondence search and cost functor: it must reproduce the oracle's straight-line
implementation (oracle/gicp_pcl.cpp: orc_gicp_pcl_align) bit for bit — transform, converged flag, outer iterations, functor
evaluations.  This checks the cut-up control flow (outer loop, BFGS, bracketing / sectioning line search) without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from tests import oraclelib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    O.lib()
    so = str(tmp_path_factory.mktemp("gp") / "libgp_sm_host.so")
    cmd = ["g++", "-std=c++17", "-O2", "-msse4.2", "-fPIC", "-shared", "-Wall", os.path.join(ROOT, "tests/cpp/gicp_pcl_sm_host.cpp"), "-o", so,
           f"-L{ROOT}/oracle", "-loracle", f"-Wl,-rpath,{ROOT}/oracle"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    L = ctypes.CDLL(so)
    L.gp_sm_align.restype = ctypes.c_int
    L.gp_sm_align.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(O.GicpPclParams), ctypes.c_void_p,
                              ctypes.POINTER(O.Result), ctypes.POINTER(ctypes.c_int)]
    return L


def _sm_align(L, target, source, guess, params):
    t, s = np.ascontiguousarray(target, np.float32), np.ascontiguousarray(source, np.float32)
    g = O.colmajor(guess)
    r, rounds = O.Result(), ctypes.c_int()
    rc = L.gp_sm_align(t.ctypes.data, len(t), s.ctypes.data, len(s), ctypes.byref(params), g.ctypes.data, ctypes.byref(r), ctypes.byref(rounds))
    assert rc == 0, rc
    return r, rounds.value


@pytest.mark.parametrize("case", ["identity_guess", "offset_guess", "coarse", "tight", "few_inner", "one_outer"])
def test_state_machine_reproduces_the_oracle_bit_for_bit(harness, small_pair, case):
    a, b, gt = small_pair
    guess = np.eye(4)
    prm = O.gicp_pcl_params()
    if case == "offset_guess":
        guess = gt.copy()
        guess[0, 3] += 0.4
        guess[1, 3] -= 0.3
    elif case == "coarse":
        prm = O.gicp_pcl_params(transformation_epsilon=0.5, rotation_epsilon=0.05)
    elif case == "tight":
        prm = O.gicp_pcl_params(transformation_epsilon=1e-3, rotation_epsilon=1e-4, maximum_iterations=12)
    elif case == "few_inner":
        prm = O.gicp_pcl_params(max_optimizer_iterations=2, transformation_epsilon=0.01)
    elif case == "one_outer":
        prm = O.gicp_pcl_params(maximum_iterations=1)
    want = O.gicp_pcl_align(a, b, guess, prm)
    got, rounds = _sm_align(harness, a, b, guess, prm)
    assert (got.converged, got.iterations, got.lm_evals) == (want.converged, want.iterations, want.lm_evals), case
    assert list(got.T) == list(want.T), case
    assert rounds == want.lm_evals + want.iterations  # one round per request: every functor evaluation + one search per outer iteration


def test_state_machine_with_too_few_correspondences(harness, small_pair):
    a, b, _ = small_pair
    far = np.eye(4)
    far[0, 3] = 500.0  # nothing within max_correspondence_distance: estimateRigidTransformationBFGS throws, converged_ stays false
    prm = O.gicp_pcl_params()
    want = O.gicp_pcl_align(a, b, far, prm)
    got, _ = _sm_align(harness, a, b, far, prm)
    assert want.converged == 0 and got.converged == 0 and got.iterations == want.iterations == 0
    assert list(got.T) == list(want.T)
