"""registration_method GICP / GICP_OMP (pcl / pclomp GeneralizedIterativeClosestPoint, /root/reference/src/mrg_slam/registrations.cpp:93-116;
SURVEY 8a row G) on the device, against the oracle restatement (oracle/gicp_pcl.cpp).

The state machine behind it is verified on the host bit for bit (tests/test_gicp_pcl_sm.py).  The two CUDA kernels that answer
its requests and PCL's covariance kernel are checked here.  Tolerances: BFGS with PCL's coarse stopping rule
(|delta T| < transformation_epsilon = 0.1 m / rotation_epsilon = 2e-3 per outer iteration) stops wherever the last outer iteration
happens to land, so the default configuration is compared at centimetre level; the north-star tolerance (1e-4 m / 1e-4 rad)
does not apply to this method (see the note below): with a tight stopping rule the comparison is at the millimetre level, which
is the oracle's own reproducibility."""
import numpy as np
import pytest

from mrg_slam_b200 import lib as B
from tests import oraclelib as O
from tests.conftest import pose_error

# Status: all five tests passed on a B200 at the end of round 1 (GPUTEST_r01.json: 5 xpassed) with PCL-style covariances from
# pcl_cov_kernel; the marker that let them pass or fail silently is gone, a regression in pcl_cov_kernel / gicp_pcl_eval_kernel /
# gicp_pcl_step_kernel now turns the suite red.  The tight-stopping-rule cases use 2e-3 m / 1e-3 rad: the ORACLE's own result
# moves by 0.2-0.8 mm (and its outer iteration count by up to 2) when the initial guess is perturbed by 1e-7 m
# (tests/test_gicp_pcl.py::test_result_is_reproducible_only_to_about_a_millimetre).
pytestmark = pytest.mark.gpu


def _align(method_cfg, a, b, guess):
    g = B.Registration(method_cfg)
    g.setInputTarget(a)
    g.setInputSource(b)
    r = g.align(guess)
    T = g.getFinalTransformation()
    fit = g.getFitnessScore()
    g.close()
    return r, T, fit


def test_gicp_pcl_default_parameters(small_pair):
    a, b, gt = small_pair
    want = O.gicp_pcl_align(a, b, np.eye(4))
    r, T, fit = _align(B.default_config(B.GICP_PCL), a, b, np.eye(4))
    assert r.converged == want.converged == 1
    te, re = pose_error(O.from_colmajor(list(want.T)), T)
    assert te < 0.03 and re < 5e-3, (te, re)
    assert abs(r.iterations - want.iterations) <= 1
    te_gt, _ = pose_error(gt, T)
    assert te_gt < 0.15 and fit > 0


@pytest.mark.parametrize("guess_offset", [(0.0, 0.0), (0.3, -0.2)])
def test_gicp_pcl_tight_stopping_rule_matches_oracle(small_pair, guess_offset):
    a, b, gt = small_pair
    guess = np.eye(4)
    guess[0, 3], guess[1, 3] = gt[0, 3] + guess_offset[0], gt[1, 3] + guess_offset[1]
    kw = dict(transformation_epsilon=2e-5, rotation_epsilon=2e-6, maximum_iterations=40)
    want = O.gicp_pcl_align(a, b, guess, O.gicp_pcl_params(**kw))
    r, T, _ = _align(B.default_config(B.GICP_PCL, **kw), a, b, guess)
    te, re = pose_error(O.from_colmajor(list(want.T)), T)
    assert te <= 2e-3 and re <= 1e-3, (te, re, r.iterations, want.iterations)


def test_gicp_pcl_batch_and_factory(small_pair):
    a, b, gt = small_pair
    reg = B.select_registration_method({"registration_method": "GICP", "reg_max_optimizer_iterations": 20})
    ca, cb = B.Cloud(reg, a), B.Cloud(reg, b)
    far = np.eye(4)
    far[0, 3] = 500.0
    res = reg.align_batch([cb, cb, ca], [ca, ca, ca], [np.eye(4), far, np.eye(4)], with_fitness=True)
    reg.setInputTarget(ca); reg.setInputSource(cb)
    single = reg.align(np.eye(4))
    assert list(res[0].T) == list(single.T) and res[0].iterations == single.iterations  # batch == single, bit for bit
    assert res[1].converged == 0 and res[1].iterations == 0                              # no correspondences: BFGS "throws"
    assert np.allclose(np.array(list(res[1].T)).reshape(4, 4).T, far, atol=1e-6)        # final = previous * guess = guess
    assert res[2].converged == 1                                                         # a cloud against itself
    te, re = pose_error(np.eye(4), B.from_colmajor(list(res[2].T)))
    assert te < 1e-3 and re < 1e-3
    ca.close(); cb.close(); reg.close()


def test_gicp_pcl_covariances_match_oracle(small_pair):
    """pcl_cov_kernel (PCL's moment arithmetic over the exact-kNN lists) against the oracle's computeCovariances restatement."""
    a, b, _ = small_pair
    g = B.Registration(B.default_config(B.GICP_PCL))
    g.setInputTarget(a)
    g.setInputSource(b)
    got = g.debug_covariances(1)  # target
    want = O.gicp_pcl_covariances(a, k=20)
    want6 = np.stack([want[:, 0, 0], want[:, 0, 1], want[:, 0, 2], want[:, 1, 1], want[:, 1, 2], want[:, 2, 2]], axis=1)
    # identical neighbour sets and float products; the double sums run in a different order and the eigenvectors come from
    # different Jacobi sweeps: agreement to ~1e-9, except where two eigenvalues are nearly equal (ambiguous plane normal)
    close = np.isclose(got, want6, rtol=0, atol=1e-6).all(axis=1)
    assert close.mean() > 0.995, close.mean()
    w = np.linalg.eigvalsh(np.stack([[got[:, 0], got[:, 1], got[:, 2]], [got[:, 1], got[:, 3], got[:, 4]], [got[:, 2], got[:, 4], got[:, 5]]]).transpose(2, 0, 1))
    assert np.allclose(w, np.tile([1e-3, 1.0, 1.0], (len(w), 1)), atol=1e-9)
    g.close()
