"""Golden vectors (tests/golden/golden_v1.npz, written by tests/golden/make_golden.py).

CPU: the oracle still reproduces them (guards the checker itself).  GPU: the CUDA path matches the committed numbers
without consulting the oracle at run time — bit-exact for index / set / downsampled-cloud data (sha256), within the
north-star tolerances for transforms and fitness."""
import os

import numpy as np
import pytest

from tests.conftest import pose_error
from tests.golden import make_golden as G

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(PATH, allow_pickle=False))


def test_oracle_reproduces_golden(golden):
    now = G.build()
    assert set(now) == set(golden)
    for k, v in golden.items():
        if v.dtype.kind in "US" or v.dtype.kind in "iub":
            assert np.array_equal(now[k], v), k
        elif v.dtype == np.float32:  # final transforms: the oracle's OpenMP sum order depends on the thread count
            np.testing.assert_allclose(now[k], v, rtol=1e-5, atol=1e-6, err_msg=k)
        else:
            np.testing.assert_allclose(now[k], v, rtol=1e-8, atol=1e-12, err_msg=k)


@pytest.mark.gpu
def test_gpu_matches_golden(golden):
    from mrg_slam_b200 import lib as B
    from mrg_slam_b200 import synth

    raw_a, raw_b = G.inputs()
    assert G.sha(raw_a) == str(golden["raw_a_sha"]) and G.sha(raw_b) == str(golden["raw_b_sha"])  # generator unchanged
    reg = B.Registration(B.default_config(B.FAST_VGICP))
    d = reg.distance_filter(raw_a, 0.1, 35.0)
    assert (len(d), G.sha(d)) == (int(golden["dist_n"]), str(golden["dist_sha"]))
    v, ovf = reg.voxelgrid(d, 0.1, 1)
    assert not ovf and (len(v), G.sha(v)) == (int(golden["vg_n"]), str(golden["vg_sha"]))
    r = reg.radius_outlier(v, 0.5, 2)
    assert (len(r), G.sha(r)) == (int(golden["rad_n"]), str(golden["rad_sha"]))
    s = reg.statistical_outlier(v, 30, 1.2)
    assert (len(s), G.sha(s)) == (int(golden["sor_n"]), str(golden["sor_sha"]))
    A = r
    Bc = reg.prefilter(raw_b)
    gt = np.linalg.inv(synth.pose(G.SCAN_A)) @ synth.pose(G.SCAN_B)
    reg.setInputTarget(A); reg.setInputSource(Bc)
    cov, knn = reg.debug_covariances(0, want_knn=True)
    assert G.sha(np.sort(knn, 1)) == str(golden["knn_sorted_sha"])
    np.testing.assert_allclose(cov[:256], golden["cov_head"], atol=1e-11)
    coords, npts, mean, vcov = reg.debug_voxelmap()
    assert (len(coords), G.sha(coords), G.sha(npts)) == (int(golden["vox_n"]), str(golden["vox_coords_sha"]), str(golden["vox_npts_sha"]))
    np.testing.assert_allclose(mean[:64], golden["vox_mean_head"], atol=1e-12)
    np.testing.assert_allclose(vcov[:64], golden["vox_cov_head"], atol=1e-12)
    for res in (1.0, 0.5):
        tag = str(res).replace(".", "p")
        n = B.Registration(B.default_config(B.NDT_OMP, resolution=res))
        n.setInputTarget(A); n.setInputSource(Bc)
        idx, cnt, m, icov, min_b, div_b = n.debug_ndt_grid()
        assert (len(idx), G.sha(idx), G.sha(cnt)) == (int(golden[f"ndt_n_{tag}"]), str(golden[f"ndt_idx_sha_{tag}"]), str(golden[f"ndt_npts_sha_{tag}"]))
        assert np.array_equal(min_b, golden[f"ndt_min_b_{tag}"]) and np.array_equal(div_b, golden[f"ndt_div_b_{tag}"])
        n.close()
    for name, method in (("vgicp", B.FAST_VGICP), ("gicp", B.FAST_GICP), ("ndt", B.NDT_OMP)):
        g = B.Registration(B.default_config(method))
        g.setInputTarget(A); g.setInputSource(Bc)
        for i, guess in enumerate(G.guesses(gt)):
            res = g.align(guess)
            assert (res.converged, res.iterations) == (int(golden[f"{name}_conv"][i]), int(golden[f"{name}_iters"][i]))
            te, re = pose_error(B.from_colmajor(golden[f"{name}_T"][i]), g.getFinalTransformation())
            assert te <= 1e-4 and re <= 1e-4
            f = g.getFitnessScore()
            assert abs(f - golden[f"{name}_fitness"][i]) <= 1e-3 * golden[f"{name}_fitness"][i]
        if method != B.NDT_OMP:
            err, H, b, corr, valid = g.debug_linearize(gt)
            assert G.sha(corr[valid]) == str(golden[f"{name}_corr_sha"])  # correspondence set
            assert G.sha(valid.astype(bool)) == str(golden[f"{name}_valid_sha"])
            assert abs(err - golden[f"{name}_lin_err"]) <= 1e-9 * abs(golden[f"{name}_lin_err"])
            assert np.abs(H - golden[f"{name}_lin_H"]).max() <= 1e-9 * np.abs(H).max()
        g.close()
    reg.close()


# ------------------------------------------------------------------------------------------------ golden_v2
from tests.golden import make_golden_v2 as G2  # noqa: E402

PATH2 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v2.npz")


@pytest.fixture(scope="module")
def golden2():
    return dict(np.load(PATH2, allow_pickle=False))


def test_oracle_reproduces_golden_v2(golden2):
    now = G2.build()
    assert set(now) == set(golden2)
    for k, v in golden2.items():
        if v.dtype.kind in "US" or v.dtype.kind in "iub":
            assert np.array_equal(now[k], v), k
        elif v.dtype == np.float32:
            np.testing.assert_allclose(now[k], v, rtol=1e-5, atol=1e-6, err_msg=k)
        else:
            np.testing.assert_allclose(now[k], v, rtol=1e-8, atol=1e-12, err_msg=k)


@pytest.mark.gpu
def test_gpu_matches_golden_v2(golden2):
    from mrg_slam_b200 import lib as B
    from mrg_slam_b200 import synth
    from mrg_slam_b200.odometry import ScanMatchingOdometry

    reg = B.Registration(B.default_config(B.SMALL_GICP))
    pre = B.Registration(B.default_config(B.FAST_VGICP))
    A, Bc = pre.prefilter(synth.scan(synth.VLP16, 3)), pre.prefilter(synth.scan(synth.VLP16, 4))
    gt = np.linalg.inv(synth.pose(3)) @ synth.pose(4)
    reg.setInputTarget(A); reg.setInputSource(Bc)
    for i, guess in enumerate(G.guesses(gt)):
        res = reg.align(guess)
        assert (res.converged, res.iterations) == (int(golden2["sgicp_conv"][i]), int(golden2["sgicp_iters"][i]))
        te, re = pose_error(B.from_colmajor(golden2["sgicp_T"][i]), reg.getFinalTransformation())
        assert te <= 1e-4 and re <= 1e-4
        f = reg.getFitnessScore()
        assert abs(f - golden2["sgicp_fitness"][i]) <= 1e-3 * golden2["sgicp_fitness"][i]
    far = gt.copy(); far[:3, 3] += [0.8, -0.6, 0.1]
    err, H, b, corr, _ = reg.debug_linearize(far)
    assert G.sha(corr) == str(golden2["sgicp_corr_sha"])  # double-precision nearest neighbours + rejections
    assert abs(0.5 * err - golden2["sgicp_lin_err"]) <= 1e-9 * abs(golden2["sgicp_lin_err"])
    assert np.abs(H - golden2["sgicp_lin_H"]).max() <= 1e-9 * np.abs(H).max()
    clouds = [pre.prefilter(synth.scan(synth.VLP16, G2.MAP_FIRST + G2.MAP_STEP * i)) for i in range(G2.MAP_COUNT)]
    _, poses = G2.map_inputs()
    for tag, (res_, mp, far_) in {"fine": (0.05, 1, -1.0), "coarse": (0.5, 2, 25.0)}.items():
        m = pre.map_cloud(clouds, poses, [1, 0, 0, 0], res_, mp, far_, True)
        assert (len(m), G.sha(m)) == (int(golden2[f"map_{tag}_n"]), str(golden2[f"map_{tag}_sha"]))
    odo = ScanMatchingOdometry(pre, make_cloud=lambda pts: B.Cloud(pre, pts))
    for i in range(G2.ODO_COUNT):
        p = odo.matching(0.1 * i, pre.prefilter(synth.scan(synth.VLP16, G2.ODO_FIRST + i)))
        te, re = pose_error(golden2["odo_traj"][i].astype(np.float64), p.astype(np.float64))
        assert te <= 1e-4 and re <= 1e-4
    assert odo.keyframe_switches == int(golden2["odo_switches"])


# ------------------------------------------------------------------------------------------------ golden_v3: loop matching
PATH3 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v3.npz")


def _check_v3(now, golden, pose_tol, delta_tol, score_rtol):
    assert set(now) == set(golden)
    for k, v in golden.items():
        if k.endswith("_pose"):
            te, re = pose_error(now[k].astype(np.float64), v.astype(np.float64))
            assert te <= pose_tol and re <= pose_tol, (k, te, re)
        elif k.endswith("_deltas"):
            assert np.array_equal(now[k] < 0, v < 0), k  # the same checks ran
            np.testing.assert_allclose(now[k], v, rtol=0, atol=delta_tol, err_msg=k)
        elif k.endswith("_score"):
            np.testing.assert_allclose(now[k], v, rtol=score_rtol, err_msg=k)
        else:  # best candidate, decision, cloud sizes
            assert np.array_equal(now[k], v), k


def test_oracle_reproduces_golden_v3():
    from tests.golden import make_golden_v3 as G3

    _check_v3(G3.build(), dict(np.load(PATH3, allow_pickle=False)), 1e-6, 1e-6, 1e-6)


@pytest.mark.gpu
def test_gpu_matches_golden_v3():
    """LoopDetector::matching + consistency check through the product, against the committed decisions and deltas
    (no oracle at run time): transforms within the north-star tolerance, decisions identical."""
    from mrg_slam_b200 import lib as B
    from tests.golden import make_golden_v3 as G3

    class GpuBatch:
        def __init__(self):
            self.reg = B.Registration(B.default_config(B.FAST_GICP))

        def align_batch(self, sources, targets, guesses, with_fitness=False, fitness_max_range=np.finfo(np.float64).max):
            cache = {}

            def up(c):
                if id(c) not in cache:
                    cache[id(c)] = B.Cloud(self.reg, c)
                return cache[id(c)]

            out = self.reg.align_batch([up(s) for s in sources], [up(t) for t in targets], guesses, with_fitness=with_fitness,
                                       fitness_max_range=fitness_max_range)
            for c in cache.values():
                c.close()
            return out

    _check_v3(G3.build(GpuBatch), dict(np.load(PATH3, allow_pickle=False)), 1e-4, 2e-4, 1e-3)
