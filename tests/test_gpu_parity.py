"""GPU parity tests: every call goes through the C ABI (libb2r.so via ctypes) and is compared with the CPU oracle on
the same seeded inputs.  Bars (BASELINE.json north_star): voxel indices, correspondence sets, kNN sets and
downsampled point sets BIT-EXACT; final transforms within 1e-4 m / 1e-4 rad; fitness within 1e-3 relative (we hold
1e-9); plus identical converged flag and iteration count."""
import ctypes

import numpy as np
import pytest

from mrg_slam_b200 import lib as B
from mrg_slam_b200 import loop_closure as LC
from mrg_slam_b200 import synth
from tests import oraclelib as O
from tests.conftest import oracle_prefilter, pose_error

pytestmark = pytest.mark.gpu

T_TOL, R_TOL, FIT_RTOL = 1e-4, 1e-4, 1e-3  # north_star tolerances


@pytest.fixture(scope="module")
def reg():
    r = B.Registration(B.default_config(B.FAST_VGICP))
    yield r
    r.close()


# ------------------------------------------------------------------------------------------------ filters
@pytest.mark.parametrize("sensor,idx", [(synth.VLP16, 3), (synth.HDL64, 5)])
def test_filter_chain_bit_exact(reg, sensor, idx):
    raw = synth.scan(sensor, idx)
    o1 = O.distance_filter(raw, 0.1, 35.0)
    g1 = reg.distance_filter(raw, 0.1, 35.0)
    assert np.array_equal(o1, g1)
    o2, ovf, vidx = O.voxelgrid(o1, 0.1, 1, want_index=True)
    g2, govf = reg.voxelgrid(o1, 0.1, 1)
    assert not ovf and not govf
    assert np.array_equal(o2, g2)  # same voxels, same order, same float32 centroids
    assert np.all(np.diff(vidx) > 0)
    keep = O.radius_outlier(o2, 0.5, 2)
    assert np.array_equal(o2[keep], reg.radius_outlier(o2, 0.5, 2))
    keep_s, _, _ = O.statistical_outlier(o2, 30, 1.2)
    assert np.array_equal(o2[keep_s], reg.statistical_outlier(o2, 30, 1.2))
    # the fused chain (cloud_callback order, intermediates on the device) gives the same cloud
    assert np.array_equal(o2[keep], reg.prefilter(raw))


def test_prefilter_full_size_os1_1m(reg):
    """BASELINE config 3: 128 x 8192 = 1,048,576-ray cloud."""
    raw = synth.scan(synth.OS1_128_1M, 2)
    assert len(raw) > 700_000
    o1 = O.distance_filter(raw, 0.1, 35.0)
    o2, ovf = O.voxelgrid(o1, 0.1, 1)
    keep = O.radius_outlier(o2, 0.5, 2)
    assert np.array_equal(o2[keep], reg.prefilter(raw))
    # without the distance filter the 120 m extent at 0.1 m still fits INT32 here; at 0.01 m it overflows:
    # PCL warns and returns the input unchanged
    o3, ovf3 = O.voxelgrid(raw, 0.01, 1)
    g3, govf3 = reg.voxelgrid(raw, 0.01, 1)
    assert ovf3 and govf3 and np.array_equal(g3, raw) and np.array_equal(o3, raw)


def test_filter_edge_cases(reg):
    raw = synth.scan(synth.VLP16, 9)
    c = O.distance_filter(raw, 0.1, 35.0)
    # min_points_per_voxel > 1
    o, _ = O.voxelgrid(c, 0.3, 3)
    g, _ = reg.voxelgrid(c, 0.3, 3)
    assert np.array_equal(o, g) and 0 < len(g) < len(c)
    # RadiusOutlierRemoval special case min_pts == 1 (nearestKSearch(2))
    v, _ = O.voxelgrid(c, 0.2, 1)
    assert np.array_equal(v[O.radius_outlier(v, 0.3, 1)], reg.radius_outlier(v, 0.3, 1))
    assert np.array_equal(v[O.radius_outlier(v, 0.8, 5)], reg.radius_outlier(v, 0.8, 5))
    # statistical with the code defaults of prefiltering_component.cpp:100-104 (k=20, 1.0)
    ks, _, _ = O.statistical_outlier(v, 20, 1.0)
    assert np.array_equal(v[ks], reg.statistical_outlier(v, 20, 1.0))
    # a leaf of many metres: thousands of points per voxel — the sort's crowded-cell fallback (bitonic) must give the same cloud
    for leaf in (8.0, 60.0):
        o, _ = O.voxelgrid(c, leaf, 1)
        g, _ = reg.voxelgrid(c, leaf, 1)
        assert np.array_equal(o, g) and 0 < len(g) < 200
    # empty and tiny inputs
    empty = np.zeros((0, 4), np.float32)
    assert len(reg.distance_filter(empty, 0.1, 35.0)) == 0
    assert len(reg.voxelgrid(empty, 0.1, 1)[0]) == 0
    assert len(reg.radius_outlier(empty, 0.5, 2)) == 0
    one = np.array([[1.0, 2.0, 3.0, 0.5]], np.float32)
    assert np.array_equal(reg.voxelgrid(one, 0.1, 1)[0], one)
    assert len(reg.radius_outlier(one, 0.5, 2)) == 0
    # non-finite points: dropped by the distance filter and ignored by VoxelGrid
    bad = c[:2000].copy()
    bad[5, 0] = np.nan
    bad[77, 2] = np.inf
    assert np.array_equal(O.distance_filter(bad, 0.1, 35.0), reg.distance_filter(bad, 0.1, 35.0))
    assert np.array_equal(O.voxelgrid(bad, 0.2, 1)[0], reg.voxelgrid(bad, 0.2, 1)[0])
    # pcl::PointXYZI layout (32-byte stride, intensity at byte 16)
    wide = np.zeros((len(c), 8), np.float32)
    wide[:, :3] = c[:, :3]; wide[:, 3] = 1.0; wide[:, 4] = c[:, 3]
    assert np.array_equal(reg.voxelgrid(wide, 0.1, 1)[0], reg.voxelgrid(c, 0.1, 1)[0])


# ------------------------------------------------------------------------------------------------ exact kNN
def test_knn_bit_exact(reg, vlp16_pair):
    a, _, _ = vlp16_pair
    cl = B.Cloud(reg, a)
    rng = np.random.default_rng(0)
    q = np.concatenate([a[rng.integers(0, len(a), 2000)], (rng.normal(size=(500, 4)) * 25).astype(np.float32),
                        np.array([[500.0, -300.0, 40.0, 0.0]], np.float32)])
    for k in (1, 5, 20, 31):
        oi, od = O.knn(a, q, k)
        gi, gd = reg.debug_knn(cl, q, k)
        assert np.array_equal(od, gd), k  # FLANN float association reproduced
        same = (oi == gi).all(1)
        # indices can only differ where neighbours have bit-identical distances: swapped inside the list, or a tie
        # between the k-th and the (k+1)-th neighbour (either is a correct answer)
        for r in np.where(~same)[0]:
            diff = set(oi[r]) ^ set(gi[r])
            for i in diff:
                d = a[i, :3] - q[r, :3]
                dd = np.float32(np.float32(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])) + np.float32(d[2] * d[2]))
                assert dd == od[r, -1]
            assert diff or len(set(od[r])) < k
        assert (~same).sum() <= 5
    cl.close()


def test_knn_cov_tma_tile_variant_matches(vlp16_pair):
    """B2R_KNN_TILE=1 selects knn_cov_tile_kernel (candidate rows staged in shared memory by cp.async.bulk): the same neighbour
    sets and covariances as the default kernel.  The switch is read once per process, hence the subprocess."""
    import os, subprocess, sys, tempfile
    a, b, _ = vlp16_pair
    with tempfile.TemporaryDirectory() as d:
        np.save(os.path.join(d, "b.npy"), b)
        code = ("import sys, numpy as np; sys.path.insert(0, %r); from mrg_slam_b200 import lib as B\n"
                "b = np.load(%r)\n"
                "g = B.Registration(B.default_config(B.FAST_GICP)); g.setInputTarget(b); g.setInputSource(b)\n"
                "cov, knn = g.debug_covariances(0, want_knn=True); np.save(%r, cov); np.save(%r, np.sort(knn, 1))\n"
                "g2 = B.Registration(B.default_config(B.FAST_GICP, nn_cell_size=3.0)); g2.setInputTarget(b); g2.setInputSource(b)\n"
                "np.save(%r, g2.debug_covariances(0))\n"
                % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(d, "b.npy"), os.path.join(d, "cov.npy"),
                   os.path.join(d, "knn.npy"), os.path.join(d, "cov3.npy")))
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, B2R_KNN_TILE="1"), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        ocov, oknn = O.knn_covariances(b, 20, want_idx=True)
        np.testing.assert_allclose(np.load(os.path.join(d, "cov.npy")), ocov, atol=1e-11)
        np.testing.assert_allclose(np.load(os.path.join(d, "cov3.npy")), ocov, atol=1e-11)  # coarse grid: strips exceed the tile
        assert np.array_equal(np.sort(oknn, 1), np.load(os.path.join(d, "knn.npy")))


def test_knn_cov_candidate_log_overflow_path(vlp16_pair):
    """knn_cov keeps the candidates that entered the top-k in a bounded shared-memory log; a coarse NN grid makes the log
    overflow, which must take the second-traversal path and give the same neighbours and covariances."""
    a, b, _ = vlp16_pair
    ocov, oknn = O.knn_covariances(b, 20, want_idx=True)
    for cell in (0.0, 3.0):
        g = B.Registration(B.default_config(B.FAST_GICP, nn_cell_size=cell))
        before = g.knn_list_overflows()
        g.setInputTarget(a); g.setInputSource(b)
        g.align(np.eye(4))  # builds covariances through the production path (no neighbour export)
        gcov = g.debug_covariances(0)
        np.testing.assert_allclose(gcov, ocov, atol=1e-11)
        took_fallback = g.knn_list_overflows() - before
        if cell > 0:
            assert took_fallback > 0
        _, gknn = g.debug_covariances(0, want_knn=True)
        assert np.array_equal(np.sort(oknn, 1), np.sort(gknn, 1))


# ------------------------------------------------------------------------------------------------ GICP / VGICP
@pytest.mark.parametrize("method", [B.FAST_VGICP, B.FAST_GICP])
def test_lsq_intermediates(vlp16_pair, method):
    a, b, gt = vlp16_pair
    g = B.Registration(B.default_config(method))
    o = O.Registration(O.default_params(method))
    g.setInputTarget(a); g.setInputSource(b)
    o.setInputTarget(a); o.setInputSource(b)
    ocov, oknn = O.knn_covariances(b, 20, want_idx=True)
    gcov, gknn = g.debug_covariances(0, want_knn=True)
    assert np.array_equal(np.sort(oknn, 1), np.sort(gknn, 1))  # kNN sets bit-exact
    np.testing.assert_allclose(gcov, ocov, atol=1e-11)
    if method == B.FAST_VGICP:
        oc, on, om, ov = O.vgicp_voxelmap(a, O.knn_covariances(a, 20), 1.0)
        gc, gn, gm, gv = g.debug_voxelmap()
        assert np.array_equal(oc, gc) and np.array_equal(on, gn)  # voxel indices and counts bit-exact
        np.testing.assert_allclose(gm, om, rtol=0, atol=1e-12)
        np.testing.assert_allclose(gv, ov, rtol=0, atol=1e-12)
    for T in (np.eye(4), gt):
        oe, oH, ob, ocorr, oval = o.linearize(T)
        ge, gH, gb, gcorr, gval = g.debug_linearize(T)
        assert np.array_equal(oval, gval) and np.array_equal(ocorr[oval], gcorr[gval])  # correspondence set bit-exact
        assert abs(oe - ge) <= 1e-9 * abs(oe)
        assert np.abs(oH - gH).max() <= 1e-9 * np.abs(oH).max()
        assert np.abs(ob - gb).max() <= 1e-9 * np.abs(ob).max()
    # compute_error keeps the correspondences / Mahalanobis matrices of the linearisation pose
    T1 = gt.copy(); T1[1, 3] += 0.03
    o.linearize(gt)
    assert abs(o.compute_error(T1) - g.debug_compute_error(gt, T1)) <= 1e-9 * abs(o.compute_error(T1))
    g.close()


GUESS_OFFSETS = [(0.0, 0.0), (0.3, 0.0), (-0.2, 0.02), (0.45, -0.01)]


@pytest.mark.parametrize("method", [B.FAST_VGICP, B.FAST_GICP, B.NDT_OMP, B.SMALL_GICP])
def test_align_matches_oracle(vlp16_pair, method):
    a, b, gt = vlp16_pair
    g = B.Registration(B.default_config(method))
    o = O.Registration(O.default_params(method))
    g.setInputTarget(a); g.setInputSource(b)
    o.setInputTarget(a); o.setInputSource(b)
    for dx, dyaw in GUESS_OFFSETS + [(None, None)]:
        if dx is None:
            guess = np.eye(4)
        else:
            guess = gt.copy(); guess[0, 3] -= dx
            c, s = np.cos(dyaw), np.sin(dyaw)
            guess[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]) @ guess[:3, :3]
        ro, rg = o.align(guess), g.align(guess)
        assert rg.converged == ro.converged and rg.iterations == ro.iterations and rg.evals == ro.lm_evals
        te, re = pose_error(o.getFinalTransformation(), g.getFinalTransformation())
        assert te <= T_TOL and re <= R_TOL, (te, re)
        assert abs(rg.error - ro.error) <= 1e-6 * max(1.0, abs(ro.error))
        assert g.hasConverged() == o.hasConverged()
        for mr in (np.finfo(np.float64).max, 1.0, 0.05):
            fo, fg = o.getFitnessScore(mr), g.getFitnessScore(mr)
            assert abs(fo - fg) <= FIT_RTOL * abs(fo), (mr, fo, fg)
            # (SMALL_GICP accumulates in the source frame: the double pose agrees to ~1e-13, its float cast may differ by an ulp)
            assert abs(fo - fg) <= (1e-7 if method == B.SMALL_GICP else 1e-9) * abs(fo)
        # the `output` cloud of align(): float transform with PCL's association, bit-exact
        assert np.array_equal(g.aligned_cloud(), O.transform_cloud(b, g.getFinalTransformation()))
    g.close()


@pytest.mark.parametrize("method,over", [
    (B.FAST_VGICP, dict(neighbor_search=B.DIRECT7)),
    (B.FAST_VGICP, dict(resolution=0.5, transformation_epsilon=0.01)),
    (B.FAST_VGICP, dict(correspondence_randomness=10)),
    (B.FAST_GICP, dict(max_correspondence_distance=0.5, transformation_epsilon=0.001)),
    (B.SMALL_GICP, dict(max_correspondence_distance=0.5, transformation_epsilon=0.001, rotation_epsilon=1e-4)),
    (B.SMALL_GICP, dict(correspondence_randomness=10, maximum_iterations=2, transformation_epsilon=1e-6)),
    (B.NDT_OMP, dict(resolution=0.5)),
    (B.NDT_OMP, dict(neighbor_search=B.DIRECT1, transformation_epsilon=0.01)),
    (B.NDT_OMP, dict(resolution=2.0, neighbor_search=B.DIRECT27, maximum_iterations=5)),
    (B.NDT_OMP, dict(neighbor_search=B.KDTREE)),                                          # registrations.cpp:140-141
    (B.NDT_OMP, dict(resolution=2.0, neighbor_search=B.KDTREE, transformation_epsilon=0.01, maximum_iterations=8)),
])
def test_align_parameter_variants(vlp16_pair, method, over):
    a, b, gt = vlp16_pair
    g = B.Registration(B.default_config(method, **over))
    o = O.Registration(O.default_params(method, **over))
    g.setInputTarget(a); g.setInputSource(b)
    o.setInputTarget(a); o.setInputSource(b)
    guess = gt.copy(); guess[0, 3] -= 0.25; guess[1, 3] += 0.1
    ro, rg = o.align(guess), g.align(guess)
    assert (rg.converged, rg.iterations, rg.evals) == (ro.converged, ro.iterations, ro.lm_evals)
    te, re = pose_error(o.getFinalTransformation(), g.getFinalTransformation())
    assert te <= T_TOL and re <= R_TOL, (te, re)
    g.close()


def test_ndt_intermediates(vlp16_pair):
    a, b, gt = vlp16_pair
    for res in (1.0, 0.5):
        g = B.Registration(B.default_config(B.NDT_OMP, resolution=res))
        o = O.Registration(O.default_params(O.NDT_OMP, resolution=res))
        g.setInputTarget(a); g.setInputSource(b)
        o.setInputTarget(a); o.setInputSource(b)
        oi, on, om, oic, omin, odiv = O.ndt_grid(a, res)
        gi, gn, gm, gic, gmin, gdiv = g.debug_ndt_grid()
        assert np.array_equal(oi, gi) and np.array_equal(on, gn)  # voxel indices, counts and usable flags bit-exact
        assert np.array_equal(omin, gmin) and np.array_equal(odiv, gdiv)
        np.testing.assert_allclose(gm, om, rtol=0, atol=1e-12)
        use = on >= 6
        assert np.abs(oic - gic)[use].max() <= 1e-8 * np.abs(oic[use]).max()
        for p in (np.zeros(6), np.array([0.4, 0.01, 0.0, 0.001, -0.002, 0.01]), np.array([0.1, 0.0, 0.0, 3.13, 3.1, -3.0])):
            os_, og, oH, ohits = o.ndt_derivatives(p)
            gs_, gg, gH, ghits = g.debug_ndt_derivatives(p)
            assert np.array_equal(ohits, ghits)  # per-point voxel hit counts bit-exact
            assert abs(os_ - gs_) <= 1e-7 * abs(os_)
            assert np.abs(og - gg).max() <= 1e-6 * np.abs(og).max()
            assert np.abs(oH - gH).max() <= 1e-6 * np.abs(oH).max()
        g.close()
        # KDTREE neighbourhoods (radius search over the leaves' float centroids): the same per-point hit counts
        gk = B.Registration(B.default_config(B.NDT_OMP, resolution=res, neighbor_search=B.KDTREE))
        ok = O.Registration(O.default_params(O.NDT_OMP, resolution=res, neighbor_search=O.KDTREE))
        gk.setInputTarget(a); gk.setInputSource(b)
        ok.setInputTarget(a); ok.setInputSource(b)
        for p in (np.zeros(6), np.array([0.4, 0.01, 0.0, 0.001, -0.002, 0.01])):
            os_, og, oH, ohits = ok.ndt_derivatives(p)
            gs_, gg, gH, ghits = gk.debug_ndt_derivatives(p)
            assert np.array_equal(ohits, ghits) and ohits.max() >= 3
            assert abs(os_ - gs_) <= 1e-7 * abs(os_)
            assert np.abs(og - gg).max() <= 1e-6 * np.abs(og).max()
            assert np.abs(oH - gH).max() <= 1e-6 * np.abs(oH).max()
        gk.close()


def test_small_gicp_intermediates(vlp16_pair):
    """small_gicp GICPFactor: double-precision nearest neighbours bit-exact, H / b / e of the linearisation and the
    error of a trial pose (correspondences and Mahalanobis matrices of the linearisation pose) against the oracle."""
    a, b, gt = vlp16_pair
    g = B.Registration(B.default_config(B.SMALL_GICP))
    o = O.Registration(O.default_params(O.SMALL_GICP))
    g.setInputTarget(a); g.setInputSource(b)
    o.setInputTarget(a); o.setInputSource(b)
    far = gt.copy(); far[:3, 3] += [0.8, -0.6, 0.1]
    for T in (np.eye(4), gt, far):
        oe, oH, ob, ocorr, _ = o.linearize(T)
        ge, gH, gb, gcorr, _ = g.debug_linearize(T)
        assert np.array_equal(ocorr, gcorr)  # correspondence set bit-exact (incl. the max_dist_sq rejections)
        # debug_linearize reports sum r^T M r; GICPFactor's error carries the factor 1/2
        assert abs(oe - 0.5 * ge) <= 1e-9 * abs(oe)
        assert np.abs(oH - gH).max() <= 1e-9 * np.abs(oH).max()
        assert np.abs(ob - gb).max() <= 1e-9 * np.abs(ob).max()
    T1 = gt.copy(); T1[1, 3] += 0.03
    o.linearize(gt)
    assert abs(o.compute_error(T1) - 0.5 * g.debug_compute_error(gt, T1)) <= 1e-9 * abs(o.compute_error(T1))
    g.close()


# ------------------------------------------------------------------------------------------------ batch path
@pytest.mark.parametrize("method", [B.FAST_VGICP, B.FAST_GICP, B.NDT_OMP, B.SMALL_GICP])
def test_batch_equals_single_and_oracle(method):
    """LoopDetector::matching shape: shared targets, several candidates each (loop_detector.cpp:104-145)."""
    scans = [oracle_prefilter(synth.scan(synth.VLP16, 20 + i)) for i in range(5)]
    poses = [synth.pose(20 + i) for i in range(5)]
    g = B.Registration(B.default_config(method))
    clouds = [B.Cloud(g, s) for s in scans]
    pairs = [(0, 1), (0, 2), (0, 3), (4, 3), (4, 2), (1, 1)]  # (target, source); last one aligns a cloud with itself
    rng = np.random.default_rng(5)
    guesses = []
    for t, s in pairs:
        gt = np.linalg.inv(poses[t]) @ poses[s]
        gt[:3, 3] += rng.uniform(-0.15, 0.15, 3) * [1, 1, 0.2]
        guesses.append(gt)
    res = g.align_batch([clouds[s] for _, s in pairs], [clouds[t] for t, _ in pairs], guesses, with_fitness=True)
    res2 = g.align_batch([clouds[s] for _, s in pairs], [clouds[t] for t, _ in pairs], guesses, with_fitness=True)
    for i, (t, s) in enumerate(pairs):
        # run-to-run determinism: bitwise identical
        assert list(res[i].T) == list(res2[i].T) and res[i].fitness == res2[i].fitness
        # same as the single-pair call
        g.setInputTarget(clouds[t]); g.setInputSource(clouds[s])
        r1 = g.align(guesses[i])
        assert list(r1.T) == list(res[i].T) and r1.iterations == res[i].iterations
        assert g.getFitnessScore() == res[i].fitness
        # and as the oracle
        o = O.Registration(O.default_params(method))
        o.setInputTarget(scans[t]); o.setInputSource(scans[s])
        ro = o.align(guesses[i])
        assert (res[i].converged, res[i].iterations) == (ro.converged, ro.iterations)
        te, re = pose_error(o.getFinalTransformation(), B.from_colmajor(list(res[i].T)))
        assert te <= T_TOL and re <= R_TOL
        fo = o.getFitnessScore()
        assert abs(fo - res[i].fitness) <= (1e-7 if method == B.SMALL_GICP else 1e-9) * max(abs(fo), 1e-12)
    # the candidate reduction on top (tie rule, threshold) through the host mirror of LoopDetector::matching
    loops, table = LC.detect_loops(g, clouds, pairs[:5], guesses[:5])
    assert [l.target for l in loops] == [0, 4]
    for l, idxs in zip(loops, ([0, 1, 2], [3, 4])):
        best, score = LC.select_best([res[i].fitness for i in idxs], [bool(res[i].converged) for i in idxs])
        assert l.best_candidate == (best if score <= 1.25 else None)
    for c in clouds:
        c.close()
    g.close()


def test_information_matrix_fitness(reg, vlp16_pair):
    """InformationMatrixCalculator::calc_fitness_score (information_matrix_calculator.cpp:46-81) as a GPU call."""
    a, b, gt = vlp16_pair
    ca, cb = B.Cloud(reg, a), B.Cloud(reg, b)
    for T, mr in ((gt, np.finfo(np.float64).max), (np.eye(4), 2.0), (gt, 0.01)):
        fo, _ = O.fitness_score(a, b, T, mr)
        fg = reg.fitness_pair(ca, cb, T, mr)
        assert abs(fo - fg) <= 1e-9 * abs(fo)
    far = b.copy(); far[:, :3] += 500.0
    cf = B.Cloud(reg, far)
    assert reg.fitness_pair(ca, cf, np.eye(4), 1.0) == np.finfo(np.float64).max  # no match within range -> DBL_MAX
    for c in (ca, cb, cf):
        c.close()


def test_nearest_neighbor_search_small_and_odd_clouds(reg):
    """The 1-NN sweeps read the target two points at a time from a pair-interleaved copy (knn.cuh VISIT_PAIRS): clouds of 1, 2, 3 ...
    points, odd sizes (the +inf pad), duplicate points, points on one line (a single grid row) and queries far outside the target's
    box must give the brute-force answer, bit for bit (FLANN's float association)."""
    rng = np.random.default_rng(20261017)

    def brute(tgt, q):
        d = q[:, None, :3].astype(np.float32) - tgt[None, :, :3].astype(np.float32)
        d2 = ((d[..., 0] * d[..., 0]).astype(np.float32) + (d[..., 1] * d[..., 1]).astype(np.float32)).astype(np.float32)
        return (d2 + (d[..., 2] * d[..., 2]).astype(np.float32)).astype(np.float32).min(axis=1)

    def cloud(n, kind):
        c = np.zeros((n, 4), dtype=np.float32)
        if kind == "box":
            c[:, :3] = rng.uniform(-5, 5, (n, 3))
        elif kind == "line":  # one row of the grid: everything along x
            c[:, 0] = rng.uniform(-20, 20, n)
        elif kind == "dup":   # many exact duplicates
            c[:, :3] = rng.integers(-2, 3, (n, 3)).astype(np.float32)
        else:                 # flat, like a ground plane
            c[:, :2] = rng.uniform(-8, 8, (n, 2))
        return c

    g = reg
    for kind in ("box", "line", "dup", "plane"):
        for n in (1, 2, 3, 4, 5, 7, 8, 31, 32, 33, 257, 1001):
            tgt = cloud(n, kind)
            q = cloud(97, kind)
            q[:8, :3] += 40.0          # far outside the target's box
            q[8:16, :3] = tgt[rng.integers(0, n, 8), :3]  # exactly on target points
            ct, cq = B.Cloud(g, tgt), B.Cloud(g, q)
            want = brute(tgt, q)
            f = g.fitness_pair(ct, cq, np.eye(4))
            assert abs(f - float(want.astype(np.float64).mean())) <= 1e-12 * max(1.0, abs(f)), (kind, n)
            ct.close(); cq.close()


def test_split_upload_of_pinned_batches_gives_identical_results():
    """A batch of >= 16 pinned host clouds is uploaded in two halves, the second one on a copy stream while the first half's search
    structures are built (cloud.cu: clouds_upload / api.cu: run_align).  The rows must be bit-identical to the same batch from
    pageable host memory (single synchronous path), also when the clouds are released without ever being aligned."""
    torch = pytest.importorskip("torch")
    g = B.Registration(B.default_config(B.FAST_VGICP))
    scans = [oracle_prefilter(synth.scan(synth.VLP16, 40 + i)) for i in range(18)]
    poses = [synth.pose(40 + i) for i in range(18)]
    pairs = [(t, s) for t in range(2, 16) for s in (t - 2, t + 1, t + 2)]
    guesses = []
    for t, s in pairs:
        gt = np.linalg.inv(poses[t]) @ poses[s]
        gt[0, 3] += 0.2
        guesses.append(gt)
    pinned = [torch.from_numpy(c).pin_memory() for c in scans]
    # never aligned, released at once: the pending half must not be freed under the running copy
    for c in B.create_clouds(g, [p.data_ptr() for p in pinned], [p.shape[0] for p in pinned], B.HOST):
        c.close()
    rows = []
    for bufs, memspace in ((pinned, B.HOST), (None, None), (pinned, B.HOST)):
        if bufs is None:
            cl = [B.Cloud(g, c) for c in scans]  # pageable numpy arrays, one by one
        else:
            cl = B.create_clouds(g, [p.data_ptr() for p in bufs], [p.shape[0] for p in bufs], memspace)
        res = g.align_batch([cl[s] for _, s in pairs], [cl[t] for t, _ in pairs], guesses, with_fitness=True)
        rows.append([(list(r.T), r.iterations, r.converged, r.fitness) for r in res])
        for c in cl:
            c.close()
    assert rows[0] == rows[1] == rows[2]
    g.close()


# ------------------------------------------------------------------------------------------------ size-independent properties
def test_properties_full_size_hdl64(reg):
    """KITTI-shape clouds (BASELINE config 2): properties that need no oracle."""
    a = reg.prefilter(synth.scan(synth.HDL64, 40))
    b = reg.prefilter(synth.scan(synth.HDL64, 41))
    assert 30_000 < len(a) < 90_000
    gt = np.linalg.inv(synth.pose(40)) @ synth.pose(41)
    for method in (B.FAST_VGICP, B.FAST_GICP, B.NDT_OMP):
        g = B.Registration(B.default_config(method))
        ca, cb = B.Cloud(g, a), B.Cloud(g, b)
        g.setInputTarget(ca); g.setInputSource(cb)
        guess = gt.copy(); guess[0, 3] -= 0.2
        r = g.align(guess)
        assert r.converged
        te, re = pose_error(gt, g.getFinalTransformation())
        # reg_transformation_epsilon = 0.1 m (config/mrg_slam.yaml:102) stops the optimisers within ~0.1 m of the optimum
        assert te < 0.12 and re < 0.01, (method, te, re)
        # a cloud against itself: fitness exactly 0 at identity, and the LSQ optimisers do not move
        g.setInputSource(ca)
        r0 = g.align(np.eye(4))
        if method != B.NDT_OMP:
            assert pose_error(np.eye(4), g.getFinalTransformation())[0] < 5e-3
        assert g.fitness_pair(ca, ca, np.eye(4)) == 0.0
        # source -> target promotion at a keyframe switch reuses the cached structures: identical result either way
        g.setInputTarget(cb); g.setInputSource(ca)
        r1 = g.align(np.linalg.inv(gt))
        g2 = B.Registration(B.default_config(method))
        g2.setInputTarget(b); g2.setInputSource(a)
        r2 = g2.align(np.linalg.inv(gt))
        assert list(r1.T) == list(r2.T)
        ca.close(); cb.close(); g.close(); g2.close()


# ------------------------------------------------------------------------------------------------ error behaviour
def test_error_behaviour():
    L = B.load()
    g = B.Registration(B.default_config(B.FAST_VGICP))
    r = B.Result()
    guess = B.colmajor(np.eye(4) * 1.0)
    # align before setInputSource/Target: status, converged = 0, T = guess (PCL style: callers branch on hasConverged)
    assert L.b2r_align(g._h, guess.ctypes.data, ctypes.byref(r)) == B.ERR_STATE
    assert r.converged == 0 and list(r.T) == list(guess)
    assert b"source and target" in L.b2r_last_error(g._h)
    tiny = np.random.default_rng(0).normal(size=(10, 4)).astype(np.float32)
    g.setInputTarget(tiny); g.setInputSource(tiny)
    with pytest.raises(B.B2RError) as e:  # fewer points than correspondence_randomness
        g.align(np.eye(4))
    assert e.value.status == B.ERR_INVALID_ARG
    # ... but inside a batch a degenerate candidate only loses itself (the reference aligns candidates one by one,
    # loop_detector.cpp:126-145): converged = 0, T = guess, fitness = DBL_MAX, and the other pairs run
    ok = synth.scan(synth.VLP16, 3)[::4]
    c_ok, c_tiny = B.Cloud(g, ok), B.Cloud(g, tiny)
    shifted = np.eye(4); shifted[0, 3] = 0.25
    res = g.align_batch([c_ok, c_tiny, c_ok], [c_ok, c_ok, c_tiny], [np.eye(4), shifted, np.eye(4)], with_fitness=True)
    assert res[0].converged == 1 and res[0].fitness < 1e-6
    for r_bad, g_bad in ((res[1], shifted), (res[2], np.eye(4))):
        assert r_bad.converged == 0 and r_bad.iterations == 0 and r_bad.fitness == np.finfo(np.float64).max
        assert list(r_bad.T) == list(B.colmajor(g_bad))
    c_ok.close(); c_tiny.close()
    with pytest.raises(B.B2RError):
        g.setInputSource(np.zeros((0, 4), np.float32))
    with pytest.raises(B.B2RError):
        B.Registration(B.default_config(B.FAST_VGICP, correspondence_randomness=64))
    with pytest.raises(B.B2RError):
        B.Registration(B.default_config(B.FAST_VGICP, device=99))
    with pytest.raises(B.B2RError):
        g.statistical_outlier(np.zeros((100, 4), np.float32), 40, 1.0)
    # a target far too large for a dense 1e-3 m voxel table
    spread = (np.random.default_rng(1).uniform(-100, 100, size=(1000, 4))).astype(np.float32)
    g3 = B.Registration(B.default_config(B.FAST_VGICP, resolution=0.01))
    g3.setInputTarget(spread); g3.setInputSource(spread)
    with pytest.raises(B.B2RError) as e:
        g3.align(np.eye(4))
    assert e.value.status == B.ERR_CAPACITY
    launches = g.kernel_launches()
    assert launches > 0
    g.close(); g3.close()


def test_cloud_may_outlive_its_handle(vlp16_pair):
    """b2r_cloud_destroy after b2r_destroy of the creating handle (Python GC order is arbitrary) must be safe, and a
    cloud created through one handle is usable from another handle on the same device."""
    a, b, _ = vlp16_pair
    r1 = B.Registration(B.default_config(B.FAST_VGICP))
    ca, cb = B.Cloud(r1, a), B.Cloud(r1, b)
    r1.setInputTarget(ca); r1.setInputSource(cb)
    ref = r1.align(np.eye(4))
    T1 = r1.getFinalTransformation()
    r1.close()
    r2 = B.Registration(B.default_config(B.FAST_VGICP))
    r2.setInputTarget(ca); r2.setInputSource(cb)
    res = r2.align(np.eye(4))
    assert res.iterations == ref.iterations and np.array_equal(T1, r2.getFinalTransformation())
    r2.close()
    ca.close(); cb.close()


def test_two_handles_on_two_threads(vlp16_pair):
    """SURVEY 8b "Threading": odometry and loop detection own one registration object each and run on different threads of
    one process.  Two handles (one CUDA stream each) driven concurrently from two host threads give bitwise the results
    of the same calls made one after the other (bench.py's overlapped e2e leg relies on this)."""
    import threading
    a, b, gt = vlp16_pair
    methods = [B.FAST_VGICP, B.NDT_OMP]
    guesses = [gt.copy() for _ in range(4)]
    for k, g in enumerate(guesses):
        g[0, 3] += 0.05 * k
    serial, regs = [], []
    for m in methods:
        r = B.Registration(B.default_config(m))
        regs.append(r)
        ca, cb = B.Cloud(r, a), B.Cloud(r, b)
        serial.append([list(x.T) + [x.iterations, x.fitness] for x in r.align_batch([cb] * 4, [ca] * 4, guesses, with_fitness=True)])
        ca.close(); cb.close()
    out = [[None] * 6 for _ in methods]
    errors = []

    def work(i):
        try:
            r = regs[i]
            for rep in range(6):
                ca, cb = B.Cloud(r, a), B.Cloud(r, b)  # uploads and structure builds race with the other thread's kernels
                out[i][rep] = [list(x.T) + [x.iterations, x.fitness] for x in r.align_batch([cb] * 4, [ca] * 4, guesses, with_fitness=True)]
                ca.close(); cb.close()
        except Exception as e:  # pragma: no cover
            errors.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(methods))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    for i in range(len(methods)):
        for rep in range(6):
            assert out[i][rep] == serial[i]
    for r in regs:
        r.close()


def test_prefilter_chain_variants_match_separate_filters(reg):
    """b2r_prefilter folds the distance filter into VoxelGrid's own passes (no compaction in between); every variant of
    the chain must still equal the separate filters applied one after the other, which are bit-exact against the oracle
    above — including the VoxelGrid overflow branch (output = the distance-filtered input) and a range nobody survives."""
    raw = synth.scan(synth.VLP16, 12)
    raw = raw.copy()
    raw[7, 1] = np.nan
    raw[99, 0] = np.inf

    def cfg(**kw):
        c = B.PrefilterConfig()
        B.load().b2r_default_prefilter_config(ctypes.byref(c))
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    def separate(c):
        cur = raw
        if c.enable_distance_filter:
            cur = reg.distance_filter(cur, c.distance_near_thresh, c.distance_far_thresh)
        if c.downsample_method == 1:
            cur, _ = reg.voxelgrid(cur, c.downsample_resolution, c.downsample_min_points_per_voxel)
        if c.outlier_removal_method == 2:
            cur = reg.radius_outlier(cur, c.radius_radius, c.radius_min_neighbors)
        elif c.outlier_removal_method == 1:
            cur = reg.statistical_outlier(cur, c.statistical_mean_k, c.statistical_stddev)
        return cur

    variants = [cfg(), cfg(distance_near_thresh=2.0, distance_far_thresh=12.0), cfg(downsample_resolution=0.35, downsample_min_points_per_voxel=2),
                cfg(enable_distance_filter=0), cfg(outlier_removal_method=1), cfg(outlier_removal_method=0),
                cfg(downsample_resolution=0.001, outlier_removal_method=0),            # INT32 overflow: VoxelGrid passes its input through
                cfg(distance_near_thresh=500.0, distance_far_thresh=600.0)]            # nothing survives
    for c in variants:
        want, got = separate(c), reg.prefilter(raw, c)
        assert np.array_equal(want, got), (c.distance_near_thresh, c.downsample_resolution, c.outlier_removal_method, len(want), len(got))
    assert len(reg.prefilter(raw, variants[-1])) == 0
    assert len(reg.prefilter(raw, variants[-2])) == len(reg.distance_filter(raw, 0.1, 35.0))
