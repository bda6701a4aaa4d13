"""Oracle self-checks (CPU): independent cross-checks against scipy / numpy and analytic known answers.

The reference has no tests and none of its numeric dependencies are vendored: parity is
UNPINNED against the real pclomp / fast_gicp / PCL binaries.  What can be pinned here is that the oracle's
building blocks (exact kNN, radius counts, tiny linear algebra, voxel keying) agree with independent
implementations, and that the registration restatements recover known transforms.
"""
import numpy as np
import pytest
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation

from tests import oraclelib as O
from tests.conftest import pose_error


def _dist2_flann(q, p):
    d = (q[:3] - p[:3]).astype(np.float32)
    r = np.float32(d[0] * d[0])
    r = np.float32(r + np.float32(d[1] * d[1]))
    return np.float32(r + np.float32(d[2] * d[2]))


# ------------------------------------------------------------------------------ search structures
def test_kdtree_knn_matches_scipy(small_pair):
    a, _, _ = small_pair
    rng = np.random.default_rng(1)
    q = np.concatenate([a[rng.integers(0, len(a), 300)], (rng.normal(size=(100, 4)) * 15).astype(np.float32)])
    idx, d2 = O.knn(a, q, 20)
    tree = cKDTree(a[:, :3].astype(np.float64))
    sd, si = tree.query(q[:, :3].astype(np.float64), k=20)
    # identical neighbour sets (float32 distances can reorder near-ties, so compare as sets) ...
    assert all(set(idx[i]) == set(si[i]) for i in range(len(q)))
    # ... ascending order and the exact FLANN float association for the distances
    assert np.all(np.diff(d2, axis=1) >= 0)
    for i in (0, 7, 311):
        for j in range(20):
            assert d2[i, j] == _dist2_flann(q[i], a[idx[i, j]])
    np.testing.assert_allclose(np.sqrt(d2), sd, rtol=1e-5, atol=1e-6)


def test_radius_outlier_matches_scipy(small_pair):
    a, _, _ = small_pair
    keep = O.radius_outlier(a, 0.8, 3)
    tree = cKDTree(a[:, :3].astype(np.float64))
    counts = np.array([len(x) for x in tree.query_ball_point(a[:, :3].astype(np.float64), 0.8)])
    # no point may sit on the decision boundary in float32 for this check to be meaningful
    assert np.array_equal(keep, counts > 3)
    assert 0 < keep.sum() < len(a)


def test_statistical_outlier_matches_scipy(small_pair):
    a, _, _ = small_pair
    keep, dist, thr = O.statistical_outlier(a, 10, 1.0)
    tree = cKDTree(a[:, :3].astype(np.float64))
    sd, _ = tree.query(a[:, :3].astype(np.float64), k=11)
    ref = sd[:, 1:].mean(axis=1)
    np.testing.assert_allclose(dist, ref, rtol=2e-6)
    ref_thr = ref.mean() + 1.0 * ref.std(ddof=1)
    assert abs(thr - ref_thr) < 1e-6 * ref_thr
    margin = np.abs(ref - ref_thr) > 1e-5
    assert np.array_equal(keep[margin], (ref <= ref_thr)[margin])


def test_fitness_score_matches_scipy(small_pair):
    a, b, gt = small_pair
    f, nr = O.fitness_score(a, b, gt)
    tb = O.transform_cloud(b, gt)
    d, _ = cKDTree(a[:, :3].astype(np.float64)).query(tb[:, :3].astype(np.float64))
    assert nr == len(b)
    assert abs(f - np.mean(d ** 2)) < 1e-5 * f
    # max_range is compared against the SQUARED distance (information_matrix_calculator.cpp:70)
    f2, nr2 = O.fitness_score(a, b, gt, max_range=0.25)
    sel = d ** 2 <= 0.25
    assert abs(nr2 - sel.sum()) <= 2
    assert abs(f2 - np.mean(d[sel] ** 2)) < 1e-3 * f2
    # nothing in range -> DBL_MAX
    far = b.copy(); far[:, :3] += 1000.0
    assert O.fitness_score(a, far, np.eye(4), max_range=1.0)[0] == np.finfo(np.float64).max


# ------------------------------------------------------------------------------ tiny linear algebra
def test_linalg_against_numpy():
    rng = np.random.default_rng(2)
    for _ in range(20):
        M = rng.normal(size=(6, 6))
        A = M @ M.T + 1e-3 * np.eye(6)
        rhs = rng.normal(size=6)
        np.testing.assert_allclose(O.test_ldlt6_solve(A, rhs), np.linalg.solve(A, rhs), rtol=1e-9, atol=1e-12)
        G = rng.normal(size=(6, 6))
        np.testing.assert_allclose(O.test_svd6_solve(G, rhs), np.linalg.solve(G, rhs), rtol=1e-8, atol=1e-10)
        S = rng.normal(size=(3, 3)); S = S @ S.T
        ev, V = O.test_sym3_eigen(S)
        np.testing.assert_allclose(ev, np.linalg.eigvalsh(S), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(V @ np.diag(ev) @ V.T, S, atol=1e-12)
        B4 = rng.normal(size=(4, 4))
        np.testing.assert_allclose(O.test_m4_inverse(B4), np.linalg.inv(B4), rtol=1e-9, atol=1e-10)
    # rank-deficient system: pseudo-inverse solution, as Eigen's JacobiSVD::solve gives
    G = np.diag([3.0, 2.0, 1.0, 0.5, 0.0, 0.0]); rhs = np.arange(1.0, 7.0)
    np.testing.assert_allclose(O.test_svd6_solve(G, rhs), np.linalg.pinv(G) @ rhs, atol=1e-12)


def test_so3_exp_and_euler():
    rng = np.random.default_rng(3)
    for w in list(rng.normal(size=(10, 3)) * 0.5) + [np.array([1e-7, -2e-7, 3e-7]), np.zeros(3)]:
        np.testing.assert_allclose(O.test_so3_exp(w), Rotation.from_rotvec(w).as_matrix(), atol=1e-12)
    # eulerAngles(0,1,2) followed by the Rx*Ry*Rz composition must reproduce the rotation, on both Euler branches
    for ang in ([0.02, -0.01, 0.3], [-0.02, 0.01, -0.3], [0.0, 0.0, 0.0], [1.0, -0.7, 2.0]):
        R = Rotation.from_euler("XYZ", ang).as_matrix()  # intrinsic = Rx*Ry*Rz
        e = O.test_euler_angles_012(R)
        M = O.test_ndt_matrix_from_p(np.concatenate([[1.0, 2.0, 3.0], e.astype(np.float64)]))
        np.testing.assert_allclose(M[:, :3], R, atol=5e-6)
        np.testing.assert_allclose(M[:, 3], [1, 2, 3])
        assert 0.0 <= e[0] <= np.pi + 1e-6  # Eigen >= 3.3 range of the first angle


def test_more_thuente_trial_values():
    # case 1 (f_t > f_l): quadratic f(a) = (a-1)^2, a_l=0, a_t=3 -> both interpolants find the minimiser 1
    f = lambda a: (a - 1.0) ** 2
    g = lambda a: 2 * (a - 1.0)
    v = O.test_mt_trial_value([0.0, f(0), g(0), 0.0, f(0), g(0), 3.0, f(3), g(3)])
    assert abs(v - 1.0) < 1e-12
    # case 2 (f_t <= f_l, derivatives of opposite sign): secant of the derivative hits the minimiser exactly
    v = O.test_mt_trial_value([0.0, f(0), g(0), 0.0, f(0), g(0), 1.5, f(1.5), g(1.5)])
    assert abs(v - 1.0) < 1e-12


# ------------------------------------------------------------------------------ filters
def test_distance_filter_literal():
    pts = np.array([[0.05, 0, 0, 1], [0.2, 0, 0, 2], [30, 10, 0, 3], [34, 9, 0, 4], [np.nan, 0, 0, 5], [0, 0, 35.0, 6]], dtype=np.float32)
    out = O.distance_filter(pts, 0.1, 35.0)
    assert out[:, 3].tolist() == [2.0, 3.0]  # strict bounds on both sides; NaN dropped; order preserved


def test_voxelgrid_semantics():
    rng = np.random.default_rng(4)
    pts = np.concatenate([rng.uniform(-5, 5, size=(4000, 3)), rng.uniform(0, 1, size=(4000, 1))], axis=1).astype(np.float32)
    out, ovf, vidx = O.voxelgrid(pts, 0.5, 1, want_index=True)
    assert not ovf
    assert np.all(np.diff(vidx) > 0)  # one point per voxel, ascending dense voxel index
    # independent recomputation of the keying and of the float32 sequential centroid
    inv = np.float32(1.0) / np.float32(0.5)
    mn = pts[:, :3].min(0); mx = pts[:, :3].max(0)
    min_b = np.floor(mn * inv).astype(np.int32); max_b = np.floor(mx * inv).astype(np.int32)
    div = max_b - min_b + 1
    ijk = (np.floor(pts[:, :3] * inv) - min_b.astype(np.float32)).astype(np.int32)
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    assert np.array_equal(np.unique(key), vidx)
    for v in (0, len(vidx) // 2, len(vidx) - 1):
        members = np.where(key == vidx[v])[0]
        s = np.zeros(4, dtype=np.float32)
        for m in members:
            s = (s + pts[m]).astype(np.float32)
        assert np.array_equal(out[v], s / np.float32(len(members)))
    # min_points_per_voxel drops sparse voxels
    out3, _ = O.voxelgrid(pts, 0.5, 3)
    cnt = np.bincount(np.searchsorted(vidx, key))
    assert len(out3) == (cnt >= 3).sum()
    # INT32 overflow branch: the input comes back unchanged
    big = pts.copy(); big[0, :3] = [1e4, 1e4, 1e4]
    o2, ovf2 = O.voxelgrid(big, 0.001, 1)
    assert ovf2 and np.array_equal(o2, big)
    # empty input
    e, eo = O.voxelgrid(np.zeros((0, 4), np.float32), 0.5, 1)
    assert len(e) == 0 and not eo


# ------------------------------------------------------------------------------ registration building blocks
def test_covariances_against_numpy(small_pair):
    a, _, _ = small_pair
    cov, knn = O.knn_covariances(a, 20, want_idx=True)
    for i in (0, 100, len(a) - 1):
        nb = a[knn[i], :3].astype(np.float64)
        C = np.cov(nb.T, bias=True)
        U, s, Vt = np.linalg.svd(C)
        ref = U @ np.diag([1.0, 1.0, 1e-3]) @ Vt
        got = np.array([[cov[i, 0], cov[i, 1], cov[i, 2]], [cov[i, 1], cov[i, 3], cov[i, 4]], [cov[i, 2], cov[i, 4], cov[i, 5]]])
        np.testing.assert_allclose(got, ref, atol=1e-9)


def test_vgicp_voxelmap_keying(small_pair):
    a, _, _ = small_pair
    cov = O.knn_covariances(a, 20)
    coords, npts, mean, vcov = O.vgicp_voxelmap(a, cov, 1.0)
    ref = np.floor(a[:, :3].astype(np.float64) / 1.0 - 0.5).astype(np.int32)  # fast_gicp voxel_coord
    uniq, inv, cnt = np.unique(ref, axis=0, return_inverse=True, return_counts=True)
    assert np.array_equal(coords, uniq) and np.array_equal(npts, cnt)
    v = 5
    members = np.where(inv.ravel() == v)[0]
    np.testing.assert_allclose(mean[v], a[members, :3].astype(np.float64).mean(0), rtol=1e-13)
    np.testing.assert_allclose(vcov[v], cov[members].mean(0), rtol=1e-12)


def test_linearize_matches_finite_differences(small_pair):
    """H, b of the oracle's VGICP cost are the Gauss-Newton terms of err(T): b is half its gradient w.r.t. a left
    perturbation delta = [so3_exp(w) | t] (with the correspondences / Mahalanobis matrices held fixed)."""
    a, b, gt = small_pair
    r = O.Registration(O.default_params(O.FAST_VGICP))
    r.setInputTarget(a); r.setInputSource(b)
    T0 = gt.copy(); T0[0, 3] -= 0.05
    err0, H, bb, corr, valid = r.linearize(T0)
    assert valid.sum() > 0.5 * len(b)
    assert np.allclose(H, H.T, rtol=1e-10)
    assert np.all(np.linalg.eigvalsh(H) > 0)
    assert abs(r.compute_error(T0) - err0) < 1e-9 * err0
    eps = 1e-6
    grad = np.zeros(6)
    for k in range(6):
        d = np.zeros(6); d[k] = eps
        D = np.eye(4); D[:3, :3] = Rotation.from_rotvec(d[:3]).as_matrix(); D[:3, 3] = d[3:]
        grad[k] = (r.compute_error(D @ T0) - r.compute_error(np.linalg.inv(D) @ T0)) / (2 * eps)
    np.testing.assert_allclose(grad, 2 * bb, rtol=2e-4, atol=1e-3 * np.abs(bb).max())


def test_small_gicp_linearisation_is_the_gauss_newton_model(small_pair):
    """small_gicp's factor: e(T exp(d)) with the linearisation's correspondences and Mahalanobis matrices must follow
    e0 + b.d + d.H.d/2 (right-multiplied se3 update; pins the Jacobian [R skew(p) | -R] and its signs)."""
    a, b, gt = small_pair
    r = O.Registration(O.default_params(O.SMALL_GICP))
    r.setInputTarget(a); r.setInputSource(b)
    T0 = gt.copy(); T0[0, 3] += 0.05
    e0, H, bb, corr, _ = r.linearize(T0)
    assert (corr >= 0).sum() > 0.5 * len(b) and np.allclose(H, H.T)
    assert abs(r.compute_error(T0) - e0) < 1e-12 * e0
    rng = np.random.default_rng(3)
    for _ in range(4):
        d = rng.normal(size=6) * 1e-3
        D = np.eye(4); D[:3, :3] = Rotation.from_rotvec(d[:3]).as_matrix(); D[:3, 3] = d[3:]  # se3_exp to first order in |d|
        pred = e0 + bb @ d + 0.5 * d @ H @ d
        got = r.compute_error(T0 @ D)
        assert abs(got - pred) <= 2e-3 * abs(bb @ d) + 1e-9 * e0, (got, pred)
    # and the solver's descent direction reduces the error
    step = np.linalg.solve(H + 1e-3 * np.eye(6), -bb)
    D = np.eye(4); D[:3, :3] = Rotation.from_rotvec(step[:3]).as_matrix(); D[:3, 3] = step[3:]
    assert r.compute_error(T0 @ D) < e0


@pytest.mark.parametrize("method", [O.FAST_GICP, O.FAST_VGICP, O.NDT_OMP, O.SMALL_GICP])
def test_identity_on_identical_clouds(small_pair, method):
    a, _, _ = small_pair
    r = O.Registration(O.default_params(method))
    r.setInputTarget(a); r.setInputSource(a)
    res = r.align(np.eye(4))
    assert res.converged
    te, re = pose_error(np.eye(4), r.getFinalTransformation())
    # NDT at eps=0.1 takes a clamped >= 0.05 step even from the optimum; GICP (point-to-point
    # correspondences) stays put exactly; VGICP's voxel means pull it a fraction of a millimetre
    t_tol = {O.NDT_OMP: 0.11, O.FAST_VGICP: 5e-3, O.FAST_GICP: 1e-6, O.SMALL_GICP: 1e-6}[method]
    r_tol = {O.NDT_OMP: 0.05, O.FAST_VGICP: 1e-3, O.FAST_GICP: 1e-6, O.SMALL_GICP: 1e-6}[method]
    assert te < t_tol and re < r_tol
    assert r.getFitnessScore() < {O.NDT_OMP: 0.02, O.FAST_VGICP: 1e-4, O.FAST_GICP: 1e-10, O.SMALL_GICP: 1e-10}[method]


@pytest.mark.parametrize("method,tol", [(O.FAST_GICP, 0.02), (O.FAST_VGICP, 0.02), (O.NDT_OMP, 0.03), (O.SMALL_GICP, 0.02)])
def test_recovers_known_transform(small_pair, method, tol):
    """Source = target moved by a known SE(3): the alignment must find its inverse (tight epsilon)."""
    a, _, _ = small_pair
    Tk = np.eye(4)
    Tk[:3, :3] = Rotation.from_euler("xyz", [0.01, -0.015, 0.04]).as_matrix()
    Tk[:3, 3] = [0.25, -0.15, 0.05]
    src = O.transform_cloud(a, np.linalg.inv(Tk))
    # the ~3k-point test cloud is sparse (0.4 m voxels): NDT needs 2 m leaves to have enough >=6-point cells
    extra = dict(transformation_epsilon=0.01, resolution=2.0) if method == O.NDT_OMP else dict(transformation_epsilon=1e-3)
    r = O.Registration(O.default_params(method, **extra))
    r.setInputTarget(a); r.setInputSource(src)
    res = r.align(np.eye(4))
    assert res.converged
    te, re = pose_error(Tk, r.getFinalTransformation())
    assert te < tol and re < 0.01, (te, re)


def test_lm_reports_iterations_and_convergence(vlp16_pair):
    a, b, gt = vlp16_pair
    r = O.Registration(O.default_params(O.FAST_VGICP))
    r.setInputTarget(a); r.setInputSource(b)
    g = gt.copy(); g[0, 3] -= 0.3
    res = r.align(g)
    assert res.converged and 0 <= res.iterations < 64 and res.lm_evals >= 2
    te, _ = pose_error(gt, r.getFinalTransformation())
    assert te < 0.05
    # maximum_iterations = 1 with a far guess: loop ends unconverged, hasConverged() false (callers drop the frame)
    r2 = O.Registration(O.default_params(O.FAST_VGICP, maximum_iterations=1, transformation_epsilon=1e-6, rotation_epsilon=1e-9))
    r2.setInputTarget(a); r2.setInputSource(b)
    assert not r2.align(g).converged


def test_ndt_kdtree_neighbourhood_is_a_radius_search_over_float_centroids(vlp16_pair):
    """pclomp KDTREE / pcl::NormalDistributionsTransform (registrations.cpp:121-141): the neighbourhood of a point is the set of
    leaves (>= 6 points) whose FLOAT centroid — sequential float sum in point order / float count — lies closer than the
    resolution (FLANN float distance, strict).  Checked against scipy's kd-tree over independently computed centroids."""
    from scipy.spatial import cKDTree
    a, b, _ = vlp16_pair
    res = 1.0
    o = O.Registration(O.default_params(O.NDT_OMP, resolution=res, neighbor_search=O.KDTREE))
    o.setInputTarget(a); o.setInputSource(b)
    _, _, _, hits = o.ndt_derivatives(np.zeros(6))  # identity: the transformed cloud is the source itself
    inv = np.float32(1.0) / np.float32(res)
    key = np.floor(a[:, :3] * inv).astype(np.int64)
    order = np.lexsort((np.arange(len(a)), key[:, 0], key[:, 1], key[:, 2]))  # voxels, point order inside a voxel
    ks = key[order]
    starts = np.flatnonzero(np.r_[True, (ks[1:] != ks[:-1]).any(axis=1)])
    ends = np.r_[starts[1:], len(a)]
    cen = []
    for s, e in zip(starts, ends):
        if e - s >= 6:
            pts = a[order[s:e], :3]
            cen.append(np.cumsum(pts, axis=0, dtype=np.float32)[-1] / np.float32(e - s))  # sequential float32 sums
    cen = np.array(cen, dtype=np.float32)
    tree = cKDTree(cen.astype(np.float64))
    want = np.zeros(len(b), dtype=np.int32)
    r2 = np.float32(np.float64(np.float32(res)) ** 2)
    for i, cand in enumerate(tree.query_ball_point(b[:, :3].astype(np.float64), res * 1.001)):
        if cand:
            d = b[i, :3] - cen[cand]
            d2 = (d[:, 0] * d[:, 0]).astype(np.float32)
            d2 = (d2 + (d[:, 1] * d[:, 1]).astype(np.float32)).astype(np.float32)
            d2 = (d2 + (d[:, 2] * d[:, 2]).astype(np.float32)).astype(np.float32)
            want[i] = int((d2 < r2).sum())
    # leaves whose covariance turned out unusable are dropped by the oracle (nr_points = -1): allow for those few
    assert (hits <= want).all() and (hits == want).mean() > 0.98 and hits.max() >= 3
    d7 = O.Registration(O.default_params(O.NDT_OMP, resolution=res))
    d7.setInputTarget(a); d7.setInputSource(b)
    _, _, _, hits7 = d7.ndt_derivatives(np.zeros(6))
    assert not np.array_equal(hits, hits7)  # a different neighbourhood from DIRECT7's face neighbours
