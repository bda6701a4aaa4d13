"""pcl::GeneralizedIterativeClosestPoint ("GICP" / "GICP_OMP", /root/reference/src/mrg_slam/registrations.cpp:93-116; SURVEY 8a
row G) — checks of the ORACLE restatement (oracle/gicp_pcl.cpp).  The product has no engine for this method yet; these tests
pin what can be pinned without the upstream binaries: the functor's analytic gradient against finite differences, applyState's
Euler convention, the covariance regularisation, the line search's acceptance conditions and known-transform recovery."""
import numpy as np
import pytest
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation as Rot

from tests import oraclelib as O


def _state_matrix(x):
    T = np.eye(4)
    T[:3, :3] = (Rot.from_euler("z", x[5]) * Rot.from_euler("y", x[4]) * Rot.from_euler("x", x[3])).as_matrix()
    T[:3, 3] = x[:3]
    return T


def test_apply_state_is_zyx_euler_on_the_left():
    rng = np.random.default_rng(0)
    for _ in range(20):
        x = np.concatenate([rng.normal(size=3), rng.uniform(-1.2, 1.2, 3)])
        np.testing.assert_allclose(O.gicp_pcl_apply_state(x), _state_matrix(x), atol=2e-6)
        base = _state_matrix(np.concatenate([rng.normal(size=3), rng.uniform(-1, 1, 3)]))
        want = base.copy()
        want[:3, :3] = _state_matrix(x)[:3, :3] @ base[:3, :3]  # rotation applied on the left of the rotation block ...
        want[:3, 3] = base[:3, 3] + x[:3]                        # ... and the translation ADDED (not rotated), as upstream
        np.testing.assert_allclose(O.gicp_pcl_apply_state(x, base), want, atol=5e-6)


def _problem(small_pair, rng, n=600):
    a, b, gt = small_pair
    src, tgt = b[:n].copy(), a
    tree = cKDTree(tgt[:, :3])
    q = src[:, :3] @ gt[:3, :3].T + gt[:3, 3]
    _, nn = tree.query(q)
    idx_src = np.arange(n, dtype=np.int32)[::2]
    idx_tgt = nn[::2].astype(np.int32)
    M = np.zeros((n, 3, 3))
    for i in range(n):
        A = rng.normal(size=(3, 3))
        M[i] = A @ A.T + 0.5 * np.eye(3)  # any SPD Mahalanobis matrix
    return src, tgt, idx_src, idx_tgt, M


def test_functor_gradient_matches_finite_differences(small_pair):
    rng = np.random.default_rng(1)
    src, tgt, idx_src, idx_tgt, M = _problem(small_pair, rng)
    # double-precision model of the same cost (the oracle evaluates the point transform in float like upstream)
    def cost(x):
        T = _state_matrix(x)
        d = src[idx_src, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3] - tgt[idx_tgt, :3]
        return float(np.einsum("ni,nij,nj->", d, M[idx_src], d) / len(idx_src))

    for _ in range(5):
        x = np.concatenate([rng.normal(scale=0.3, size=3), rng.uniform(-0.3, 0.3, 3)])
        f, g = O.gicp_pcl_fdf(src, tgt, idx_src, idx_tgt, M, x)
        assert abs(f - cost(x)) <= 2e-4 * max(1.0, abs(f))  # float transform vs double model
        num = np.zeros(6)
        for k in range(6):
            e = np.zeros(6)
            e[k] = 1e-6
            num[k] = (cost(x + e) - cost(x - e)) / 2e-6
        np.testing.assert_allclose(g, num, rtol=2e-3, atol=2e-3 * np.abs(num).max())


def test_covariances_are_plane_regularised(small_pair):
    a, _, _ = small_pair
    C = O.gicp_pcl_covariances(a[:800], k=20)
    w = np.linalg.eigvalsh(C)
    np.testing.assert_allclose(w, np.tile([1e-3, 1.0, 1.0], (len(C), 1)), atol=1e-9)
    # the small direction is the normal of the local plane: the smallest-variance direction of the 20 neighbours
    pts = a[:800, :3].astype(np.float64)
    tree = cKDTree(pts)
    _, nn = tree.query(pts[:50], k=20)
    for i in range(50):
        nb = pts[nn[i]]
        cov = np.cov(nb.T, bias=True)
        wv, V = np.linalg.eigh(cov)
        if wv[1] - wv[0] < 1e-3 * max(wv[2], 1e-12):  # ambiguous normal
            continue
        n_c = np.linalg.eigh(C[i])[1][:, 0]
        assert abs(abs(n_c @ V[:, 0]) - 1.0) < 1e-3


def test_bfgs_descends_and_line_search_satisfies_wolfe(small_pair):
    rng = np.random.default_rng(2)
    src, _, idx_src, _, M = _problem(small_pair, rng)
    # exact correspondences: target point i = T_true * source point i, so the minimum is x_true with f = 0
    x_true = np.array([0.4, -0.1, 0.05, 0.02, -0.03, 0.08])
    Tt = _state_matrix(x_true)
    tgt = src.copy()
    tgt[:, :3] = (src[:, :3].astype(np.float64) @ Tt[:3, :3].T + Tt[:3, 3]).astype(np.float32)
    idx_tgt = idx_src.copy()
    x0 = x_true + np.array([0.3, -0.2, 0.1, 0.05, -0.04, 0.06])
    f0, g0 = O.gicp_pcl_fdf(src, tgt, idx_src, idx_tgt, M, x0)
    # one step: strong Wolfe conditions with rho = sigma = 0.01 along the steepest-descent direction
    x1, inner, status, f1, evals = O.gicp_pcl_bfgs(src, tgt, idx_src, idx_tgt, M, x0, max_inner=1, gradient_tol=0.0)
    assert inner == 1 and status in (0, 1)
    p = -g0 / np.linalg.norm(g0)
    alpha = float((x1 - x0) @ p)
    np.testing.assert_allclose(x1 - x0, alpha * p, atol=1e-9)
    _, g1 = O.gicp_pcl_fdf(src, tgt, idx_src, idx_tgt, M, x1)
    assert alpha > 0 and f1 <= f0 + 0.01 * alpha * float(g0 @ p) + 1e-12
    assert abs(float(g1 @ p)) <= 0.01 * abs(float(g0 @ p)) * (1 + 1e-6) + 1e-9
    # twenty steps: monotone decrease, close to the transform the correspondences were made with
    fs = [f0]
    x = x0
    for it in range(1, 21):
        x, inner, status, f, _ = O.gicp_pcl_bfgs(src, tgt, idx_src, idx_tgt, M, x0, max_inner=it, gradient_tol=0.0)
        fs.append(f)
    assert all(b <= a + 1e-12 for a, b in zip(fs, fs[1:]))
    assert fs[-1] < 1e-3 * f0
    assert np.linalg.norm(x[:3] - x_true[:3]) < 0.02 and np.linalg.norm(x[3:] - x_true[3:]) < 5e-3


@pytest.mark.parametrize("offset", [(0.2, -0.15, 0.05, 0.03), (-0.3, 0.2, 0.0, -0.05)])
def test_align_recovers_a_known_transform(small_pair, offset):
    a, b, gt = small_pair
    # source = target moved by a known rigid transform: the answer is exact up to the convergence thresholds
    T = np.eye(4)
    T[:3, :3] = Rot.from_euler("z", offset[3]).as_matrix()
    T[:3, 3] = offset[:3]
    src = a.copy()
    src[:, :3] = (a[:, :3] - T[:3, 3]) @ T[:3, :3]  # p_src = T^-1 p_tgt
    r = O.gicp_pcl_align(a, src.astype(np.float32), np.eye(4))
    assert r.converged and 1 <= r.iterations <= 64
    got = O.from_colmajor(list(r.T))
    d = np.linalg.inv(T) @ got
    assert np.linalg.norm(d[:3, 3]) < 0.02 and np.arccos(np.clip((np.trace(d[:3, :3]) - 1) / 2, -1, 1)) < 5e-3
    # and from a guess: final = transformation * guess
    g = T.copy()
    g[0, 3] += 0.1
    r2 = O.gicp_pcl_align(a, src.astype(np.float32), g)
    d2 = np.linalg.inv(T) @ O.from_colmajor(list(r2.T))
    assert r2.converged and np.linalg.norm(d2[:3, 3]) < 0.02
    # consecutive real scans: agrees with the FAST_GICP oracle of the same pair within a few centimetres
    r3 = O.gicp_pcl_align(a, b, np.eye(4))
    fg = O.Registration(O.default_params(O.FAST_GICP))
    fg.setInputTarget(a); fg.setInputSource(b)
    fg.align(np.eye(4))
    d3 = np.linalg.inv(fg.getFinalTransformation()) @ O.from_colmajor(list(r3.T))
    assert r3.converged and np.linalg.norm(d3[:3, 3]) < 0.05


def test_result_is_reproducible_only_to_about_a_millimetre(small_pair):
    """Why the device engine is compared with this oracle at 2e-3 m rather than the north star's 1e-4 m: even with a tight outer
    stopping rule the answer moves by a fraction of a millimetre when the initial guess moves by 1e-7 m — the inner BFGS stops at
    |g| < 1e-2 and the outer rule stops wherever an iteration happens to land ("1e-4 parity fragile")."""
    a, b, gt = small_pair
    kw = dict(transformation_epsilon=2e-5, rotation_epsilon=2e-6, maximum_iterations=40)
    guess = np.eye(4)
    guess[0, 3], guess[1, 3] = gt[0, 3], gt[1, 3]
    base = O.from_colmajor(list(O.gicp_pcl_align(a, b, guess, O.gicp_pcl_params(**kw)).T))
    worst = 0.0
    for eps in (1e-7, 1e-6, 1e-5):
        g2 = guess.copy()
        g2[0, 3] += eps
        T = O.from_colmajor(list(O.gicp_pcl_align(a, b, g2, O.gicp_pcl_params(**kw)).T))
        d = np.linalg.inv(base) @ T
        worst = max(worst, float(np.linalg.norm(d[:3, 3])))
    assert 1e-5 < worst < 2e-3, worst
