"""ctypes binding of oracle/liboracle.so — the CPU restatement used as the parity checker.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")

NDT_OMP, FAST_GICP, FAST_VGICP, SMALL_GICP = 0, 1, 2, 3
DIRECT1, DIRECT7, DIRECT27, KDTREE = 0, 1, 2, 3
# orc_set_variant ids (oracle/oracle.h): the alternative reading of every version-dependent detail of SURVEY Appendix A
VARIANTS = {"vgicp_coord_no_half": 0, "ndt_angle_eps_1e5": 1, "ndt_inner_double": 2, "mt_clamp_max_first": 3, "ndt_cov_newer_pcl": 4,
            "ndt_lookup_mul": 5, "voxelgrid_descending": 6, "radius_nonstrict": 7, "transform_left_to_right": 8, "norm_left_to_right": 9,
            "euler_no_fixup": 10}


class variant:
    """with variant("ndt_inner_double"): ...  — the oracle follows the alternative upstream reading inside the block."""

    def __init__(self, name, value=1):
        self.id, self.value = VARIANTS[name], value

    def __enter__(self):
        lib().orc_set_variant(self.id, self.value)

    def __exit__(self, *a):
        lib().orc_set_variant(self.id, 0)


class Params(ctypes.Structure):
    _fields_ = [
        ("method", ctypes.c_int),
        ("num_threads", ctypes.c_int),
        ("transformation_epsilon", ctypes.c_double),
        ("maximum_iterations", ctypes.c_int),
        ("max_correspondence_distance", ctypes.c_double),
        ("correspondence_randomness", ctypes.c_int),
        ("resolution", ctypes.c_double),
        ("neighbor_search", ctypes.c_int),
        ("rotation_epsilon", ctypes.c_double),
        ("lm_max_iterations", ctypes.c_int),
        ("lm_init_lambda_factor", ctypes.c_double),
        ("ndt_step_size", ctypes.c_double),
        ("ndt_outlier_ratio", ctypes.c_double),
    ]


class Result(ctypes.Structure):
    _fields_ = [
        ("T", ctypes.c_float * 16),
        ("converged", ctypes.c_int),
        ("iterations", ctypes.c_int),
        ("error", ctypes.c_double),
        ("lm_evals", ctypes.c_int),
    ]


class GicpPclParams(ctypes.Structure):
    _fields_ = [
        ("transformation_epsilon", ctypes.c_double),
        ("maximum_iterations", ctypes.c_int),
        ("use_reciprocal_correspondences", ctypes.c_int),
        ("max_correspondence_distance", ctypes.c_double),
        ("correspondence_randomness", ctypes.c_int),
        ("max_optimizer_iterations", ctypes.c_int),
        ("rotation_epsilon", ctypes.c_double),
        ("gicp_epsilon", ctypes.c_double),
    ]


_lib = None


def build():
    subprocess.run(["make", "-C", _ORACLE_DIR], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        L.orc_default_params.argtypes = [ci, ctypes.POINTER(Params)]
        L.orc_reg_create.restype = vp
        L.orc_reg_create.argtypes = [ctypes.POINTER(Params)]
        L.orc_reg_destroy.argtypes = [vp]
        L.orc_reg_set_target.argtypes = [vp, vp, ci]
        L.orc_reg_set_source.argtypes = [vp, vp, ci]
        L.orc_reg_align.argtypes = [vp, vp, ctypes.POINTER(Result)]
        L.orc_reg_fitness.restype = cd
        L.orc_reg_fitness.argtypes = [vp, cd]
        L.orc_knn_covariances.argtypes = [vp, ci, ci, vp, vp]
        L.orc_vgicp_voxelmap.restype = ci
        L.orc_vgicp_voxelmap.argtypes = [vp, ci, vp, cd, vp, vp, vp, vp]
        L.orc_reg_linearize.restype = cd
        L.orc_reg_linearize.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_reg_compute_error.restype = cd
        L.orc_reg_compute_error.argtypes = [vp, vp]
        L.orc_ndt_grid.restype = ci
        L.orc_ndt_grid.argtypes = [vp, ci, cd, vp, vp, vp, vp, vp, vp]
        L.orc_reg_ndt_derivatives.restype = cd
        L.orc_reg_ndt_derivatives.argtypes = [vp, vp, vp, vp, vp]
        L.orc_distance_filter.restype = ci
        L.orc_distance_filter.argtypes = [vp, ci, cd, cd, vp]
        L.orc_voxelgrid.restype = ci
        L.orc_voxelgrid.argtypes = [vp, ci, ctypes.c_float, ci, vp, vp]
        L.orc_radius_outlier.restype = ci
        L.orc_radius_outlier.argtypes = [vp, ci, cd, ci, vp]
        L.orc_statistical_outlier.restype = ci
        L.orc_statistical_outlier.argtypes = [vp, ci, ci, cd, vp, vp, ctypes.POINTER(cd)]
        L.orc_transform_cloud.argtypes = [vp, ci, vp, vp]
        L.orc_fitness_score.restype = cd
        L.orc_fitness_score.argtypes = [vp, ci, vp, ci, vp, cd, ctypes.POINTER(ci)]
        L.orc_knn.argtypes = [vp, ci, vp, ci, ci, vp, vp]
        L.orc_map_cloud.restype = ci
        L.orc_map_cloud.argtypes = [vp, vp, vp, vp, ci, ctypes.c_float, ci, ctypes.c_float, ci, vp, vp]
        L.orc_gicp_pcl_default_params.argtypes = [ctypes.POINTER(GicpPclParams)]
        L.orc_gicp_pcl_align.restype = ci
        L.orc_gicp_pcl_align.argtypes = [vp, ci, vp, ci, ctypes.POINTER(GicpPclParams), vp, ctypes.POINTER(Result)]
        L.orc_gicp_pcl_apply_state.argtypes = [vp, vp]
        L.orc_gicp_pcl_fdf.restype = cd
        L.orc_gicp_pcl_fdf.argtypes = [vp, vp, vp, vp, ci, vp, ci, vp, vp]
        L.orc_gicp_pcl_bfgs.restype = ci
        L.orc_gicp_pcl_bfgs.argtypes = [vp, vp, vp, vp, ci, vp, ci, vp, ci, cd, ctypes.POINTER(ci), ctypes.POINTER(cd), ctypes.POINTER(ci)]
        L.orc_gicp_pcl_covariances.argtypes = [vp, ci, ci, cd, vp]
        L.orc_set_num_threads.argtypes = [ci]
        L.orc_get_max_threads.restype = ci
        _lib = L
    return _lib


def _pts(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4
    return a


def _p(a):
    return a.ctypes.data


def default_params(method, **overrides):
    p = Params()
    lib().orc_default_params(method, ctypes.byref(p))
    for k, v in overrides.items():
        assert hasattr(p, k), k
        setattr(p, k, v)
    return p


def colmajor(T):
    """4x4 (row-indexed numpy) -> 16 floats column-major (Eigen::Matrix4f storage)."""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T.reshape(16))


def from_colmajor(t16):
    return np.asarray(t16, dtype=np.float64).reshape(4, 4).T.copy()


class Registration:
    """Mirror of the pcl::Registration surface the reference's callers use."""

    def __init__(self, params):
        self.params = params
        self._h = lib().orc_reg_create(ctypes.byref(params))
        self._keep = []

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_reg_destroy(self._h)
            self._h = None

    def setInputTarget(self, cloud):
        c = _pts(cloud)
        lib().orc_reg_set_target(self._h, _p(c), len(c))
        self.nt = len(c)

    def setInputSource(self, cloud):
        c = _pts(cloud)
        lib().orc_reg_set_source(self._h, _p(c), len(c))
        self.ns = len(c)

    def align(self, guess=None):
        g = colmajor(np.eye(4) if guess is None else guess)
        r = Result()
        lib().orc_reg_align(self._h, _p(g), ctypes.byref(r))
        self.result = r
        return r

    def getFinalTransformation(self):
        return from_colmajor(list(self.result.T))

    def hasConverged(self):
        return bool(self.result.converged)

    def getFitnessScore(self, max_range=np.finfo(np.float64).max):
        return lib().orc_reg_fitness(self._h, max_range)

    # --- intermediates ---
    def linearize(self, T):
        T = np.ascontiguousarray(T, dtype=np.float64)
        H = np.zeros((6, 6)); b = np.zeros(6)
        if self.params.method == FAST_VGICP:
            corr = np.zeros((self.ns, 3), dtype=np.int32)
            valid = np.zeros(self.ns, dtype=np.uint8)
            err = lib().orc_reg_linearize(self._h, _p(T), _p(H), _p(b), _p(corr), _p(valid))
            return err, H, b, corr, valid.astype(bool)
        corr = np.zeros(self.ns, dtype=np.int32)
        err = lib().orc_reg_linearize(self._h, _p(T), _p(H), _p(b), _p(corr), None)
        return err, H, b, corr, corr >= 0

    def compute_error(self, T):
        T = np.ascontiguousarray(T, dtype=np.float64)
        return lib().orc_reg_compute_error(self._h, _p(T))

    def ndt_derivatives(self, p6):
        p6 = np.ascontiguousarray(p6, dtype=np.float64)
        g = np.zeros(6); H = np.zeros((6, 6)); hits = np.zeros(self.ns, dtype=np.int32)
        s = lib().orc_reg_ndt_derivatives(self._h, _p(p6), _p(g), _p(H), _p(hits))
        return s, g, H, hits


def knn_covariances(cloud, k=20, want_idx=False):
    c = _pts(cloud)
    cov = np.zeros((len(c), 6))
    idx = np.zeros((len(c), k), dtype=np.int32) if want_idx else None
    lib().orc_knn_covariances(_p(c), len(c), k, _p(cov), _p(idx) if want_idx else None)
    return (cov, idx) if want_idx else cov


def vgicp_voxelmap(cloud, cov6, resolution):
    c = _pts(cloud)
    cov6 = np.ascontiguousarray(cov6, dtype=np.float64)
    n = len(c)
    coords = np.zeros((n, 3), dtype=np.int32); npts = np.zeros(n, dtype=np.int32)
    mean = np.zeros((n, 3)); cov = np.zeros((n, 6))
    V = lib().orc_vgicp_voxelmap(_p(c), n, _p(cov6), resolution, _p(coords), _p(npts), _p(mean), _p(cov))
    return coords[:V].copy(), npts[:V].copy(), mean[:V].copy(), cov[:V].copy()


def ndt_grid(cloud, resolution):
    c = _pts(cloud)
    n = len(c)
    idx = np.zeros(n, dtype=np.int32); npts = np.zeros(n, dtype=np.int32)
    mean = np.zeros((n, 3)); icov = np.zeros((n, 9))
    min_b = np.zeros(3, dtype=np.int32); div_b = np.zeros(3, dtype=np.int32)
    V = lib().orc_ndt_grid(_p(c), n, resolution, _p(idx), _p(npts), _p(mean), _p(icov), _p(min_b), _p(div_b))
    return idx[:V].copy(), npts[:V].copy(), mean[:V].copy(), icov[:V].reshape(V, 3, 3).copy(), min_b, div_b


def distance_filter(cloud, near, far):
    c = _pts(cloud)
    out = np.empty_like(c)
    m = lib().orc_distance_filter(_p(c), len(c), near, far, _p(out))
    return out[:m].copy()


def voxelgrid(cloud, leaf, min_points=1, want_index=False):
    """Returns (points, overflow).  On the INT32 overflow branch PCL returns the input unchanged."""
    c = _pts(cloud)
    out = np.empty_like(c)
    vidx = np.zeros(len(c), dtype=np.int32)
    m = lib().orc_voxelgrid(_p(c), len(c), leaf, min_points, _p(out), _p(vidx))
    if m < 0:
        return (out.copy(), True, None) if want_index else (out.copy(), True)
    return (out[:m].copy(), False, vidx[:m].copy()) if want_index else (out[:m].copy(), False)


def radius_outlier(cloud, radius, min_neighbors):
    c = _pts(cloud)
    keep = np.zeros(len(c), dtype=np.uint8)
    lib().orc_radius_outlier(_p(c), len(c), radius, min_neighbors, _p(keep))
    return keep.astype(bool)


def statistical_outlier(cloud, mean_k, stddev_mul):
    c = _pts(cloud)
    keep = np.zeros(len(c), dtype=np.uint8)
    dist = np.zeros(len(c), dtype=np.float32)
    thr = ctypes.c_double()
    lib().orc_statistical_outlier(_p(c), len(c), mean_k, stddev_mul, _p(keep), _p(dist), ctypes.byref(thr))
    return keep.astype(bool), dist, thr.value


def transform_cloud(cloud, T):
    c = _pts(cloud)
    out = np.empty_like(c)
    g = colmajor(T)
    lib().orc_transform_cloud(_p(c), len(c), _p(g), _p(out))
    return out


def fitness_score(target, source, T, max_range=np.finfo(np.float64).max):
    t, s = _pts(target), _pts(source)
    g = colmajor(T)
    nr = ctypes.c_int()
    f = lib().orc_fitness_score(_p(t), len(t), _p(s), len(s), _p(g), max_range, ctypes.byref(nr))
    return f, nr.value


def map_cloud(clouds, poses, first_keyframe=None, resolution=0.05, min_points_per_voxel=1, distance_far_thresh=-1.0, skip_first_cloud=False):
    """MapCloudGenerator::generate.  poses: 4x4 float64 (world <- keyframe).  Returns (points, voxel keys) or None for nullptr.
    Points come in the oracle's unordered_map iteration order; sort by the keys to compare."""
    cs = [_pts(c) for c in clouds]
    n = len(cs)
    ptrs = (ctypes.c_void_p * max(n, 1))(*[c.ctypes.data for c in cs])
    ns = np.array([len(c) for c in cs], dtype=np.int32)
    P = np.ascontiguousarray(np.stack([np.asarray(p, dtype=np.float64).T.reshape(16) for p in poses])) if n else np.zeros((0, 16))
    fk = np.zeros(max(n, 1), dtype=np.uint8) if first_keyframe is None else np.asarray(first_keyframe, dtype=np.uint8)
    out = np.empty((max(int(ns.sum()), 1), 4), dtype=np.float32)
    keys = np.zeros((len(out), 3), dtype=np.int32)
    m = lib().orc_map_cloud(ptrs, _p(ns), _p(P), _p(fk), n, resolution, min_points_per_voxel, distance_far_thresh, int(skip_first_cloud), _p(out),
                            _p(keys))
    if m < 0:
        return None
    return out[:m].copy(), keys[:m].copy()


def knn(cloud, queries, k):
    c, q = _pts(cloud), _pts(queries)
    idx = np.zeros((len(q), k), dtype=np.int32); d2 = np.zeros((len(q), k), dtype=np.float32)
    lib().orc_knn(_p(c), len(c), _p(q), len(q), k, _p(idx), _p(d2))
    return idx, d2


def _call_test(name, *arrays):
    fn = getattr(lib(), name)
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p] * len(arrays)
    fn(*[a.ctypes.data for a in arrays])


def test_ldlt6_solve(A, rhs):
    A = np.ascontiguousarray(A, dtype=np.float64); rhs = np.ascontiguousarray(rhs, dtype=np.float64); x = np.zeros(6)
    _call_test("orc_test_ldlt6_solve", A, rhs, x)
    return x


def test_svd6_solve(A, rhs):
    A = np.ascontiguousarray(A, dtype=np.float64); rhs = np.ascontiguousarray(rhs, dtype=np.float64); x = np.zeros(6)
    _call_test("orc_test_svd6_solve", A, rhs, x)
    return x


def test_sym3_eigen(A):
    A = np.ascontiguousarray(A, dtype=np.float64); ev = np.zeros(3); V = np.zeros((3, 3))
    _call_test("orc_test_sym3_eigen", A, ev, V)
    return ev, V


def test_m4_inverse(A):
    A = np.ascontiguousarray(A, dtype=np.float64); inv = np.zeros((4, 4))
    _call_test("orc_test_m4_inverse", A, inv)
    return inv


def test_so3_exp(omega):
    w = np.ascontiguousarray(omega, dtype=np.float64); R = np.zeros((3, 3))
    _call_test("orc_test_so3_exp", w, R)
    return R


def test_euler_angles_012(R):
    R = np.ascontiguousarray(R, dtype=np.float32); res = np.zeros(3, dtype=np.float32)
    _call_test("orc_test_euler_angles_012", R, res)
    return res


def test_ndt_matrix_from_p(p6):
    p = np.ascontiguousarray(p6, dtype=np.float64); M = np.zeros((3, 4), dtype=np.float32)
    _call_test("orc_test_ndt_matrix_from_p", p, M)
    return M


def test_mt_trial_value(v9):
    v = np.ascontiguousarray(v9, dtype=np.float64)
    fn = lib().orc_test_mt_trial_value
    fn.restype = ctypes.c_double
    fn.argtypes = [ctypes.c_void_p]
    return fn(v.ctypes.data)


for _f in (test_ldlt6_solve, test_svd6_solve, test_sym3_eigen, test_m4_inverse, test_so3_exp, test_euler_angles_012,
           test_ndt_matrix_from_p, test_mt_trial_value):
    _f.__test__ = False  # helpers, not pytest tests


def set_num_threads(n):
    lib().orc_set_num_threads(n)


def max_threads():
    return lib().orc_get_max_threads()


# ---- pcl::GeneralizedIterativeClosestPoint ("GICP", registrations.cpp:93-103): oracle only, see oracle/gicp_pcl.cpp
def gicp_pcl_params(**overrides):
    p = GicpPclParams()
    lib().orc_gicp_pcl_default_params(ctypes.byref(p))
    for k, v in overrides.items():
        assert hasattr(p, k), k
        setattr(p, k, v)
    return p


def gicp_pcl_align(target, source, guess=None, params=None):
    t, s = _pts(target), _pts(source)
    g = colmajor(np.eye(4) if guess is None else guess)
    r = Result()
    p = params or gicp_pcl_params()
    rc = lib().orc_gicp_pcl_align(_p(t), len(t), _p(s), len(s), ctypes.byref(p), _p(g), ctypes.byref(r))
    assert rc == 0, rc
    return r


def gicp_pcl_apply_state(x6, base=None):
    T = colmajor(np.eye(4) if base is None else base).copy()
    x = np.ascontiguousarray(x6, dtype=np.float64)
    lib().orc_gicp_pcl_apply_state(_p(x), _p(T))
    return from_colmajor(T)


def gicp_pcl_fdf(src, tgt, idx_src, idx_tgt, mahalanobis, x6, want_grad=True):
    s, t = _pts(src), _pts(tgt)
    a, b = np.ascontiguousarray(idx_src, dtype=np.int32), np.ascontiguousarray(idx_tgt, dtype=np.int32)
    M = np.ascontiguousarray(mahalanobis, dtype=np.float64).reshape(len(s), 9)
    x = np.ascontiguousarray(x6, dtype=np.float64)
    g = np.zeros(6)
    f = lib().orc_gicp_pcl_fdf(_p(s), _p(t), _p(a), _p(b), len(a), _p(M), len(s), _p(x), _p(g) if want_grad else None)
    return f, g


def gicp_pcl_bfgs(src, tgt, idx_src, idx_tgt, mahalanobis, x6, max_inner=20, gradient_tol=1e-2):
    s, t = _pts(src), _pts(tgt)
    a, b = np.ascontiguousarray(idx_src, dtype=np.int32), np.ascontiguousarray(idx_tgt, dtype=np.int32)
    M = np.ascontiguousarray(mahalanobis, dtype=np.float64).reshape(len(s), 9)
    x = np.ascontiguousarray(x6, dtype=np.float64).copy()
    st, ev, f = ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
    inner = lib().orc_gicp_pcl_bfgs(_p(s), _p(t), _p(a), _p(b), len(a), _p(M), len(s), _p(x), max_inner, gradient_tol, ctypes.byref(st),
                                    ctypes.byref(f), ctypes.byref(ev))
    return x, inner, st.value, f.value, ev.value


def gicp_pcl_covariances(cloud, k=20, gicp_epsilon=1e-3):
    c = _pts(cloud)
    out = np.zeros((len(c), 9))
    lib().orc_gicp_pcl_covariances(_p(c), len(c), k, gicp_epsilon, _p(out))
    return out.reshape(len(c), 3, 3)
