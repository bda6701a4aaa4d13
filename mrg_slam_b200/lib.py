"""ctypes binding of libb2r.so (include/b2r.h) — the product's C ABI.

This is the binding the tests and bench.py go through; it deliberately has no CPU path: if the shared
library is missing, or no CUDA device is present, construction fails loudly.

The `Registration` class mirrors the pcl::Registration surface that the reference's factory
`select_registration_method()` hands out (/root/reference/src/mrg_slam/registrations.cpp:28-152) and its
callers use (apps/scan_matching_odometry_component.cpp:203-275, src/mrg_slam/loop_detector.cpp:104-144).
"""
import ctypes
import os

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2R_LIB_PATH") or os.path.join(_DIR, "libb2r.so")  # override: kernel experiments only

NDT_OMP, FAST_GICP, FAST_VGICP, SMALL_GICP, GICP_PCL = 0, 1, 2, 3, 4
DIRECT1, DIRECT7, DIRECT27, KDTREE = 0, 1, 2, 3
HOST, DEVICE = 0, 1
OK, ERR_INVALID_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_CAPACITY, ERR_STATE, ERR_COMM = range(7)
UNIQUE_ID_BYTES = 128
METHOD_BY_NAME = {"NDT_OMP": NDT_OMP, "FAST_GICP": FAST_GICP, "FAST_VGICP": FAST_VGICP, "SMALL_GICP": SMALL_GICP, "GICP": GICP_PCL,
                  "GICP_OMP": GICP_PCL}


class Config(ctypes.Structure):
    _fields_ = [
        ("method", ctypes.c_int),
        ("device", ctypes.c_int),
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
        ("nn_cell_size", ctypes.c_double),
        ("max_optimizer_iterations", ctypes.c_int),
        ("gicp_epsilon", ctypes.c_double),
    ]


class Result(ctypes.Structure):
    _fields_ = [
        ("T", ctypes.c_float * 16),
        ("converged", ctypes.c_int),
        ("iterations", ctypes.c_int),
        ("error", ctypes.c_double),
        ("evals", ctypes.c_int),
        ("reserved", ctypes.c_int),
        ("fitness", ctypes.c_double),
    ]


# numpy view of b2r_result (natural C alignment, matches ctypes' layout of Result)
RESULT_DTYPE = np.dtype({"names": ["T", "converged", "iterations", "error", "evals", "reserved", "fitness"],
                         "formats": [("<f4", 16), "<i4", "<i4", "<f8", "<i4", "<i4", "<f8"],
                         "offsets": [Result.T.offset, Result.converged.offset, Result.iterations.offset, Result.error.offset,
                                     Result.evals.offset, Result.reserved.offset, Result.fitness.offset],
                         "itemsize": ctypes.sizeof(Result)})


class PrefilterConfig(ctypes.Structure):
    _fields_ = [
        ("enable_distance_filter", ctypes.c_int),
        ("distance_near_thresh", ctypes.c_double),
        ("distance_far_thresh", ctypes.c_double),
        ("downsample_method", ctypes.c_int),
        ("downsample_resolution", ctypes.c_float),
        ("downsample_min_points_per_voxel", ctypes.c_int),
        ("outlier_removal_method", ctypes.c_int),
        ("statistical_mean_k", ctypes.c_int),
        ("statistical_stddev", ctypes.c_double),
        ("radius_radius", ctypes.c_double),
        ("radius_min_neighbors", ctypes.c_int),
    ]


# every symbol include/b2r.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "b2r_default_config", "b2r_create", "b2r_destroy", "b2r_last_error", "b2r_version",
    "b2r_cloud_create", "b2r_cloud_create_batch", "b2r_cloud_destroy", "b2r_cloud_destroy_batch", "b2r_cloud_size",
    "b2r_set_target", "b2r_set_source", "b2r_set_target_cloud", "b2r_set_source_cloud",
    "b2r_align", "b2r_fitness", "b2r_transform_source", "b2r_fitness_pair", "b2r_align_batch",
    "b2r_distance_filter", "b2r_voxelgrid", "b2r_radius_outlier", "b2r_statistical_outlier",
    "b2r_default_prefilter_config", "b2r_prefilter", "b2r_map_cloud", "b2r_remove_robot_points",
    "b2r_kernel_launches", "b2r_synchronize", "b2r_debug_knn_list_overflows", "b2r_debug_covariances", "b2r_debug_voxelmap",
    "b2r_debug_linearize", "b2r_debug_compute_error", "b2r_debug_ndt_grid", "b2r_debug_ndt_derivatives",
    "b2r_debug_knn", "b2r_last_timings", "b2r_event_record", "b2r_event_elapsed_ms", "b2r_profile_enable", "b2r_profile_read",
    "b2r_inlier_fraction", "b2r_nearest_neighbors", "b2r_comm_unique_id", "b2r_comm_init", "b2r_comm_init_host", "b2r_comm_destroy", "b2r_comm_rank",
    "b2r_comm_size", "b2r_comm_last_error", "b2r_comm_collectives", "b2r_partition_by_target", "b2r_align_batch_sharded",
    "b2r_gather_results", "b2r_select_best_candidates", "b2r_graph_launches",
]
ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)
PROFILE_KERNELS = {"knn_cov": 0, "lsq_eval": 1, "ndt_eval": 2, "grid_build": 3, "voxel_reduce": 4, "fitness": 5}

_lib = None


class B2RError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"b2r status {status}: {message}")
        self.status = status


def load():
    """Loads libb2r.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cd, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t
    L.b2r_version.restype = ctypes.c_char_p
    L.b2r_last_error.restype = ctypes.c_char_p
    L.b2r_last_error.argtypes = [vp]
    L.b2r_default_config.argtypes = [ci, ctypes.POINTER(Config)]
    L.b2r_create.argtypes = [ctypes.POINTER(Config), ctypes.POINTER(vp)]
    L.b2r_destroy.argtypes = [vp]
    L.b2r_destroy.restype = None
    L.b2r_cloud_create.argtypes = [vp, vp, sz, sz, ci, ctypes.POINTER(vp)]
    L.b2r_cloud_create_batch.argtypes = [vp, vp, vp, sz, sz, ci, vp]
    L.b2r_cloud_destroy.argtypes = [vp]
    L.b2r_cloud_destroy.restype = None
    L.b2r_cloud_destroy_batch.argtypes = [vp, sz]
    L.b2r_cloud_destroy_batch.restype = None
    L.b2r_cloud_size.argtypes = [vp]
    L.b2r_cloud_size.restype = sz
    for name in ("b2r_set_target", "b2r_set_source"):
        getattr(L, name).argtypes = [vp, vp, sz, sz, ci]
    for name in ("b2r_set_target_cloud", "b2r_set_source_cloud"):
        getattr(L, name).argtypes = [vp, vp]
    L.b2r_align.argtypes = [vp, vp, ctypes.POINTER(Result)]
    L.b2r_fitness.argtypes = [vp, cd, ctypes.POINTER(cd)]
    L.b2r_transform_source.argtypes = [vp, vp, sz, ci]
    L.b2r_fitness_pair.argtypes = [vp, vp, vp, vp, cd, ctypes.POINTER(cd)]
    L.b2r_align_batch.argtypes = [vp, vp, vp, vp, sz, ci, cd, vp]
    L.b2r_distance_filter.argtypes = [vp, vp, sz, sz, ci, cd, cd, vp, ctypes.POINTER(sz)]
    L.b2r_voxelgrid.argtypes = [vp, vp, sz, sz, ci, ctypes.c_float, ci, vp, ctypes.POINTER(sz), ctypes.POINTER(ci)]
    L.b2r_radius_outlier.argtypes = [vp, vp, sz, sz, ci, cd, ci, vp, ctypes.POINTER(sz)]
    L.b2r_statistical_outlier.argtypes = [vp, vp, sz, sz, ci, ci, cd, vp, ctypes.POINTER(sz)]
    L.b2r_default_prefilter_config.argtypes = [ctypes.POINTER(PrefilterConfig)]
    L.b2r_prefilter.argtypes = [vp, ctypes.POINTER(PrefilterConfig), vp, sz, sz, ci, vp, ctypes.POINTER(sz)]
    L.b2r_map_cloud.argtypes = [vp, vp, vp, vp, vp, sz, sz, ci, ctypes.c_float, ci, ctypes.c_float, ci, vp, ctypes.POINTER(sz),
                                ctypes.POINTER(ci)]
    L.b2r_remove_robot_points.argtypes = [vp, vp, sz, sz, ci, vp, sz, ctypes.c_float, vp, ctypes.POINTER(sz), vp, ctypes.POINTER(sz)]
    L.b2r_kernel_launches.argtypes = [vp]
    L.b2r_kernel_launches.restype = ctypes.c_uint64
    L.b2r_synchronize.argtypes = [vp]
    L.b2r_debug_knn_list_overflows.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64)]
    L.b2r_debug_covariances.argtypes = [vp, ci, vp, vp]
    L.b2r_debug_voxelmap.argtypes = [vp, vp, vp, vp, vp, ctypes.POINTER(sz)]
    L.b2r_debug_linearize.argtypes = [vp, vp, vp, vp, ctypes.POINTER(cd), vp, vp]
    L.b2r_debug_compute_error.argtypes = [vp, vp, vp, ctypes.POINTER(cd)]
    L.b2r_debug_ndt_grid.argtypes = [vp, vp, vp, vp, vp, vp, vp, ctypes.POINTER(sz)]
    L.b2r_debug_ndt_derivatives.argtypes = [vp, vp, ctypes.POINTER(cd), vp, vp, vp]
    L.b2r_debug_knn.argtypes = [vp, vp, vp, sz, ci, vp, vp]
    L.b2r_last_timings.argtypes = [vp, vp]
    L.b2r_event_record.argtypes = [vp, ci]
    L.b2r_event_elapsed_ms.argtypes = [vp, ci, ci, ctypes.POINTER(ctypes.c_float)]
    L.b2r_profile_enable.argtypes = [vp, ci]
    L.b2r_profile_read.argtypes = [vp, ci, ctypes.POINTER(cd), ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(cd)]
    L.b2r_inlier_fraction.argtypes = [vp, cd, ctypes.POINTER(cd), ctypes.POINTER(cd)]
    L.b2r_nearest_neighbors.argtypes = [vp, vp, vp, vp]
    L.b2r_graph_launches.argtypes = [vp]
    L.b2r_graph_launches.restype = ctypes.c_uint64
    L.b2r_comm_unique_id.argtypes = [vp]
    L.b2r_comm_init.argtypes = [vp, vp, ci, ci, ctypes.POINTER(vp)]
    L.b2r_comm_init_host.argtypes = [ALLGATHER_FN, vp, ci, ci, ctypes.POINTER(vp)]
    L.b2r_comm_destroy.argtypes = [vp]
    L.b2r_comm_destroy.restype = None
    L.b2r_comm_rank.argtypes = [vp]
    L.b2r_comm_size.argtypes = [vp]
    L.b2r_comm_last_error.argtypes = [vp]
    L.b2r_comm_last_error.restype = ctypes.c_char_p
    L.b2r_comm_collectives.argtypes = [vp]
    L.b2r_comm_collectives.restype = ctypes.c_uint64
    L.b2r_partition_by_target.argtypes = [vp, vp, sz, ci, vp]
    L.b2r_align_batch_sharded.argtypes = [vp, vp, vp, vp, vp, vp, vp, sz, ci, cd, vp]
    L.b2r_gather_results.argtypes = [vp, vp, vp, sz, vp, vp]
    L.b2r_select_best_candidates.argtypes = [vp, vp, sz, cd, vp, vp, ctypes.POINTER(sz)]
    _lib = L
    return L


def default_config(method, **overrides):
    if isinstance(method, str):
        method = METHOD_BY_NAME[method]
    cfg = Config()
    load().b2r_default_config(method, ctypes.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def colmajor(T):
    """4x4 numpy matrix -> 16 float32, column-major (Eigen::Matrix4f storage)."""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T.reshape(16))


def from_colmajor(t16):
    return np.asarray(t16, dtype=np.float64).reshape(4, 4).T.copy()


def _points(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] not in (4, 8):
        raise ValueError("points must be (n,4) packed or (n,8) pcl::PointXYZI float32")
    return a


class Cloud:
    """A device-resident cloud (b2r_cloud) with cached search structures."""

    def __init__(self, reg, points=None, device_ptr=None, host_ptr=None, n=None, stride=16):
        self._reg = reg
        self._lib = load()
        h = ctypes.c_void_p()
        if device_ptr is not None:
            st = self._lib.b2r_cloud_create(reg._h, ctypes.c_void_p(device_ptr), n, stride, DEVICE, ctypes.byref(h))
        elif host_ptr is not None:  # raw (e.g. pinned) host buffer
            st = self._lib.b2r_cloud_create(reg._h, ctypes.c_void_p(host_ptr), n, stride, HOST, ctypes.byref(h))
        else:
            a = _points(points)
            st = self._lib.b2r_cloud_create(reg._h, a.ctypes.data, len(a), a.shape[1] * 4, HOST, ctypes.byref(h))
        reg._check(st)
        self._h = h

    def __len__(self):
        return int(self._lib.b2r_cloud_size(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b2r_cloud_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def create_clouds(reg, pointers, sizes, memspace, stride=16):
    """b2r_cloud_create_batch over raw buffers (device or host pointers): one synchronisation for the whole batch."""
    n = len(pointers)
    P = (ctypes.c_void_p * n)(*pointers)
    N = (ctypes.c_size_t * n)(*sizes)
    H = (ctypes.c_void_p * n)()
    reg._check(load().b2r_cloud_create_batch(reg._h, P, N, n, stride, memspace, H))
    out = []
    for i in range(n):
        c = Cloud.__new__(Cloud)
        c._reg, c._lib, c._h = reg, load(), ctypes.c_void_p(H[i])
        out.append(c)
    return out


def partition_by_target(target_ids, world_size, weights=None):
    """b2r_partition_by_target: rank of every pair (numpy int32).  Pure host code in libb2r.so (runs without a GPU)."""
    t = np.ascontiguousarray(target_ids, dtype=np.int64)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    out = np.zeros(len(t), dtype=np.int32)
    st = load().b2r_partition_by_target(t.ctypes.data, None if w is None else w.ctypes.data, len(t), int(world_size), out.ctypes.data)
    if st != OK:
        raise B2RError(st, "b2r_partition_by_target failed")
    return out


def select_best_candidates(table, target_ids, fitness_score_thresh=1.25):
    """b2r_select_best_candidates over a RESULT_DTYPE table: (best pair index per target or -1, best score per target)."""
    t = np.ascontiguousarray(target_ids, dtype=np.int64)
    tab = np.ascontiguousarray(table)
    assert tab.dtype == RESULT_DTYPE and len(tab) == len(t)
    best = np.zeros(max(len(t), 1), dtype=np.int64)
    score = np.zeros(max(len(t), 1), dtype=np.float64)
    nt = ctypes.c_size_t()
    st = load().b2r_select_best_candidates(tab.ctypes.data, t.ctypes.data, len(t), fitness_score_thresh, best.ctypes.data, score.ctypes.data,
                                           ctypes.byref(nt))
    if st != OK:
        raise B2RError(st, "b2r_select_best_candidates failed")
    return best[: nt.value].copy(), score[: nt.value].copy()


class Comm:
    """b2r_comm: the ranks of a sharded loop-closure batch.  Comm.nccl(...) runs the all-gather over NCCL / NVLink on the
    handle's stream; Comm.host(...) takes a host all-gather (e.g. torch.distributed gloo) — CPU tests and single-GPU boxes."""

    def __init__(self, h, keep=None):
        self._h = h
        self._keep = keep
        self._lib = load()

    @staticmethod
    def unique_id():
        buf = ctypes.create_string_buffer(UNIQUE_ID_BYTES)
        st = load().b2r_comm_unique_id(buf)
        if st != OK:
            raise B2RError(st, "b2r_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return bytes(buf.raw)

    @classmethod
    def nccl(cls, reg, unique_id, rank, world_size):
        h = ctypes.c_void_p()
        reg._check(load().b2r_comm_init(reg._h, ctypes.c_char_p(unique_id), rank, world_size, ctypes.byref(h)))
        return cls(h)

    @classmethod
    def host(cls, allgather, rank, world_size):
        """allgather(send: bytes) -> list of world_size bytes objects (rank order)."""
        def tramp(user, send, recv, nbytes):
            try:
                parts = allgather(ctypes.string_at(send, nbytes))
                for r, blk in enumerate(parts):
                    ctypes.memmove(recv + r * nbytes, blk, nbytes)
                return 0
            except Exception:  # pragma: no cover
                import traceback
                traceback.print_exc()
                return 1
        fn = ALLGATHER_FN(tramp) if world_size > 1 or allgather is not None else ALLGATHER_FN(0)
        h = ctypes.c_void_p()
        st = load().b2r_comm_init_host(fn, None, rank, world_size, ctypes.byref(h))
        if st != OK:
            raise B2RError(st, "b2r_comm_init_host failed")
        return cls(h, keep=fn)

    @classmethod
    def torch_host(cls, group=None):
        """Host transport over torch.distributed (gloo on CPU, or any backend with CPU tensors)."""
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return cls.host(None, 0, 1)
        ws, rk = dist.get_world_size(group), dist.get_rank(group)

        def allgather(blob):
            t = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
            outs = [torch.empty_like(t) for _ in range(ws)]
            dist.all_gather(outs, t, group=group)
            return [o.numpy().tobytes() for o in outs]
        return cls.host(allgather, rk, ws)

    @property
    def rank(self):
        return int(self._lib.b2r_comm_rank(self._h))

    @property
    def size(self):
        return int(self._lib.b2r_comm_size(self._h))

    def collectives(self):
        return int(self._lib.b2r_comm_collectives(self._h))

    def gather_results(self, rank_of_pair, local_table, reg=None):
        """b2r_gather_results: local_table = this rank's rows (RESULT_DTYPE) in pair order -> all rows in pair order."""
        rp = np.ascontiguousarray(rank_of_pair, dtype=np.int32)
        loc = np.ascontiguousarray(local_table)
        assert loc.dtype == RESULT_DTYPE
        out = np.zeros(len(rp), dtype=RESULT_DTYPE)
        st = self._lib.b2r_gather_results(reg._h if reg is not None else None, self._h, rp.ctypes.data, len(rp),
                                          loc.ctypes.data if len(loc) else None, out.ctypes.data)
        if st != OK:
            raise B2RError(st, self._lib.b2r_comm_last_error(self._h).decode())
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b2r_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CloudBatch:
    """The clouds of one b2r_cloud_create_batch call as an array of handles (numpy uint64): no per-cloud Python objects, one
    destroy call — for hosts that move thousands of handles per batch (bench.py)."""

    def __init__(self, reg, pointers, sizes, memspace, stride=16):
        n = len(pointers)
        P = np.ascontiguousarray(pointers, dtype=np.uint64)
        N = np.ascontiguousarray(sizes, dtype=np.uint64)
        self.handles = np.zeros(max(n, 1), dtype=np.uint64)
        self.n = n
        reg._check(load().b2r_cloud_create_batch(reg._h, P.ctypes.data, N.ctypes.data, n, stride, memspace, self.handles.ctypes.data))

    def close(self):
        if self.handles is not None:
            load().b2r_cloud_destroy_batch(self.handles.ctypes.data, self.n)
            self.handles = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _handle_array(clouds):
    """list of Cloud / None, or a numpy uint64 array of b2r_cloud handles (0 = NULL) -> (pointer, keep-alive object)"""
    if isinstance(clouds, np.ndarray):
        a = np.ascontiguousarray(clouds, dtype=np.uint64)
        return a.ctypes.data, a
    n = len(clouds)
    arr = (ctypes.c_void_p * max(n, 1))(*[c._h if c is not None else None for c in clouds])
    return ctypes.cast(arr, ctypes.c_void_p), arr


class Registration:
    """pcl::Registration-shaped front end over one b2r_handle."""

    def __init__(self, config):
        self._lib = load()
        self.config = config
        h = ctypes.c_void_p()
        st = self._lib.b2r_create(ctypes.byref(config), ctypes.byref(h))
        if st != OK:
            raise B2RError(st, "b2r_create failed (no CUDA device?)" if st == ERR_NO_DEVICE else "b2r_create failed")
        self._h = h
        self.result = None
        self._src_n = 0
        self._tgt_n = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b2r_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != OK:
            raise B2RError(st, self._lib.b2r_last_error(self._h).decode())

    # ---- pcl::Registration surface ----
    def setInputTarget(self, cloud):
        if isinstance(cloud, Cloud):
            self._check(self._lib.b2r_set_target_cloud(self._h, cloud._h))
            self._tgt_n = len(cloud)
            self._keep_t = cloud
        else:
            a = _points(cloud)
            self._check(self._lib.b2r_set_target(self._h, a.ctypes.data, len(a), a.shape[1] * 4, HOST))
            self._tgt_n = len(a)

    def setInputSource(self, cloud):
        if isinstance(cloud, Cloud):
            self._check(self._lib.b2r_set_source_cloud(self._h, cloud._h))
            self._src_n = len(cloud)
            self._keep_s = cloud
        else:
            a = _points(cloud)
            self._check(self._lib.b2r_set_source(self._h, a.ctypes.data, len(a), a.shape[1] * 4, HOST))
            self._src_n = len(a)

    def align(self, guess=None):
        g = colmajor(np.eye(4) if guess is None else guess)
        r = Result()
        self._check(self._lib.b2r_align(self._h, g.ctypes.data, ctypes.byref(r)))
        self.result = r
        return r

    def getFinalTransformation(self):
        return from_colmajor(list(self.result.T))

    def hasConverged(self):
        return bool(self.result.converged)

    def getFitnessScore(self, max_range=np.finfo(np.float64).max):
        out = ctypes.c_double()
        self._check(self._lib.b2r_fitness(self._h, max_range, ctypes.byref(out)))
        return out.value

    def aligned_cloud(self):
        out = np.empty((self._src_n, 4), dtype=np.float32)
        self._check(self._lib.b2r_transform_source(self._h, out.ctypes.data, 16, HOST))
        return out

    # ---- batch (loop closure) ----
    def align_batch(self, sources, targets, guesses, with_fitness=False, fitness_max_range=np.finfo(np.float64).max):
        n = len(sources)
        assert len(targets) == n and len(guesses) == n
        S = (ctypes.c_void_p * n)(*[c._h for c in sources])
        T = (ctypes.c_void_p * n)(*[c._h for c in targets])
        G = np.ascontiguousarray(np.stack([colmajor(g) for g in guesses]) if n else np.zeros((0, 16), np.float32))
        R = (Result * n)()
        self._check(self._lib.b2r_align_batch(self._h, S, T, G.ctypes.data, n, int(with_fitness), fitness_max_range, R))
        return list(R)

    def align_batch_table(self, sources, targets, guesses, with_fitness=False, fitness_max_range=np.finfo(np.float64).max):
        """align_batch returning the results as one numpy structured array (a view of the b2r_result array): no per-pair
        Python objects, for large batches."""
        n = len(sources)
        S = (ctypes.c_void_p * n)(*[c._h for c in sources])
        T = (ctypes.c_void_p * n)(*[c._h for c in targets])
        G = np.ascontiguousarray(np.asarray(guesses, dtype=np.float64).reshape(n, 4, 4).transpose(0, 2, 1).reshape(n, 16).astype(np.float32))
        R = (Result * max(n, 1))()
        self._check(self._lib.b2r_align_batch(self._h, S, T, G.ctypes.data, n, int(with_fitness), fitness_max_range, R))
        return np.frombuffer(R, dtype=RESULT_DTYPE, count=n).copy()

    def align_batch_sharded(self, comm, sources, targets, target_ids, guesses, weights=None, with_fitness=False,
                            fitness_max_range=np.finfo(np.float64).max):
        """b2r_align_batch_sharded: every rank passes the same pair list; sources[i] / targets[i] may be None on ranks that do not
        own pair i.  Returns the whole table (RESULT_DTYPE, pair order) on every rank."""
        n = len(sources)
        S, _ks = _handle_array(sources)
        T, _kt = _handle_array(targets)
        ids = np.ascontiguousarray(target_ids, dtype=np.int64)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        if isinstance(guesses, np.ndarray) and guesses.dtype == np.float32 and guesses.shape == (n, 16):
            G = np.ascontiguousarray(guesses)  # already column-major float32 (e.g. prepared once for a repeated batch)
        else:
            G = np.ascontiguousarray(np.asarray(guesses, dtype=np.float64).reshape(n, 4, 4).transpose(0, 2, 1).reshape(n, 16).astype(np.float32))
        R = np.zeros(max(n, 1), dtype=RESULT_DTYPE)
        st = self._lib.b2r_align_batch_sharded(self._h, comm._h, S, T, ids.ctypes.data, None if w is None else w.ctypes.data, G.ctypes.data, n,
                                               int(with_fitness), fitness_max_range, R.ctypes.data)
        if st != OK:
            raise B2RError(st, self._lib.b2r_comm_last_error(comm._h).decode())
        return R[:n]

    def inlier_fraction(self, max_correspondence_dist=0.5):
        """(inlier fraction, fitness score) of the last alignment: scan_matching_odometry_component.cpp:403-415."""
        frac, fit = ctypes.c_double(), ctypes.c_double()
        self._check(self._lib.b2r_inlier_fraction(self._h, max_correspondence_dist, ctypes.byref(frac), ctypes.byref(fit)))
        return frac.value, fit.value

    def nearest_neighbors(self):
        """b2r_nearest_neighbors: (index of the nearest target point, squared distance, transformed point) per source point."""
        n = self._src_n
        idx = np.zeros(n, dtype=np.int32); d2 = np.zeros(n, dtype=np.float32); xyz = np.zeros((n, 3), dtype=np.float32)
        self._check(self._lib.b2r_nearest_neighbors(self._h, idx.ctypes.data, d2.ctypes.data, xyz.ctypes.data))
        return idx, d2, xyz

    def graph_launches(self):
        return int(self._lib.b2r_graph_launches(self._h))

    def fitness_pair(self, target, source, T, max_range=np.finfo(np.float64).max):
        out = ctypes.c_double()
        g = colmajor(T)
        self._check(self._lib.b2r_fitness_pair(self._h, target._h, source._h, g.ctypes.data, max_range, ctypes.byref(out)))
        return out.value

    # ---- filters ----
    def _filter(self, fn, cloud, *args, extra=None):
        a = _points(cloud)
        out = np.empty((len(a), 4), dtype=np.float32)
        m = ctypes.c_size_t()
        tail = [out.ctypes.data, ctypes.byref(m)] + ([extra] if extra is not None else [])
        self._check(fn(self._h, a.ctypes.data, len(a), a.shape[1] * 4, HOST, *args, *tail))
        return out[: m.value].copy()

    def distance_filter(self, cloud, near, far):
        return self._filter(self._lib.b2r_distance_filter, cloud, near, far)

    def voxelgrid(self, cloud, leaf, min_points=1):
        ovf = ctypes.c_int()
        pts = self._filter(self._lib.b2r_voxelgrid, cloud, leaf, min_points, extra=ctypes.byref(ovf))
        return pts, bool(ovf.value)

    def radius_outlier(self, cloud, radius, min_neighbors):
        return self._filter(self._lib.b2r_radius_outlier, cloud, radius, min_neighbors)

    def statistical_outlier(self, cloud, mean_k, stddev_mul):
        return self._filter(self._lib.b2r_statistical_outlier, cloud, mean_k, stddev_mul)

    def prefilter(self, cloud, cfg=None):
        if cfg is None:
            cfg = PrefilterConfig()
            self._lib.b2r_default_prefilter_config(ctypes.byref(cfg))
        a = _points(cloud)
        out = np.empty((len(a), 4), dtype=np.float32)
        m = ctypes.c_size_t()
        self._check(self._lib.b2r_prefilter(self._h, ctypes.byref(cfg), a.ctypes.data, len(a), a.shape[1] * 4, HOST, out.ctypes.data,
                                            ctypes.byref(m)))
        return out[: m.value].copy()

    def remove_robot_points(self, cloud, others_positions_map, map2sensor, radius):
        """mrg_slam_component.cpp:395-427.  others_positions_map: (m,3) float64 positions of the other robots in the map frame;
        map2sensor: 4x4 float64 (odom^-1 * map2odom).  Returns (kept, removed), both in input order."""
        a = _points(cloud)
        others = np.asarray(others_positions_map, dtype=np.float64).reshape(-1, 3)
        M = np.asarray(map2sensor, dtype=np.float64)
        sensor = (others @ M[:3, :3].T + M[:3, 3]).astype(np.float32)  # (map2sensor * p).cast<float>()  (:399-403)
        sensor = np.ascontiguousarray(sensor)
        r2 = np.float32(np.float64(radius) * np.float64(radius))       # float robot_radius_sqr = double * double (:405-406)
        kept = np.empty((len(a), 4), dtype=np.float32); removed = np.empty((len(a), 4), dtype=np.float32)
        nk, nr = ctypes.c_size_t(), ctypes.c_size_t()
        self._check(self._lib.b2r_remove_robot_points(self._h, a.ctypes.data, len(a), a.shape[1] * 4, HOST, sensor.ctypes.data, len(sensor),
                                                      r2, kept.ctypes.data, ctypes.byref(nk), removed.ctypes.data, ctypes.byref(nr)))
        return kept[: nk.value].copy(), removed[: nr.value].copy()

    def map_cloud(self, clouds, poses, first_keyframe=None, resolution=0.05, min_points_per_voxel=1, distance_far_thresh=-1.0,
                  skip_first_cloud=False):
        """MapCloudGenerator::generate (map_cloud_generator.cpp:14-86).  clouds: list of (n,4) float32 arrays; poses: 4x4 float64
        (world <- keyframe).  Returns the map cloud (voxels in ascending (z, y, x) order) or None where the reference returns nullptr."""
        cs = [_points(c) for c in clouds]
        n = len(cs)
        ptrs = (ctypes.c_void_p * max(n, 1))(*[c.ctypes.data for c in cs])
        ns = (ctypes.c_size_t * max(n, 1))(*[len(c) for c in cs])
        P = np.ascontiguousarray(np.stack([np.asarray(p, dtype=np.float64).T.reshape(16) for p in poses])) if n else np.zeros((0, 16))
        fk = np.zeros(max(n, 1), dtype=np.uint8) if first_keyframe is None else np.ascontiguousarray(first_keyframe, dtype=np.uint8)
        out = np.empty((max(sum(len(c) for c in cs), 1), 4), dtype=np.float32)
        m, null = ctypes.c_size_t(), ctypes.c_int()
        stride = cs[0].shape[1] * 4 if n else 16
        self._check(self._lib.b2r_map_cloud(self._h, ptrs, ns, P.ctypes.data, fk.ctypes.data, n, stride, HOST, resolution, min_points_per_voxel,
                                            distance_far_thresh, int(skip_first_cloud), out.ctypes.data, ctypes.byref(m), ctypes.byref(null)))
        return None if null.value else out[: m.value].copy()

    # ---- introspection ----
    def kernel_launches(self):
        return int(self._lib.b2r_kernel_launches(self._h))

    def last_timings(self):
        t = (ctypes.c_float * 4)()
        self._lib.b2r_last_timings(self._h, t)
        return dict(zip(("prep_ms", "optimize_ms", "fitness_ms", "total_ms"), list(t)))

    def event_record(self, slot):
        self._check(self._lib.b2r_event_record(self._h, slot))

    def event_elapsed_ms(self, a, b):
        ms = ctypes.c_float()
        self._check(self._lib.b2r_event_elapsed_ms(self._h, a, b, ctypes.byref(ms)))
        return ms.value

    def profile_enable(self, on=True):
        self._check(self._lib.b2r_profile_enable(self._h, int(on)))

    def profile_read(self, kernel):
        ms, n, by = ctypes.c_double(), ctypes.c_uint64(), ctypes.c_double()
        self._check(self._lib.b2r_profile_read(self._h, PROFILE_KERNELS[kernel], ctypes.byref(ms), ctypes.byref(n), ctypes.byref(by)))
        return dict(ms=ms.value, launches=int(n.value), bytes=by.value)

    def knn_list_overflows(self):
        v = ctypes.c_uint64(0)
        self._check(self._lib.b2r_debug_knn_list_overflows(self._h, ctypes.byref(v)))
        return int(v.value)

    def synchronize(self):
        self._check(self._lib.b2r_synchronize(self._h))

    def debug_covariances(self, which, want_knn=False):
        n = self._src_n if which == 0 else self._tgt_n
        k = self.config.correspondence_randomness
        cov = np.zeros((n, 6))
        knn = np.zeros((n, k), dtype=np.int32) if want_knn else None
        self._check(self._lib.b2r_debug_covariances(self._h, which, cov.ctypes.data, knn.ctypes.data if want_knn else None))
        return (cov, knn) if want_knn else cov

    def debug_voxelmap(self):
        n = self._tgt_n
        coords = np.zeros((n, 3), dtype=np.int32); npts = np.zeros(n, dtype=np.int32)
        mean = np.zeros((n, 3)); cov = np.zeros((n, 6))
        V = ctypes.c_size_t()
        self._check(self._lib.b2r_debug_voxelmap(self._h, coords.ctypes.data, npts.ctypes.data, mean.ctypes.data, cov.ctypes.data, ctypes.byref(V)))
        v = V.value
        return coords[:v].copy(), npts[:v].copy(), mean[:v].copy(), cov[:v].copy()

    def debug_linearize(self, T):
        T = np.ascontiguousarray(T, dtype=np.float64)
        H = np.zeros((6, 6)); b = np.zeros(6); err = ctypes.c_double()
        n = self._src_n
        if self.config.method == FAST_VGICP:
            corr = np.zeros((n, 3), dtype=np.int32); valid = np.zeros(n, dtype=np.uint8)
        else:
            corr = np.zeros(n, dtype=np.int32); valid = np.zeros(n, dtype=np.uint8)
        self._check(self._lib.b2r_debug_linearize(self._h, T.ctypes.data, H.ctypes.data, b.ctypes.data, ctypes.byref(err), corr.ctypes.data,
                                                  valid.ctypes.data))
        if self.config.method != FAST_VGICP:
            valid = corr >= 0
        return err.value, H, b, corr, valid.astype(bool)

    def debug_compute_error(self, T_lin, T_trial):
        a = np.ascontiguousarray(T_lin, dtype=np.float64); b = np.ascontiguousarray(T_trial, dtype=np.float64)
        err = ctypes.c_double()
        self._check(self._lib.b2r_debug_compute_error(self._h, a.ctypes.data, b.ctypes.data, ctypes.byref(err)))
        return err.value

    def debug_ndt_grid(self):
        n = self._tgt_n
        idx = np.zeros(n, dtype=np.int32); npts = np.zeros(n, dtype=np.int32)
        mean = np.zeros((n, 3)); icov = np.zeros((n, 9))
        min_b = np.zeros(3, dtype=np.int32); div_b = np.zeros(3, dtype=np.int32)
        V = ctypes.c_size_t()
        self._check(self._lib.b2r_debug_ndt_grid(self._h, idx.ctypes.data, npts.ctypes.data, mean.ctypes.data, icov.ctypes.data,
                                                 min_b.ctypes.data, div_b.ctypes.data, ctypes.byref(V)))
        v = V.value
        return idx[:v].copy(), npts[:v].copy(), mean[:v].copy(), icov[:v].reshape(v, 3, 3).copy(), min_b, div_b

    def debug_ndt_derivatives(self, p6):
        p6 = np.ascontiguousarray(p6, dtype=np.float64)
        score = ctypes.c_double(); g = np.zeros(6); H = np.zeros((6, 6)); hits = np.zeros(self._src_n, dtype=np.int32)
        self._check(self._lib.b2r_debug_ndt_derivatives(self._h, p6.ctypes.data, ctypes.byref(score), g.ctypes.data, H.ctypes.data,
                                                        hits.ctypes.data))
        return score.value, g, H, hits

    def debug_knn(self, cloud, queries, k):
        q = _points(queries)
        idx = np.zeros((len(q), k), dtype=np.int32); d2 = np.zeros((len(q), k), dtype=np.float32)
        self._check(self._lib.b2r_debug_knn(self._h, cloud._h, q.ctypes.data, len(q), k, idx.ctypes.data, d2.ctypes.data))
        return idx, d2


def select_registration_method(params):
    """Python mirror of select_registration_method() (/root/reference/src/mrg_slam/registrations.cpp:28-152): the same chain of
    string tests in the same order as the C++ mirror (include/b2r/registration.hpp), so one parameter set gives one engine
    whatever the entry point.

    `params` is a dict carrying the ROS parameter names the reference reads at :34-43.
      "ICP"                      outside the engine -> None (the C++ mirror's nullptr)
      "FAST_VGICP_CUDA"          this engine's FAST_VGICP (it IS a CUDA VGICP; upstream only under USE_VGICP_CUDA)
      *GICP* (with/without OMP)  pcl / pclomp GeneralizedIterativeClosestPoint (BFGS)                       (:93-116)
      anything else              NDT; strings without "NDT" warn first (:117-120); without "OMP" ->
                                 pcl::NormalDistributionsTransform = kd-tree radius search (KDTREE); with "OMP" -> pclomp NDT with
                                 reg_nn_search_method KDTREE / DIRECT1 / else DIRECT7                      (:121-146)
    reg_use_reciprocal_correspondences is accepted and has no effect, as upstream: pcl / pclomp GICP never read the flag.
    """
    import sys

    name = params.get("registration_method", "FAST_GICP")
    common = dict(
        device=params.get("device", 0),
        transformation_epsilon=params.get("reg_transformation_epsilon", 0.1),
        maximum_iterations=params.get("reg_maximum_iterations", 64),
    )
    if name == "SMALL_GICP":  # registrations.cpp:46-54: epsilon, iterations, correspondence distance and randomness are set
        cfg = default_config(SMALL_GICP, max_correspondence_distance=params.get("reg_max_correspondence_distance", 2.0),
                             correspondence_randomness=params.get("reg_correspondence_randomness", 20), **common)
    elif name == "FAST_GICP":
        cfg = default_config(FAST_GICP, max_correspondence_distance=params.get("reg_max_correspondence_distance", 2.0),
                             correspondence_randomness=params.get("reg_correspondence_randomness", 20), **common)
    elif name in ("FAST_VGICP", "FAST_VGICP_CUDA"):
        cfg = default_config(FAST_VGICP, resolution=params.get("reg_resolution", 1.0),
                             correspondence_randomness=params.get("reg_correspondence_randomness", 20), **common)
    elif name == "ICP":
        print("b2r: registration_method ICP (pcl::IterativeClosestPoint) is outside this engine's scope", file=sys.stderr)
        return None
    elif "GICP" in name:  # registrations.cpp:93-116: pcl / pclomp GeneralizedIterativeClosestPoint (BFGS)
        cfg = default_config(GICP_PCL, max_correspondence_distance=params.get("reg_max_correspondence_distance", 2.0),
                             correspondence_randomness=params.get("reg_correspondence_randomness", 20),
                             max_optimizer_iterations=params.get("reg_max_optimizer_iterations", 20), **common)
    else:
        if "NDT" not in name:
            print(f"warning: unknown registration type({name})\n       : use NDT", file=sys.stderr)
        if "OMP" not in name:  # pcl::NormalDistributionsTransform (:122-128)
            nn = KDTREE
        else:
            nn = {"KDTREE": KDTREE, "DIRECT1": DIRECT1}.get(params.get("reg_nn_search_method", "DIRECT7"), DIRECT7)
        cfg = default_config(NDT_OMP, resolution=params.get("reg_resolution", 1.0), neighbor_search=nn, **common)
    return Registration(cfg)
