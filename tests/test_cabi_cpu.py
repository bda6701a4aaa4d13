"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/b2r.h declares,
fails loudly without a GPU (no CPU fallback), and the host mirrors carry the reference's parameter values."""
import ctypes
import os
import re

import numpy as np
import pytest

from mrg_slam_b200 import lib as B
from mrg_slam_b200 import loop_closure as LC
from mrg_slam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    L = B.load()
    cfg = B.default_config(B.FAST_VGICP)
    h = ctypes.c_void_p()
    st = L.b2r_create(ctypes.byref(cfg), ctypes.byref(h))
    if st == B.OK:
        L.b2r_destroy(h)
    return st == B.OK


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "b2r.h")).read()
    declared = set(re.findall(r"\b(b2r_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(B.EXPORTED_SYMBOLS), declared ^ set(B.EXPORTED_SYMBOLS)
    L = B.load()
    for name in declared:
        assert hasattr(L, name), f"libb2r.so does not export {name}"
    assert b"sm_100a" in L.b2r_version()


def test_default_config_carries_yaml_values():
    # config/mrg_slam.yaml:100-109 and the upstream defaults mrg_slam never overrides
    c = B.default_config("FAST_VGICP")
    assert (c.transformation_epsilon, c.maximum_iterations, c.correspondence_randomness, c.resolution) == (0.1, 64, 20, 1.0)
    assert c.neighbor_search == B.DIRECT1 and c.rotation_epsilon == 2e-3 and c.lm_max_iterations == 10 and c.lm_init_lambda_factor == 1e-9
    g = B.default_config("FAST_GICP")
    assert g.max_correspondence_distance == 2.0
    n = B.default_config("NDT_OMP")
    assert n.neighbor_search == B.DIRECT7 and n.ndt_step_size == 0.1 and n.ndt_outlier_ratio == 0.55
    p = B.PrefilterConfig()
    B.load().b2r_default_prefilter_config(ctypes.byref(p))
    # config/mrg_slam.yaml:48-64
    assert (p.distance_near_thresh, p.distance_far_thresh, p.downsample_resolution) == (0.1, 35.0, np.float32(0.1))
    assert (p.outlier_removal_method, p.radius_radius, p.radius_min_neighbors, p.statistical_mean_k, p.statistical_stddev) == (2, 0.5, 2, 30, 1.2)


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to construct: nothing is computed on the host."""
    if _has_cuda():
        pytest.skip("a CUDA device is present")
    L = B.load()
    cfg = B.default_config(B.NDT_OMP)
    h = ctypes.c_void_p()
    assert L.b2r_create(ctypes.byref(cfg), ctypes.byref(h)) == B.ERR_NO_DEVICE
    assert not h.value
    with pytest.raises(B.B2RError) as e:
        B.Registration(cfg)
    assert e.value.status == B.ERR_NO_DEVICE
    with pytest.raises(B.B2RError):
        B.select_registration_method({"registration_method": "FAST_VGICP"})
    # null-handle calls are rejected, not crashed
    assert L.b2r_align(None, None, None) == B.ERR_INVALID_ARG
    assert L.b2r_kernel_launches(None) == 0


def test_product_does_not_import_the_oracle():
    """Only tests/, smoke() and bench.py's baseline legs may touch oracle/ (③)."""
    pkg = os.path.join(ROOT, "mrg_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oraclelib" not in src and "liboracle" not in src and "oracle.h" not in src, os.path.join(dirpath, f)


def test_factory_string_dispatch(monkeypatch, capsys):
    """select_registration_method mirrors registrations.cpp:46-148: string -> engine config; unknown -> warn + NDT."""
    made = []

    class Fake:
        def __init__(self, cfg):
            made.append(cfg)

    monkeypatch.setattr(B, "Registration", Fake)
    B.select_registration_method({"registration_method": "FAST_VGICP", "reg_resolution": 0.7, "reg_transformation_epsilon": 0.05})
    assert made[-1].method == B.FAST_VGICP and made[-1].resolution == 0.7 and made[-1].transformation_epsilon == 0.05
    B.select_registration_method({"registration_method": "FAST_GICP", "reg_max_correspondence_distance": 3.0})
    assert made[-1].method == B.FAST_GICP and made[-1].max_correspondence_distance == 3.0
    B.select_registration_method({"registration_method": "NDT_OMP", "reg_nn_search_method": "DIRECT1", "reg_resolution": 0.5})
    assert made[-1].method == B.NDT_OMP and made[-1].neighbor_search == B.DIRECT1 and made[-1].resolution == 0.5
    B.select_registration_method({"registration_method": "NDT_OMP", "reg_nn_search_method": "whatever"})
    assert made[-1].neighbor_search == B.DIRECT7  # registrations.cpp:144-146: anything else -> DIRECT7
    B.select_registration_method({"registration_method": "NDT_OMP", "reg_nn_search_method": "KDTREE"})   # :140-141
    assert made[-1].method == B.NDT_OMP and made[-1].neighbor_search == B.KDTREE
    # :117-129: a string without "NDT" warns; without "OMP" the reference hands out pcl::NormalDistributionsTransform, whose search is
    # the kd-tree radius search (KDTREE), whatever reg_nn_search_method says
    B.select_registration_method({"registration_method": "BOGUS", "reg_nn_search_method": "DIRECT1"})
    assert made[-1].method == B.NDT_OMP and made[-1].neighbor_search == B.KDTREE
    assert "unknown registration type(BOGUS)" in capsys.readouterr().err
    B.select_registration_method({"registration_method": "NDT", "reg_resolution": 2.0})
    assert made[-1].method == B.NDT_OMP and made[-1].neighbor_search == B.KDTREE and made[-1].resolution == 2.0
    assert "unknown" not in capsys.readouterr().err
    B.select_registration_method({"registration_method": "BOGUS_OMP"})  # unknown but "OMP": pclomp NDT, DIRECT7
    assert made[-1].method == B.NDT_OMP and made[-1].neighbor_search == B.DIRECT7
    B.select_registration_method({"registration_method": "SMALL_GICP", "reg_max_correspondence_distance": 1.5})  # the YAML default
    assert made[-1].method == B.SMALL_GICP and made[-1].max_correspondence_distance == 1.5
    B.select_registration_method({"registration_method": "GICP_OMP", "reg_max_optimizer_iterations": 9})  # registrations.cpp:104-116
    assert made[-1].method == B.GICP_PCL and made[-1].max_optimizer_iterations == 9 and made[-1].gicp_epsilon == 1e-3
    B.select_registration_method({"registration_method": "GICP"})                                          # :93-103
    assert made[-1].method == B.GICP_PCL and made[-1].max_optimizer_iterations == 20 and made[-1].max_correspondence_distance == 2.0
    B.select_registration_method({"registration_method": "MY_GICP_VARIANT", "reg_use_reciprocal_correspondences": True})  # find("GICP")
    assert made[-1].method == B.GICP_PCL
    B.select_registration_method({"registration_method": "FAST_VGICP_CUDA", "reg_resolution": 0.5})
    assert made[-1].method == B.FAST_VGICP and made[-1].resolution == 0.5
    n = len(made)
    assert B.select_registration_method({"registration_method": "ICP"}) is None and len(made) == n  # outside the engine


def test_synth_is_deterministic_and_shaped():
    a1, a2 = synth.scan(synth.VLP16, 7), synth.scan(synth.VLP16, 7)
    assert np.array_equal(a1, a2) and a1.dtype == np.float32 and a1.shape[1] == 4
    assert 0.7 * 28800 < len(a1) <= 28800
    b = synth.scan(synth.VLP16, 8)
    assert not np.array_equal(a1[: min(len(a1), len(b))], b[: min(len(a1), len(b))])
    step = np.linalg.norm((np.linalg.inv(synth.pose(7)) @ synth.pose(8))[:3, 3])
    assert 0.3 < step < 0.6
    h = synth.scan(synth.HDL64, 0)
    assert 0.8 * 121600 < len(h) <= 121600
    r = np.linalg.norm(h[:, :3], axis=1)
    assert r.max() <= 121.0 and np.isfinite(h).all() and (h[:, 3] >= 0).all() and (h[:, 3] < 1).all()


# ------------------------------------------------------------------------------ loop-closure host logic
def test_select_best_tie_rule():
    # loop_detector.cpp:138: `score > best_score -> continue`, so on equal scores the LATER candidate wins
    assert LC.select_best([0.5, 0.3, 0.3, 0.9], [True, True, True, True]) == (2, 0.3)
    assert LC.select_best([0.1, 0.3], [False, True]) == (1, 0.3)
    assert LC.select_best([0.1], [False]) == (None, LC.DBL_MAX)
    assert LC.select_best([], []) == (None, LC.DBL_MAX)


def test_partition_by_target():
    targets = [0] * 16 + [1] * 16 + [2] * 3 + [3] * 16 + [4] * 13
    for ws in (1, 2, 3, 4, 8):
        shards = LC.partition_by_target(targets, ws)
        assert len(shards) == ws
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(targets)))  # every pair exactly once
        for s in shards:  # a target never straddles two ranks
            for t in {targets[i] for i in s}:
                assert all(i in s for i, tt in enumerate(targets) if tt == t)
    big = LC.partition_by_target([i // 16 for i in range(4096)], 8)
    assert [len(s) for s in big] == [512] * 8
    assert LC.partition_by_target([], 4) == [[], [], [], []]


def test_grouped_best_candidate_rule_matches_the_sequential_one():
    """select_best_grouped (vectorised) == select_best (loop_detector.cpp:106-145 restated), ties and all-unconverged groups included."""
    from mrg_slam_b200 import loop_closure as LC
    rng = np.random.default_rng(11)
    scores = rng.integers(0, 4, size=(200, 7)).astype(np.float64) * 0.25
    conv = rng.random((200, 7)) > 0.3
    conv[5] = False
    best, sc = LC.select_best_grouped(scores, conv)
    for g in range(200):
        b, s_ = LC.select_best(scores[g], conv[g])
        assert (b if b is not None else -1) == best[g] and s_ == sc[g]


def test_partition_by_target_balances_weights():
    from mrg_slam_b200 import loop_closure as LC
    targets = [i // 16 for i in range(4096)]
    w = [30000 - 5 * (i // 16) for i in range(4096)]  # clouds shrink along the trajectory
    for ws in (2, 4, 8):
        shards = LC.partition_by_target(targets, ws, w)
        assert sorted(i for s in shards for i in s) == list(range(4096))
        loads = [sum(w[i] for i in s) for s in shards]
        assert max(loads) <= 1.02 * sum(w) / ws  # within one target group of the ideal share
        for s in shards:  # targets never split
            assert len({targets[i] for i in s}) * 16 == len(s)


def test_partition_by_target_matches_the_stated_rule_on_interleaved_ids():
    """b2r_partition_by_target against a direct restatement of its rule (targets in order of first appearance, contiguous blocks,
    a target goes where most of its weight falls), with target ids that repeat non-consecutively and arbitrary weights."""
    from mrg_slam_b200 import lib as B
    rng = np.random.default_rng(7)
    for trial in range(30):
        n = int(rng.integers(1, 400))
        ids = rng.integers(-5, 40, n).astype(np.int64) * 1000003
        w = rng.uniform(0.5, 3.0, n) if trial % 2 else None
        ws = int(rng.integers(1, 9))
        got = B.partition_by_target(ids, ws, w)
        order, gw, total = [], {}, 0.0
        for i, t in enumerate(ids):
            if t not in gw:
                order.append(t); gw[t] = 0.0
            gw[t] += 1.0 if w is None else w[i]
            total += 1.0 if w is None else w[i]  # the same summation order as the library
        rank, acc, rank_of_target = 0, 0.0, {}
        for t in order:
            while rank < ws - 1 and acc + 0.5 * gw[t] >= (rank + 1) * total / ws:
                rank += 1
            rank_of_target[t] = rank
            acc += gw[t]
        assert list(got) == [rank_of_target[t] for t in ids]
