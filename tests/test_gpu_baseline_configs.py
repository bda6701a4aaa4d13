"""GPU parity at the BASELINE.json configuration sizes: the CUDA path through the C ABI against the CPU
oracle on the very workloads the benchmarks run —

  configs[1]  the first 8 consecutive-scan pairs of bench.py's chain (prefiltered HDL-64, ~44k points, identity guess)
  configs[3]  the first 32 pairs of bench.py's 4096-pair loop-closure batch (~20k points, guesses off by U(+-1 m, +-0.1 rad)),
              FAST_VGICP / FAST_GICP / NDT_OMP, align + getFitnessScore — where an LM-iteration flip would be most likely
              (/root/reference/src/mrg_slam/loop_detector.cpp:129-145; config/mrg_slam.yaml:100-109)
  configs[4]  4 keyframe -> accumulated-submap pairs (targets of ~135k points), FAST_VGICP

Bars (north_star): converged / iterations / evaluation counts identical, final transform within 1e-4 m / 1e-4 rad, fitness
within 1e-3 relative.  Also here: the A15 inlier fraction and the sharded batch call over NCCL (one rank) at full pair count.
"""
import numpy as np
import pytest

import bench
from mrg_slam_b200 import lib as B
from mrg_slam_b200 import loop_closure as LC
from mrg_slam_b200 import synth
from tests import oraclelib as O
from tests.conftest import oracle_prefilter, pose_error

pytestmark = pytest.mark.gpu

T_TOL, R_TOL, FIT_RTOL = 1e-4, 1e-4, 1e-3  # north_star tolerances


def _check(row_T, row_conv, row_iter, row_evals, row_fit, o, ro, with_fitness=True):
    To = o.getFinalTransformation()
    te, re = pose_error(To, B.from_colmajor(row_T))
    assert bool(row_conv) == bool(ro.converged)
    assert row_iter == ro.iterations and row_evals == ro.lm_evals, (row_iter, ro.iterations, row_evals, ro.lm_evals)
    assert te <= T_TOL and re <= R_TOL, (te, re)
    if with_fitness:
        fo = o.getFitnessScore()
        assert abs(row_fit - fo) <= FIT_RTOL * abs(fo), (row_fit, fo)
    return te, re


def test_config1_chain_pairs_match_oracle():
    """bench.py's chain workload itself (configs[1] shape): scan i+1 onto scan i, identity guess, 44k-point clouds."""
    n = 8
    raws = [synth.scan(synth.HDL64, bench.CHAIN_SCAN0 + i) for i in range(n + 1)]
    reg = B.Registration(B.default_config(B.FAST_VGICP))
    clouds = [reg.prefilter(r) for r in raws]
    for r, c in zip(raws[:2], clouds[:2]):
        assert np.array_equal(c, oracle_prefilter(r))  # the engine's prefilter is the oracle's, bit for bit
    cl = [B.Cloud(reg, c) for c in clouds]
    res = reg.align_batch(cl[1:], cl[:-1], [np.eye(4)] * n, with_fitness=True)
    o = O.Registration(O.default_params(O.FAST_VGICP))
    worst = 0.0
    for i in range(n):
        o.setInputTarget(clouds[i]); o.setInputSource(clouds[i + 1])
        ro = o.align(np.eye(4))
        te, _ = _check(list(res[i].T), res[i].converged, res[i].iterations, res[i].evals, res[i].fitness, o, ro)
        worst = max(worst, te)
        assert len(clouds[i]) > 40000
    assert worst <= T_TOL
    for c in cl:
        c.close()
    reg.close()


@pytest.fixture(scope="module")
def config3():
    """The first 32 pairs of bench.py's loop-closure batch: 2 new keyframes x 16 candidates."""
    n_targets, n_cand = 256, 16
    poses = [synth.pose(bench.FIRST_SCAN + i) for i in range(n_targets + n_cand)]
    pairs, guesses = bench.batch_pairs(n_targets, n_cand, poses)
    idx = list(range(32))
    needed = sorted({c for i in idx for c in pairs[i]})
    pool = bench.oracle_pool(O, needed)
    assert 15000 < np.mean([len(c) for c in pool.values()]) < 30000
    return pairs, guesses, idx, pool


@pytest.mark.parametrize("method", ["FAST_VGICP", "FAST_GICP", "NDT_OMP"])
def test_config3_loop_closure_pairs_match_oracle(config3, method):
    pairs, guesses, idx, pool = config3
    reg = B.Registration(B.default_config(getattr(B, method)))
    cl = {c: B.Cloud(reg, pts) for c, pts in pool.items()}
    tab = reg.align_batch_table([cl[pairs[i][1]] for i in idx], [cl[pairs[i][0]] for i in idx], [guesses[i] for i in idx], with_fitness=True)
    o = O.Registration(O.default_params(getattr(O, method)))
    last_t = None
    scores, conv = [], []
    for j, i in enumerate(idx):
        ti, ci = pairs[i]
        if ti != last_t:
            o.setInputTarget(pool[ti]); last_t = ti
        o.setInputSource(pool[ci])
        ro = o.align(guesses[i])
        _check(tab["T"][j], tab["converged"][j], tab["iterations"][j], tab["evals"][j], tab["fitness"][j], o, ro)
        scores.append(o.getFitnessScore()); conv.append(bool(ro.converged))
    # the best-candidate decision (loop_detector.cpp:137-144) of both new keyframes is the oracle's
    ids = np.array([pairs[i][0] for i in idx])
    best, score = B.select_best_candidates(tab, ids, 1.25)
    for k, t in enumerate(sorted(set(ids))):
        rows = [j for j in range(len(idx)) if ids[j] == t]
        want, want_score = LC.select_best([scores[j] for j in rows], [conv[j] for j in rows])
        got = -1 if best[k] < 0 else rows.index(int(best[k]))
        assert got == (-1 if (want is None or want_score > 1.25) else want)
    for c in cl.values():
        c.close()
    reg.close()


def test_config4_keyframe_to_submap_matches_oracle():
    """3-robot keyframe -> accumulated submap (union of 10 neighbouring prefiltered scans, VoxelGrid 0.1: ~135k points)."""
    reg = B.Registration(B.default_config(B.FAST_VGICP))
    rng = np.random.default_rng(0x5EED0005)
    nb = 10
    cases = []
    for r in range(2):
        first = bench.FIRST_SCAN + 400 * r
        for k in (0, 1):
            base = synth.pose(first + k)
            parts = []
            for j in range(k, k + nb):
                rel = (np.linalg.inv(base) @ synth.pose(first + j)).astype(np.float32)
                raw = reg.prefilter(synth.scan(synth.HDL64, first + j))
                parts.append(np.concatenate([raw[:, :3] @ rel[:3, :3].T + rel[:3, 3], raw[:, 3:]], axis=1).astype(np.float32))
            sub, _ = reg.voxelgrid(np.concatenate(parts), 0.1)
            kf, _ = reg.voxelgrid(reg.prefilter(synth.scan(synth.HDL64, first + k + nb // 2)), bench.LEAF)
            gt = np.linalg.inv(base) @ synth.pose(first + k + nb // 2)
            cases.append((sub, kf, gt @ bench.perturbation(rng, 0.5, 0.05)))
    assert min(len(s) for s, _, _ in cases) > 100000
    subs = [B.Cloud(reg, s) for s, _, _ in cases]
    kfs = [B.Cloud(reg, k) for _, k, _ in cases]
    tab = reg.align_batch_table(kfs, subs, [g for _, _, g in cases], with_fitness=True)
    o = O.Registration(O.default_params(O.FAST_VGICP))
    for j, (sub, kf, g) in enumerate(cases):
        o.setInputTarget(sub); o.setInputSource(kf)
        ro = o.align(g)
        _check(tab["T"][j], tab["converged"][j], tab["iterations"][j], tab["evals"][j], tab["fitness"][j], o, ro)
    for c in subs + kfs:
        c.close()
    reg.close()


def test_inlier_fraction_matches_kdtree(vlp16_pair):
    """A15: ScanMatchingOdometryComponent::publish_scan_matching_status (scan_matching_odometry_component.cpp:403-415)."""
    from scipy.spatial import cKDTree
    a, b, gt = vlp16_pair
    reg = B.Registration(B.default_config(B.FAST_VGICP))
    reg.setInputTarget(a); reg.setInputSource(b)
    reg.align(np.eye(4))
    aligned = reg.aligned_cloud()  # the `aligned` cloud of align(): float transform of the source
    frac, fit = reg.inlier_fraction(0.5)
    # exact float32 squared distances with FLANN's association, brute force over the kd-tree's 4 nearest candidates
    _, nn = cKDTree(a[:, :3].astype(np.float64)).query(aligned[:, :3].astype(np.float64), k=4)
    d = aligned[:, None, :3] - a[nn, :3]
    d2 = ((d[..., 0] * d[..., 0]).astype(np.float32) + (d[..., 1] * d[..., 1]).astype(np.float32)).astype(np.float32)
    d2 = (d2 + (d[..., 2] * d[..., 2]).astype(np.float32)).astype(np.float32).min(axis=1)
    want = np.float32(np.count_nonzero(d2.astype(np.float64) < 0.5 * 0.5)) / np.float32(len(aligned))
    assert frac == float(want)
    assert abs(fit - reg.getFitnessScore()) <= 1e-12 * abs(fit)
    # the table behind pcl::Registration::getFitnessScore's nearestKSearch loop (b2r_nearest_neighbors, pcl_adapter.hpp)
    idx, nd2, xyz = reg.nearest_neighbors()
    assert np.array_equal(xyz, aligned[:, :3]) and np.array_equal(nd2, d2)
    dd = aligned[:, :3] - a[idx, :3]
    chk = ((dd[:, 0] * dd[:, 0]).astype(np.float32) + (dd[:, 1] * dd[:, 1]).astype(np.float32)).astype(np.float32)
    assert np.array_equal((chk + (dd[:, 2] * dd[:, 2]).astype(np.float32)).astype(np.float32), nd2)  # the index is that of a nearest point
    assert abs(float(np.mean(nd2.astype(np.float64))) - fit) <= 1e-12 * fit
    assert 0.5 < frac <= 1.0
    f2, _ = reg.inlier_fraction(0.05)
    assert f2 < frac
    reg.close()


def test_sharded_batch_over_nccl_equals_plain_batch(config3):
    """b2r_align_batch_sharded through libb2r's own NCCL communicator (one rank on this GPU: ncclCommInitRank + ncclAllGather run for
    real) returns the rows of b2r_align_batch bit for bit, and a degenerate candidate only loses itself."""
    pairs, guesses, idx, pool = config3
    reg = B.Registration(B.default_config(B.FAST_VGICP))
    cl = {c: B.Cloud(reg, pts) for c, pts in pool.items()}
    src = [cl[pairs[i][1]] for i in idx]; tgt = [cl[pairs[i][0]] for i in idx]
    g = [guesses[i] for i in idx]
    ids = np.array([pairs[i][0] for i in idx], dtype=np.int64)
    plain = reg.align_batch_table(src, tgt, g, with_fitness=True)
    comm = LC.make_comm(reg, 0, 1, nccl=True)
    n0 = comm.collectives()
    shard = reg.align_batch_sharded(comm, src, tgt, ids, g, with_fitness=True)
    assert comm.collectives() == n0 + 1
    assert shard.tobytes() == plain.tobytes()
    tiny = B.Cloud(reg, pool[pairs[0][1]][:7])
    src2 = list(src); src2[3] = tiny
    shard2 = reg.align_batch_sharded(comm, src2, tgt, ids, g, with_fitness=True)
    assert shard2["converged"][3] == 0 and shard2["fitness"][3] == np.finfo(np.float64).max
    keep = np.arange(len(idx)) != 3
    assert shard2[keep].tobytes() == plain[keep].tobytes()
    comm.close(); tiny.close()
    for c in cl.values():
        c.close()
    reg.close()
