"""Host logic of the loop-closure batch path: candidate loop, best-score rule, and sharding over ranks.

Mirrors LoopDetector::matching (/root/reference/src/mrg_slam/loop_detector.cpp:97-180): one target (the new
keyframe) against K candidate sources; per candidate align + getFitnessScore, keep the best converged one
(`score > best_score -> skip`, so on equal scores the LATER candidate wins, :138-140); accept iff
best_score <= fitness_score_thresh (:156).  The K aligns are independent, so batches of (target, candidate)
pairs are sharded across GPUs by target id — all candidates of one keyframe on one rank, so its voxel map /
covariances are built once — and the fixed-size results are all-gathered (NCCL on GPUs, gloo in CPU tests).
There is no collective inside the optimiser.

perform_loop_closure_consistency_check (:190-303, enabled in config/mrg_slam.yaml:177) follows the candidate loop: the
new keyframe is aligned once more against the best candidate's previous (and, if that fails, next) keyframe and the
composition new -> best -> prev -> new must be the identity within 0.3 m / 3 degrees.  check_consistency() runs those
aligns as up to two further (much smaller) batches through the same sharded path; match_keyframes() is the whole of
LoopDetector::matching for many new keyframes at once.
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

DBL_MAX = float(np.finfo(np.float64).max)
RESULT_WIDTH = 24  # T(16) converged iterations error evals fitness pair_index pad pad


def _int_ids(ids):
    """Arbitrary hashable target ids -> int64 labels in order of first appearance."""
    seen = {}
    return np.array([seen.setdefault(t, len(seen)) for t in ids], dtype=np.int64)


def partition_by_target(target_ids: Sequence[int], world_size: int, weights: Optional[Sequence[float]] = None) -> List[List[int]]:
    """Static block partition of pair indices by target id, balanced by pair count or, if given, by per-pair `weights`
    (e.g. the source cloud sizes: an alignment's cost is proportional to the points it evaluates).

    Targets keep their order of first appearance; each rank gets a contiguous block of targets.  Every pair of
    a given target lands on exactly one rank.  The partition itself is libb2r's (b2r_partition_by_target, the one
    b2r_align_batch_sharded applies on every rank); this wrapper only reshapes it into per-rank index lists.
    """
    from .lib import partition_by_target as c_partition

    rank_of = c_partition(_int_ids(target_ids), world_size, weights)
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i, r in enumerate(rank_of):
        shards[int(r)].append(i)
    return shards


def select_best(scores: Sequence[float], converged: Sequence[bool]):
    """The candidate reduction of loop_detector.cpp:106-145.  Returns (best_index or None, best_score)."""
    best_score, best = DBL_MAX, None
    for i, (s, c) in enumerate(zip(scores, converged)):
        if (not c) or s > best_score:
            continue
        best_score, best = s, i
    return best, best_score


def pack_table(table, pair_indices) -> np.ndarray:
    """pack_results for the structured array of Registration.align_batch_table (vectorised)."""
    out = np.zeros((len(table), RESULT_WIDTH), dtype=np.float64)
    if len(table):
        out[:, :16] = table["T"]
        out[:, 16] = table["converged"]
        out[:, 17] = table["iterations"]
        out[:, 18] = table["error"]
        out[:, 19] = table["evals"]
        out[:, 20] = table["fitness"]
        out[:, 21] = np.asarray(pair_indices, dtype=np.float64)
    return out


def select_best_grouped(scores: np.ndarray, converged: np.ndarray):
    """select_best over equally sized candidate groups at once: scores, converged are (groups, candidates).
    Returns (best index per group or -1, best score per group)."""
    s = np.where(converged, scores, np.inf)
    # minimal score, ties -> the LATER candidate (loop_detector.cpp:138: `score > best_score -> skip`)
    rev = s[:, ::-1]
    best = s.shape[1] - 1 - np.argmin(rev, axis=1)
    best_score = s[np.arange(len(s)), best]
    none = ~np.isfinite(best_score) & ~converged.any(axis=1)
    return np.where(none, -1, best), np.where(none, DBL_MAX, best_score)


def pack_results(results, pair_indices) -> np.ndarray:
    out = np.zeros((len(results), RESULT_WIDTH), dtype=np.float64)
    for j, (r, idx) in enumerate(zip(results, pair_indices)):
        out[j, :16] = list(r.T)
        out[j, 16] = r.converged
        out[j, 17] = r.iterations
        out[j, 18] = r.error
        out[j, 19] = r.evals
        out[j, 20] = r.fitness
        out[j, 21] = idx
    return out


def table_from_results(tab) -> np.ndarray:
    """RESULT_DTYPE rows (pair order) -> the (n, RESULT_WIDTH) float64 table used by the host logic below."""
    return pack_table(tab, np.arange(len(tab)))


def results_from_table(local: np.ndarray):
    """(n, RESULT_WIDTH) float64 rows -> RESULT_DTYPE rows (the wire format of b2r_gather_results)."""
    from .lib import RESULT_DTYPE

    out = np.zeros(len(local), dtype=RESULT_DTYPE)
    if len(local):
        out["T"] = local[:, :16].astype(np.float32)
        out["converged"] = local[:, 16].astype(np.int32)
        out["iterations"] = local[:, 17].astype(np.int32)
        out["error"] = local[:, 18]
        out["evals"] = local[:, 19].astype(np.int32)
        out["fitness"] = local[:, 20]
    return out


def make_comm(reg=None, rank=0, world_size=1, group=None, nccl=False):
    """The communicator of a sharded batch.  nccl=True: ncclCommInitRank on reg's device (the unique id is created by rank 0 and
    broadcast over torch.distributed); otherwise a host all-gather over torch.distributed (gloo) or, for one rank, none at all."""
    from .lib import Comm

    if world_size == 1 and not nccl:
        return Comm.host(None, 0, 1)
    import torch.distributed as dist
    if nccl:
        box = [Comm.unique_id() if rank == 0 else None]
        if world_size > 1:
            dist.broadcast_object_list(box, src=0, group=group)
        return Comm.nccl(reg, box[0], rank, world_size)
    return Comm.torch_host(group)


def gather_results(local: np.ndarray, n_pairs: int, device=None, group=None, counts: Optional[Sequence[int]] = None, comm=None,
                   rank_of_pair=None) -> np.ndarray:
    """All-gathers the per-rank result rows into an (n_pairs, RESULT_WIDTH) array ordered by pair index, through
    b2r_gather_results (libb2r's own gather step; host transport over torch.distributed unless `comm` says otherwise).
    local[:, 21] carries the global pair index of each row."""
    import torch.distributed as dist

    own = comm is None
    if own:
        ws = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        rk = dist.get_rank(group) if ws > 1 else 0
        comm = make_comm(None, rk, ws, group)
    if rank_of_pair is None:  # reconstruct who owns which pair from the indices every rank holds
        import torch
        mine = np.zeros(n_pairs, dtype=np.int32)
        mine[local[:, 21].astype(np.int64)] = 1
        if comm.size > 1:
            t = torch.from_numpy(mine * (comm.rank + 1))
            dist.all_reduce(t, group=group)
            rank_of_pair = t.numpy() - 1
        else:
            rank_of_pair = np.zeros(n_pairs, dtype=np.int32)
    order = np.argsort(local[:, 21], kind="stable") if len(local) else np.zeros(0, dtype=np.int64)
    tab = comm.gather_results(rank_of_pair, results_from_table(local[order]))
    if own:
        comm.close()
    return table_from_results(tab)


LAST_TIMINGS = {}  # host wall clock of the last detect_loops call on this rank: align_ms (partition + batch), gather_ms


@dataclass
class Loop:
    target: int
    best_candidate: Optional[int]  # index into that target's candidate list
    best_score: float
    relative_pose: Optional[np.ndarray]  # new keyframe <- best candidate, 4x4
    source: Optional[int] = None  # cloud index of the best candidate


def _cloud(clouds, i):
    return clouds(i) if callable(clouds) else clouds[i]


def sharded_align(reg, clouds, pairs, guesses, with_fitness=True, fitness_score_max_range=DBL_MAX, rank=0, world_size=1, device=None,
                  group=None, pair_weights=None, comm=None) -> np.ndarray:
    """One batch of (target, source) pairs through b2r_align_batch_sharded: partition by target, this rank's slice aligned, one
    all-gather, the full (len(pairs), RESULT_WIDTH) table on every rank.  `clouds` is a list or a callable index -> cloud (only
    the clouds of this rank's slice need to exist).  `comm`: a lib.Comm (NCCL on GPUs); by default a host transport over
    torch.distributed.  A registration stand-in with only the plain align_batch surface (CPU tests) runs its slice itself and
    the rows go through b2r_gather_results."""
    import time

    t_start = time.perf_counter()
    ids = _int_ids([p[0] for p in pairs])
    own = comm is None
    if own:
        comm = make_comm(None, rank, world_size, group)
    try:
        if hasattr(reg, "align_batch_sharded"):
            from .lib import partition_by_target as c_partition
            rank_of = c_partition(ids, comm.size, pair_weights)
            src = [_cloud(clouds, p[1]) if rank_of[i] == comm.rank else None for i, p in enumerate(pairs)]
            tgt = [_cloud(clouds, p[0]) if rank_of[i] == comm.rank else None for i, p in enumerate(pairs)]
            tab = reg.align_batch_sharded(comm, src, tgt, ids, guesses, weights=pair_weights, with_fitness=with_fitness,
                                          fitness_max_range=fitness_score_max_range)
            t_aligned = t_gathered = time.perf_counter()
            table = table_from_results(tab)
        else:
            from .lib import partition_by_target as c_partition
            rank_of = c_partition(ids, comm.size, pair_weights)
            mine = [i for i in range(len(pairs)) if rank_of[i] == comm.rank]
            if mine:
                local = pack_results(reg.align_batch([_cloud(clouds, pairs[i][1]) for i in mine], [_cloud(clouds, pairs[i][0]) for i in mine],
                                                     [guesses[i] for i in mine], with_fitness=with_fitness,
                                                     fitness_max_range=fitness_score_max_range), mine)
            else:
                local = np.zeros((0, RESULT_WIDTH))
            t_aligned = time.perf_counter()
            table = table_from_results(comm.gather_results(rank_of, results_from_table(local)))
            t_gathered = time.perf_counter()
    finally:
        if own:
            comm.close()
    LAST_TIMINGS.update(align_ms=1e3 * (t_aligned - t_start), gather_ms=1e3 * (t_gathered - t_aligned))
    return table


def detect_loops(reg, clouds, pairs, guesses, fitness_score_max_range=DBL_MAX, fitness_score_thresh=1.25, rank=0, world_size=1,
                 device=None, group=None, pair_weights=None, comm=None):
    """Batched LoopDetector::matching over many new keyframes.

    clouds: list of mrg_slam_b200.lib.Cloud (only those this rank needs may be non-None)
    pairs:  list of (target_cloud_index, source_cloud_index), candidates of one target contiguous and in
            candidate order (the tie rule depends on it)
    Returns (loops per target in order of first appearance, full result table).
    """
    from .lib import from_colmajor

    target_ids = [p[0] for p in pairs]
    table = sharded_align(reg, clouds, pairs, guesses, True, fitness_score_max_range, rank, world_size, device, group, pair_weights, comm)
    # the candidate reduction of loop_detector.cpp:106-160 is libb2r's (b2r_select_best_candidates), the same call a C++ host makes
    from .lib import select_best_candidates

    best_pair, best_score = select_best_candidates(results_from_table(table), _int_ids(target_ids), fitness_score_thresh)
    loops, seen = [], {}
    for i, t in enumerate(target_ids):
        seen.setdefault(t, []).append(i)
    for (t, idxs), bp, sc in zip(seen.items(), best_pair, best_score):
        if bp < 0:
            loops.append(Loop(t, None, float(sc), None))
        else:
            loops.append(Loop(t, idxs.index(int(bp)), float(sc), from_colmajor(table[int(bp), :16]), pairs[int(bp)][1]))
    return loops, table


# ------------------------------------------------------------------------------------------------ consistency check
@dataclass
class KeyframeLinks:
    """What perform_loop_closure_consistency_check reads from the best-matched KeyFrame (loop_detector.cpp:190-303)."""
    first_keyframe: bool = False
    static_keyframe: bool = False
    prev: Optional[int] = None                     # cloud index of prev_edge->to_keyframe
    rel_pose_to_prev: Optional[np.ndarray] = None  # prev_edge->relative_pose(), 4x4
    next: Optional[int] = None                     # cloud index of next_edge->from_keyframe
    rel_pose_from_next: Optional[np.ndarray] = None  # next_edge->relative_pose(), 4x4


def _quat_from_matrix(R, dtype=np.float64):
    """Eigen::Quaternion(rotation matrix) (w, x, y, z), in the scalar type of the matrix; not normalised."""
    R = np.asarray(R, dtype=dtype)
    one, half = dtype(1.0), dtype(0.5)
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        r = np.sqrt(t + one)
        w = half * r
        r = half / r
        return w, (R[2, 1] - R[1, 2]) * r, (R[0, 2] - R[2, 0]) * r, (R[1, 0] - R[0, 1]) * r
    i = 0
    if R[1, 1] > R[0, 0]:
        i = 1
    if R[2, 2] > R[i, i]:
        i = 2
    j, k = (i + 1) % 3, (i + 2) % 3
    r = np.sqrt(R[i, i] - R[j, j] - R[k, k] + one)
    q = [dtype(0)] * 3
    q[i] = half * r
    r = half / r
    w = (R[k, j] - R[j, k]) * r
    q[j] = (R[j, i] + R[i, j]) * r
    q[k] = (R[k, i] + R[i, k]) * r
    return w, q[0], q[1], q[2]


def normalize_estimate(T: np.ndarray) -> np.ndarray:
    """LoopDetector::normalize_estimate (:182-188): rotation -> Eigen::Quaterniond -> normalized -> rotation matrix."""
    T = np.asarray(T, dtype=np.float64)
    w, x, y, z = _quat_from_matrix(T[:3, :3])
    n = np.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    out = np.eye(4)
    out[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                   [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                   [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]]
    out[:3, 3] = T[:3, 3]
    return out


def registration_guess(new_estimate: np.ndarray, other_estimate: np.ndarray, planar: bool = False) -> np.ndarray:
    """( new_keyframe_estimate.inverse() * other_estimate ).matrix().cast<float>() with both estimates normalised
    (loop_detector.cpp:124-131, :235-240, :278-283); z translation zeroed for use_planar_registration_guess."""
    a, b = normalize_estimate(new_estimate), normalize_estimate(other_estimate)
    inv = np.eye(4)
    inv[:3, :3] = a[:3, :3].T
    inv[:3, 3] = -a[:3, :3].T @ a[:3, 3]
    g = (inv @ b).astype(np.float32)
    if planar:
        g[2, 3] = 0.0
    return g


def _delta(M: np.ndarray):
    """Translation norm and Quaternionf(rotation block).angularDistance(Identity) = 2 atan2(|vec|, |w|) of a Matrix4f that
    should be the identity (:248-250, :294-296)."""
    M = np.asarray(M, dtype=np.float32)
    delta_trans = float(np.sqrt(np.float32(M[0, 3] * M[0, 3] + M[1, 3] * M[1, 3] + M[2, 3] * M[2, 3])))
    w, x, y, z = _quat_from_matrix(M[:3, :3], np.float32)
    delta_angle = float(np.float32(2.0) * np.arctan2(np.sqrt(x * x + y * y + z * z), np.abs(w)))
    return delta_trans, delta_angle


def identity_check_prev(T_new_prev, T_new_best, T_cand_prev):
    """rel_pose_new_to_prev.inverse() * rel_pose_new_to_best_matched * rel_pose_candidate_to_prev (:247), Matrix4f."""
    f = lambda M: np.asarray(M, dtype=np.float32)
    return _delta(np.linalg.inv(f(T_new_prev)) @ f(T_new_best) @ f(T_cand_prev))


def identity_check_next(T_new_next, T_new_best, T_next_cand):
    """rel_pose_new_to_best_matched.inverse() * rel_pose_new_to_next * rel_pose_next_to_candidate (:290), Matrix4f."""
    f = lambda M: np.asarray(M, dtype=np.float32)
    return _delta(np.linalg.inv(f(T_new_best)) @ f(T_new_next) @ f(T_next_cand))


def check_consistency(reg, clouds, loops: Sequence[Loop], estimates, links, max_delta_trans=0.3, max_delta_angle=0.0523599,
                      use_planar_registration_guess=False, enable=True, rank=0, world_size=1, device=None, group=None, comm=None):
    """perform_loop_closure_consistency_check (:190-218) for every loop of detect_loops at once.

    estimates: cloud index -> 4x4 graph estimate of that keyframe (node->estimate()); links: cloud index -> KeyframeLinks.
    The prev-keyframe aligns of all loops form one batch; the next-keyframe aligns of those that failed (or have no
    prev edge) a second one, as the reference only tries `next` after `prev` failed.  As in the reference the aligns'
    converged flags are not consulted.  Returns (passed per loop, details per loop)."""
    from .lib import from_colmajor

    n = len(loops)
    passed = [False] * n
    details = [dict() for _ in range(n)]
    todo = []
    for i, lp in enumerate(loops):
        if lp.best_candidate is None:
            continue  # best_matched == nullptr or best_score > fitness_score_thresh: false (:201-204)
        lk = links[lp.source]
        if lk.first_keyframe or lk.static_keyframe:
            passed[i] = True  # :197-199, before the enable flag is looked at
            details[i]["skipped"] = "first_or_static_keyframe"
            continue
        if enable:
            todo.append(i)

    def run(stage, idxs):
        key = "prev" if stage == 0 else "next"
        sel = [i for i in idxs if getattr(links[loops[i].source], key) is not None]
        if not sel:
            return {}
        pairs, guesses = [], []
        for i in sel:
            other = getattr(links[loops[i].source], key)
            pairs.append((loops[i].target, other))  # the target is still the new keyframe (:104)
            guesses.append(registration_guess(estimates[loops[i].target], estimates[other], use_planar_registration_guess).astype(np.float64))
        table = sharded_align(reg, clouds, pairs, guesses, False, DBL_MAX, rank, world_size, device, group, None, comm)
        out = {}
        for row, i in zip(table, sel):
            lk = links[loops[i].source]
            T_other = from_colmajor(row[:16])
            if stage == 0:
                dt, da = identity_check_prev(T_other, loops[i].relative_pose, lk.rel_pose_to_prev)
            else:
                dt, da = identity_check_next(T_other, loops[i].relative_pose, lk.rel_pose_from_next)
            ok = not (dt > max_delta_trans or da > max_delta_angle)
            details[i][key] = {"delta_trans": dt, "delta_angle": da, "consistent": ok, "keyframe": getattr(lk, key)}
            out[i] = ok
        return out

    ok_prev = run(0, todo)
    for i, ok in ok_prev.items():
        passed[i] = ok
    ok_next = run(1, [i for i in todo if not passed[i]])
    for i, ok in ok_next.items():
        passed[i] = ok
    return passed, details


def match_keyframes(reg, clouds, pairs, guesses, estimates, links, fitness_score_max_range=DBL_MAX, fitness_score_thresh=1.25,
                    enable_loop_closure_consistency_check=True, max_delta_trans=0.3, max_delta_angle=0.0523599,
                    use_planar_registration_guess=False, rank=0, world_size=1, device=None, group=None, pair_weights=None, comm=None):
    """LoopDetector::matching (:97-180) for many new keyframes at once: candidate batch, best-candidate rule, consistency
    check, acceptance.  Returns (accepted loops, all loops, consistency details, result table)."""
    loops, table = detect_loops(reg, clouds, pairs, guesses, fitness_score_max_range, fitness_score_thresh, rank, world_size, device,
                                group, pair_weights, comm)
    passed, details = check_consistency(reg, clouds, loops, estimates, links, max_delta_trans, max_delta_angle,
                                        use_planar_registration_guess, enable_loop_closure_consistency_check, rank, world_size, device, group,
                                        comm)
    accepted = []
    for lp, ok in zip(loops, passed):
        if lp.best_candidate is None:
            continue  # :156-160 loop not found
        if enable_loop_closure_consistency_check and not links[lp.source].first_keyframe and not ok:
            continue  # :162-166
        accepted.append(lp)
    return accepted, loops, details, table


# ------------------------------------------------------------------------------------------------ information matrix of an edge
def information_weight(a, max_x, min_y, max_y, x):
    """InformationMatrixCalculator::weight (information_matrix_calculator.cpp:83-88)."""
    y = (1.0 - np.exp(-a * x)) / (1.0 - np.exp(-a * max_x))
    return min_y + (max_y - min_y) * y


def calc_information_matrix(reg, cloud1, cloud2, relpose, use_const_inf_matrix=False, const_stddev_x=0.5, const_stddev_q=0.1, var_gain_a=2.0,
                            min_stddev_x=0.1, max_stddev_x=0.75, min_stddev_q=0.05, max_stddev_q=0.2, fitness_score_thresh=1.25):
    """InformationMatrixCalculator::calc_information_matrix (information_matrix_calculator.cpp:14-44; defaults = config/mrg_slam.yaml:216-223,
    :173): the 6x6 information matrix of an odometry / loop edge from the fitness score of cloud2 (moved by relpose) against cloud1.
    `reg` is anything with fitness_pair(target, source, T) — the GPU engine (one b2r_fitness_pair call on retained clouds) or the oracle."""
    inf = np.eye(6)
    if use_const_inf_matrix:
        inf[:3, :3] /= const_stddev_x
        inf[3:, 3:] /= const_stddev_q
        return inf
    fitness = reg.fitness_pair(cloud1, cloud2, np.asarray(relpose, dtype=np.float64))
    w_x = information_weight(var_gain_a, fitness_score_thresh, min_stddev_x ** 2, max_stddev_x ** 2, fitness)
    w_q = information_weight(var_gain_a, fitness_score_thresh, min_stddev_q ** 2, max_stddev_q ** 2, fitness)
    inf[:3, :3] /= w_x
    inf[3:, 3:] /= w_q
    return inf
