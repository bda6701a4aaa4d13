"""Host logic of the loop-closure batch path: candidate loop, best-score rule, and sharding over ranks.

Mirrors LoopDetector::matching (/root/reference/src/mrg_slam/loop_detector.cpp:97-180): one target (the new
keyframe) against K candidate sources; per candidate align + getFitnessScore, keep the best converged one
(`score > best_score -> skip`, so on equal scores the LATER candidate wins, :138-140); accept iff
best_score <= fitness_score_thresh (:156).  The K aligns are independent, so batches of (target, candidate)
pairs are sharded across GPUs by target id — all candidates of one keyframe on one rank, so its voxel map /
covariances are built once — and the fixed-size results are all-gathered (NCCL on GPUs, gloo in CPU tests).
There is no collective inside the optimiser.
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

DBL_MAX = float(np.finfo(np.float64).max)
RESULT_WIDTH = 24  # T(16) converged iterations error evals fitness pair_index pad pad


def partition_by_target(target_ids: Sequence[int], world_size: int, weights: Optional[Sequence[float]] = None) -> List[List[int]]:
    """Static block partition of pair indices by target id, balanced by pair count or, if given, by per-pair `weights`
    (e.g. the source cloud sizes: an alignment's cost is proportional to the points it evaluates).

    Targets keep their order of first appearance; each rank gets a contiguous block of targets.  Every pair of
    a given target lands on exactly one rank.
    """
    order, groups = [], {}
    for i, t in enumerate(target_ids):
        if t not in groups:
            groups[t] = []
            order.append(t)
        groups[t].append(i)
    w = [1.0] * len(target_ids) if weights is None else [float(x) for x in weights]
    total = sum(w)
    shards: List[List[int]] = [[] for _ in range(world_size)]
    rank, acc = 0, 0.0
    for t in order:
        gw = sum(w[i] for i in groups[t])
        # move on when this rank holds its share (a group goes where most of it falls); never starve the later ranks
        while rank < world_size - 1 and acc + 0.5 * gw >= (rank + 1) * total / world_size:
            rank += 1
        shards[rank].extend(groups[t])
        acc += gw
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


def gather_results(local: np.ndarray, n_pairs: int, device=None, group=None) -> np.ndarray:
    """All-gathers the per-rank result rows into an (n_pairs, RESULT_WIDTH) array ordered by pair index."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        full = np.zeros((n_pairs, RESULT_WIDTH))
        full[local[:, 21].astype(np.int64)] = local
        return full
    ws = dist.get_world_size(group)
    dev = device if device is not None else torch.device("cpu")
    cnt = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(ws)]
    dist.all_gather(cnts, cnt, group=group)
    mx = int(max(int(c.item()) for c in cnts))
    buf = torch.zeros((mx, RESULT_WIDTH), dtype=torch.float64, device=dev)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(local).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(ws)]
    dist.all_gather(bufs, buf, group=group)
    full = np.zeros((n_pairs, RESULT_WIDTH))
    for c, b in zip(cnts, bufs):
        k = int(c.item())
        if k:
            rows = b[:k].cpu().numpy()
            full[rows[:, 21].astype(np.int64)] = rows
    return full


LAST_TIMINGS = {}  # host wall clock of the last detect_loops call on this rank: align_ms (partition + batch), gather_ms


@dataclass
class Loop:
    target: int
    best_candidate: Optional[int]  # index into that target's candidate list
    best_score: float
    relative_pose: Optional[np.ndarray]  # new keyframe <- best candidate, 4x4


def detect_loops(reg, clouds, pairs, guesses, fitness_score_max_range=DBL_MAX, fitness_score_thresh=1.25, rank=0, world_size=1,
                 device=None, group=None, pair_weights=None):
    """Batched LoopDetector::matching over many new keyframes.

    clouds: list of mrg_slam_b200.lib.Cloud (only those this rank needs may be non-None)
    pairs:  list of (target_cloud_index, source_cloud_index), candidates of one target contiguous and in
            candidate order (the tie rule depends on it)
    Returns (loops per target in order of first appearance, full result table).
    """
    import time

    from .lib import from_colmajor

    t_start = time.perf_counter()
    target_ids = [p[0] for p in pairs]
    shards = partition_by_target(target_ids, world_size, pair_weights)  # every rank computes the same partition
    mine = shards[rank]
    if mine and not hasattr(reg, "align_batch_table"):  # any object with the plain align_batch surface
        local = pack_results(reg.align_batch([clouds[pairs[i][1]] for i in mine], [clouds[pairs[i][0]] for i in mine],
                                             [guesses[i] for i in mine], with_fitness=True, fitness_max_range=fitness_score_max_range), mine)
    elif mine:
        res = reg.align_batch_table([clouds[pairs[i][1]] for i in mine], [clouds[pairs[i][0]] for i in mine], [guesses[i] for i in mine],
                                    with_fitness=True, fitness_max_range=fitness_score_max_range)
        local = pack_table(res, mine)
    else:
        local = np.zeros((0, RESULT_WIDTH))
    t_aligned = time.perf_counter()
    table = gather_results(local, len(pairs), device=device, group=group)
    t_gathered = time.perf_counter()
    LAST_TIMINGS.update(align_ms=1e3 * (t_aligned - t_start), gather_ms=1e3 * (t_gathered - t_aligned))
    loops, seen = [], {}
    for i, t in enumerate(target_ids):
        seen.setdefault(t, []).append(i)
    sizes = {len(v) for v in seen.values()}
    contiguous = all(v == list(range(v[0], v[0] + len(v))) for v in seen.values())
    if len(sizes) == 1 and contiguous and list(seen) == sorted(seen, key=lambda t: seen[t][0]):
        # the usual shape (every keyframe has the same number of candidates, stored contiguously): one vectorised reduction
        k = sizes.pop()
        first = np.array([v[0] for v in seen.values()])
        rows = first[:, None] + np.arange(k)[None, :]
        best, score = select_best_grouped(table[rows, 20], table[rows, 16] != 0)
        for (t, idxs), b, sc in zip(seen.items(), best, score):
            if b < 0 or sc > fitness_score_thresh:
                loops.append(Loop(t, None, float(sc), None))
            else:
                loops.append(Loop(t, int(b), float(sc), from_colmajor(table[idxs[int(b)], :16])))
        return loops, table
    for t, idxs in seen.items():
        best, score = select_best(table[idxs, 20], table[idxs, 16] != 0)
        if best is None or score > fitness_score_thresh:
            loops.append(Loop(t, None, score, None))
        else:
            loops.append(Loop(t, best, score, from_colmajor(table[idxs[best], :16])))
    return loops, table
