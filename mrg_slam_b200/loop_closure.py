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


def partition_by_target(target_ids: Sequence[int], world_size: int) -> List[List[int]]:
    """Static block partition of pair indices by target id, balanced by pair count.

    Targets keep their order of first appearance; each rank gets a contiguous block of targets.  Every pair of
    a given target lands on exactly one rank.
    """
    order, groups = [], {}
    for i, t in enumerate(target_ids):
        if t not in groups:
            groups[t] = []
            order.append(t)
        groups[t].append(i)
    total = len(target_ids)
    shards: List[List[int]] = [[] for _ in range(world_size)]
    rank, acc = 0, 0
    for t in order:
        # move to the next rank when this rank already holds its share (never leave later ranks without a chance)
        while rank < world_size - 1 and acc >= (rank + 1) * total / world_size:
            rank += 1
        shards[rank].extend(groups[t])
        acc += len(groups[t])
    return shards


def select_best(scores: Sequence[float], converged: Sequence[bool]):
    """The candidate reduction of loop_detector.cpp:106-145.  Returns (best_index or None, best_score)."""
    best_score, best = DBL_MAX, None
    for i, (s, c) in enumerate(zip(scores, converged)):
        if (not c) or s > best_score:
            continue
        best_score, best = s, i
    return best, best_score


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


@dataclass
class Loop:
    target: int
    best_candidate: Optional[int]  # index into that target's candidate list
    best_score: float
    relative_pose: Optional[np.ndarray]  # new keyframe <- best candidate, 4x4


def detect_loops(reg, clouds, pairs, guesses, fitness_score_max_range=DBL_MAX, fitness_score_thresh=1.25, rank=0, world_size=1,
                 device=None, group=None):
    """Batched LoopDetector::matching over many new keyframes.

    clouds: list of mrg_slam_b200.lib.Cloud (only those this rank needs may be non-None)
    pairs:  list of (target_cloud_index, source_cloud_index), candidates of one target contiguous and in
            candidate order (the tie rule depends on it)
    Returns (loops per target in order of first appearance, full result table).
    """
    from .lib import from_colmajor

    target_ids = [p[0] for p in pairs]
    shards = partition_by_target(target_ids, world_size)
    mine = shards[rank]
    res = reg.align_batch([clouds[pairs[i][1]] for i in mine], [clouds[pairs[i][0]] for i in mine], [guesses[i] for i in mine],
                          with_fitness=True, fitness_max_range=fitness_score_max_range) if mine else []
    table = gather_results(pack_results(res, mine), len(pairs), device=device, group=group)
    loops, seen = [], {}
    for i, t in enumerate(target_ids):
        seen.setdefault(t, []).append(i)
    for t, idxs in seen.items():
        best, score = select_best(table[idxs, 20], table[idxs, 16] != 0)
        if best is None or score > fitness_score_thresh:
            loops.append(Loop(t, None, score, None))
        else:
            loops.append(Loop(t, best, score, from_colmajor(table[idxs[best], :16])))
    return loops, table
