"""world_size-2 gloo test of the loop-closure sharding path (CPU): partition by target, per-rank work, all-gather,
rank-agnostic best-candidate reduction.  The registration itself is replaced by a deterministic stand-in so the
test exercises exactly the host-side N>1 logic (the GPU test runs the real thing)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class _FakeResult:
    def __init__(self, pair):
        rng = np.random.default_rng(1000 + pair)
        self.T = rng.normal(size=16).astype(np.float32)
        self.converged = int(pair % 7 != 3)
        self.iterations = pair % 5
        self.error = float(pair) * 0.5
        self.evals = 2 + pair % 3
        self.fitness = float(rng.integers(0, 4)) * 0.25  # coarse values => ties occur


class _FakeReg:
    """align_batch stand-in: result depends only on the global pair index carried in the guess."""

    def align_batch(self, sources, targets, guesses, with_fitness=False, fitness_max_range=0.0):
        return [_FakeResult(int(g[0, 3])) for g in guesses]


def _worker(rank, world_size, port, out_dir):
    from mrg_slam_b200 import loop_closure as LC

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    n_targets, k = 9, 5
    pairs = [(t, n_targets + t * k + c) for t in range(n_targets) for c in range(k)]
    guesses = []
    for i in range(len(pairs)):
        g = np.eye(4)
        g[0, 3] = i
        guesses.append(g)
    clouds = [None] * (n_targets + n_targets * k)
    loops, table = LC.detect_loops(_FakeReg(), clouds, pairs, guesses, fitness_score_thresh=0.5, rank=rank, world_size=world_size)
    np.save(os.path.join(out_dir, f"table_{rank}.npy"), table)
    np.save(os.path.join(out_dir, f"loops_{rank}.npy"), np.array([[l.target, -1 if l.best_candidate is None else l.best_candidate, l.best_score]
                                                                  for l in loops]))
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_rank(tmp_path):
    from mrg_slam_b200 import loop_closure as LC

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    t0, t1 = np.load(tmp_path / "table_0.npy"), np.load(tmp_path / "table_1.npy")
    l0, l1 = np.load(tmp_path / "loops_0.npy"), np.load(tmp_path / "loops_1.npy")
    assert np.array_equal(t0, t1) and np.array_equal(l0, l1)  # every rank ends with the same full table and decisions
    # single-process reference
    n_targets, k = 9, 5
    expect = LC.pack_results([_FakeResult(i) for i in range(n_targets * k)], list(range(n_targets * k)))
    assert np.array_equal(t0, expect)
    for t in range(n_targets):
        rows = expect[t * k:(t + 1) * k]
        best, score = LC.select_best(rows[:, 20], rows[:, 16] != 0)
        want = -1 if (best is None or score > 0.5) else best
        assert l0[t, 0] == t and l0[t, 1] == want


class _PoseReg:
    """align_batch stand-in whose result pose depends only on the global pair id carried in the guess: translation
    (0.01 * id, 0, 0), so some of the identity checks of the consistency stage pass and some fail."""

    def align_batch(self, sources, targets, guesses, with_fitness=False, fitness_max_range=0.0):
        out = []
        for g in guesses:
            r = _FakeResult(int(round(g[1, 3])))
            T = np.eye(4, dtype=np.float32)
            T[0, 3] = 0.01 * int(round(g[1, 3]))
            r.T = T.T.reshape(-1).copy()  # column-major
            r.converged = 1
            out.append(r)
        return out


def _consistency_case():
    from mrg_slam_b200 import loop_closure as LC

    n_targets, k = 8, 3
    ncl = n_targets + n_targets * k + 2 * n_targets * k
    pairs = [(t, n_targets + t * k + c) for t in range(n_targets) for c in range(k)]
    est, links = {}, {}
    for i in range(ncl):
        est[i] = np.eye(4)
        est[i][1, 3] = float(i)  # registration_guess(new, other)[1, 3] = other - new: _PoseReg reads the pair id from it
        links[i] = LC.KeyframeLinks()
    guesses = [LC.registration_guess(est[t], est[s]).astype(np.float64) for t, s in pairs]
    base = n_targets + n_targets * k
    for j, (t, s) in enumerate(pairs):
        rel = np.eye(4)
        rel[0, 3] = 0.2 if j % 2 else -5.0  # every other candidate has a prev edge that cannot close
        links[s] = LC.KeyframeLinks(prev=base + 2 * j, rel_pose_to_prev=rel, next=base + 2 * j + 1, rel_pose_from_next=np.eye(4),
                                    first_keyframe=(j % 7 == 0))
    return pairs, guesses, est, links, ncl


def _consistency_worker(rank, world_size, port, out_dir):
    from mrg_slam_b200 import loop_closure as LC

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    pairs, guesses, est, links, ncl = _consistency_case()
    accepted, loops, details, table = LC.match_keyframes(_PoseReg(), [None] * ncl, pairs, guesses, est, links, fitness_score_thresh=10.0,
                                                         rank=rank, world_size=world_size)
    np.save(os.path.join(out_dir, f"acc_{rank}.npy"), np.array([[l.target, l.source, l.best_score] for l in accepted]))
    np.save(os.path.join(out_dir, f"det_{rank}.npy"), np.array([[d.get("prev", {}).get("delta_trans", -1.0), d.get("next", {}).get("delta_trans", -1.0)]
                                                                 for d in details]))
    dist.destroy_process_group()


def test_two_rank_consistency_check_matches_single_rank(tmp_path):
    from mrg_slam_b200 import loop_closure as LC

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_consistency_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a0, a1 = np.load(tmp_path / "acc_0.npy"), np.load(tmp_path / "acc_1.npy")
    d0, d1 = np.load(tmp_path / "det_0.npy"), np.load(tmp_path / "det_1.npy")
    assert np.array_equal(a0, a1) and np.array_equal(d0, d1)
    pairs, guesses, est, links, ncl = _consistency_case()
    accepted, loops, details, table = LC.match_keyframes(_PoseReg(), [None] * ncl, pairs, guesses, est, links, fitness_score_thresh=10.0)
    want = np.array([[l.target, l.source, l.best_score] for l in accepted])
    assert np.array_equal(a0, want)
    assert 0 < len(accepted) <= len(loops)
    # both stages ran: some loops were settled by `prev`, some needed `next`
    assert (d0[:, 0] >= 0).any() and (d0[:, 1] >= 0).any()
