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
