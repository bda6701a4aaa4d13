#!/usr/bin/env python
"""Where one headline step (bench.py, N = 1) spends its time: host phases (wall clock) and device stages (CUDA events)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mrg_slam_b200 import lib as B
from mrg_slam_b200 import loop_closure as LC

method = sys.argv[1] if len(sys.argv) > 1 else "FAST_VGICP"
reg = B.Registration(B.default_config(getattr(B, method)))
n_targets, n_cand = (int(sys.argv[2]) if len(sys.argv) > 2 else 256), 16
pool_np, poses = bench.keyframe_pool(reg.prefilter, lambda c, leaf: reg.voxelgrid(c, leaf)[0], n_targets + n_cand)
pairs, guesses = bench.batch_pairs(n_targets, n_cand, poses)
ids = np.array([p[0] for p in pairs], dtype=np.int64)
weights = np.array([len(pool_np[c]) for _, c in pairs], dtype=np.float64)
needed = list(range(len(pool_np)))
dev = torch.device("cuda", 0)
dev_bufs = {c: torch.from_numpy(pool_np[c]).to(dev) for c in needed}
comm = LC.make_comm(reg, 0, 1, nccl=True)
rank_of = B.partition_by_target(ids, 1, weights)
acc = {}
def tick(name, t0):
    t = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t - t0) * 1e3; return t
for it in range(8):
    if it == 3: acc.clear()
    torch.cuda.synchronize()
    t = time.perf_counter(); t_all = t
    cl = B.create_clouds(reg, [dev_bufs[c].data_ptr() for c in needed], [dev_bufs[c].shape[0] for c in needed], B.DEVICE)
    t = tick("create_clouds", t)
    byid = dict(zip(needed, cl))
    src = [byid.get(p[1]) for p in pairs]; tgt = [byid.get(p[0]) for p in pairs]
    t = tick("py_lists", t)
    table = reg.align_batch_sharded(comm, src, tgt, ids, guesses, weights=weights, with_fitness=True)
    t = tick("align_batch_sharded", t)
    for c in cl: c.close()
    t = tick("close", t)
    torch.cuda.synchronize()
    t = tick("final_sync", t)
    acc["total"] = acc.get("total", 0.0) + (t - t_all) * 1e3
    st = reg.last_timings()
    for k, v in st.items(): acc["dev_" + k] = acc.get("dev_" + k, 0.0) + v
print(method, {k: round(v / 5, 3) for k, v in acc.items()})
