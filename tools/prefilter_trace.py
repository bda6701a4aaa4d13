#!/usr/bin/env python
"""Stage times of b2r_prefilter (B2R_TRACE=1) for pageable and pinned host input."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mrg_slam_b200 import lib as B, synth
reg = B.Registration(B.default_config(B.FAST_VGICP))
raw = synth.scan(synth.HDL64, 100)
pinned = torch.from_numpy(raw).pin_memory().numpy()
for name, a in (("pageable", raw), ("pinned", pinned)):
    for _ in range(5): reg.prefilter(a)
    t = []
    for _ in range(30):
        t0 = time.perf_counter(); out = reg.prefilter(a); t.append(1e3 * (time.perf_counter() - t0))
    print(name, "n_in", len(a), "n_out", len(out), "ms p50", round(float(np.median(t)), 4), flush=True)
