#!/usr/bin/env python
"""Where one odometry scan's ~1.3 ms goes (run under gpurun): host wall clock of each call of the serial loop next to the device
time of the align (b2r_last_timings) — separates Python / ctypes overhead, host work inside the C calls and GPU time."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mrg_slam_b200 import lib as B  # noqa: E402
from mrg_slam_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    reg = B.Registration(B.default_config(B.FAST_VGICP))
    raws = [synth.scan(synth.HDL64, 100 + i) for i in range(n)]
    seg = {k: [] for k in ("prefilter", "cloud", "set_source", "align", "result", "dev_prep", "dev_opt", "dev_total")}
    tgt = B.Cloud(reg, reg.prefilter(raws[0]))
    reg.setInputTarget(tgt)
    for i in range(1, n):
        t0 = time.perf_counter()
        pts = reg.prefilter(raws[i])
        t1 = time.perf_counter()
        c = B.Cloud(reg, pts)
        t2 = time.perf_counter()
        reg.setInputSource(c)
        t3 = time.perf_counter()
        reg.align(np.eye(4))
        t4 = time.perf_counter()
        T = reg.getFinalTransformation()
        ok = reg.hasConverged()
        t5 = time.perf_counter()
        lt = reg.last_timings()
        if i > 5:
            for k, v in zip(("prefilter", "cloud", "set_source", "align", "result"), (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
                seg[k].append(1e3 * v)
            vals = list(lt.values()) if isinstance(lt, dict) else list(lt)
            seg["dev_prep"].append(vals[0]); seg["dev_opt"].append(vals[1]); seg["dev_total"].append(vals[3])
        # every scan becomes the next target (keyframe switch every scan: the worst case for structure builds)
        reg.setInputTarget(c)
        tgt.close()
        tgt = c
        assert ok and T is not None
    print({k: round(float(np.median(v)), 4) for k, v in seg.items()})


if __name__ == "__main__":
    main()
