"""Where one bench step spends its time: host wall clock per API call vs device stage times (run under gpurun)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from mrg_slam_b200 import lib as B, synth

reg = B.Registration(B.default_config(B.FAST_VGICP))
P = 32
clouds_np = [reg.prefilter(synth.scan(synth.HDL64, 100 + i)) for i in range(P + 1)]
dev = [torch.from_numpy(c).cuda() for c in clouds_np]
pin = [torch.from_numpy(c).pin_memory() for c in clouds_np]
guesses = [np.eye(4)] * P
for name, bufs, space in (("device", dev, B.DEVICE), ("pinned host", pin, B.HOST)):
    rows = []
    for it in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cl = B.create_clouds(reg, [b.data_ptr() for b in bufs], [b.shape[0] for b in bufs], space)
        t1 = time.perf_counter()
        res = reg.align_batch(cl[1:], cl[:-1], guesses)
        t2 = time.perf_counter()
        for c in cl:
            c.close()
        reg.synchronize()
        t3 = time.perf_counter()
        tm = reg.last_timings()
        rows.append([1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), tm["prep_ms"], tm["optimize_ms"], tm["total_ms"]])
    r = np.median(np.array(rows[3:]), axis=0)
    print(f"{name:12s} create_clouds {r[0]:.3f} ms | align_batch {r[1]:.3f} ms (device: prep {r[3]:.3f} optimise {r[4]:.3f} total {r[5]:.3f}) | destroy {r[2]:.3f} ms")
