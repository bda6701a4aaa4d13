#!/usr/bin/env python
"""One prefilter call per sensor size (after warm-up) between cudaProfilerStart/Stop, for an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file X.csv python tools/prefilter_probe.py [1M]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mrg_slam_b200 import lib as B, synth
reg = B.Registration(B.default_config(B.FAST_VGICP))
sensor = synth.OS1_128_1M if len(sys.argv) > 1 and sys.argv[1] == "1M" else synth.HDL64
raw = synth.scan(sensor, 7)
for _ in range(3):
    reg.prefilter(raw)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = reg.prefilter(raw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(len(raw), len(out))
