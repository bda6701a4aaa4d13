import sys; sys.path.insert(0,".")
import numpy as np
from mrg_slam_b200 import lib as B, synth
reg = B.Registration(B.default_config(B.FAST_VGICP))
scans=[reg.prefilter(synth.scan(synth.HDL64, 100+i)) for i in range(9)]
cl=[B.Cloud(reg,s) for s in scans]
reg.align_batch(cl[1:],cl[:-1],[np.eye(4)]*8)
print("queries", sum(len(s) for s in scans), "overflows", reg.knn_list_overflows())
