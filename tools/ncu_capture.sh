#!/bin/bash
# ncu --set full capture of one kernel family under tests/prof_driver.py (run under gpurun, one GPU).
#   tools/ncu_capture.sh <kernel regex> <out name> [method] [pairs] [skip]
set -e
K=${1:?kernel regex}; OUT=${2:?output name}; METHOD=${3:-FAST_VGICP}; PAIRS=${4:-8}; SKIP=${5:-0}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/$OUT \
    python tests/prof_driver.py $METHOD $PAIRS 1 > gpurun_out/$OUT.log 2>&1 || { tail -20 gpurun_out/$OUT.log; exit 1; }
tail -2 gpurun_out/$OUT.log
