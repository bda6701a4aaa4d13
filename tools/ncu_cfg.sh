#!/bin/bash
# ncu --set full capture of one kernel inside a bench_configs.py run (run under gpurun, one GPU).
#   tools/ncu_cfg.sh <tag> <kernel regex> <launch skip> <bench_configs args...>
set -u
TAG=${1:?tag}; K=${2:?kernel regex}; SKIP=${3:?skip}; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/ncu_$TAG \
    python bench_configs.py --steps 1 --warmup 1 "$@" > gpurun_out/ncu_$TAG.log 2>&1 || tail -5 gpurun_out/ncu_$TAG.log
