#!/bin/bash
# Profiles of the bench workload itself (run under gpurun, one GPU).  bench.py brackets its timed steps with
# cudaProfilerStart/Stop, so `--profile-from-start off` sees exactly the kernels of the timed region.
#   tools/ncu_bench.sh <tag> [bench args...]       e.g. tools/ncu_bench.sh r1_vgicp --method FAST_VGICP
#   KERNELS="knn_cov lsq_eval" selects the kernels captured with --set full (one launch each).
# The optimiser loop runs host-polled here (B2R_GRAPH_LOOP=0) so that ncu sees plain kernel launches; the kernels are the same.
set -u
export B2R_GRAPH_LOOP=0
TAG=${1:?tag}; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras "$@" > gpurun_out/launches_$TAG.log 2>&1 || tail -5 gpurun_out/launches_$TAG.log
for K in ${KERNELS:-knn_cov lsq_eval}; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$K -c 1 -f -o gpurun_out/ncu_${TAG}_$K \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras "$@" > gpurun_out/ncu_${TAG}_$K.log 2>&1 || tail -5 gpurun_out/ncu_${TAG}_$K.log
done
ls -la gpurun_out | tail -12
