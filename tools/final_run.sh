#!/bin/bash
# The command list behind the committed profiles/<round>/ files: GPU tests, both bench arms, the ncu launch list and --set full
# captures of the bench step, every BASELINE configuration.   tools/final_run.sh <tag>   (run under gpurun, one GPU)
set -u
TAG=${1:-r2}
O=gpurun_out/final_$TAG; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/gpu_tests.txt
timeout 400 python bench.py > $O/bench.jsonl 2> $O/bench.err; tail -c 300 $O/bench.jsonl
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.jsonl 2>> $O/bench.err
for M in FAST_GICP NDT_OMP SMALL_GICP; do timeout 300 python bench.py --method $M --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_$M.jsonl 2>> $O/bench.err; done
timeout 200 python bench_configs.py --config ndt_vlp16 --cpu > $O/ndt_vlp16.jsonl 2>> $O/cfg.err
timeout 400 python bench_configs.py --config odometry --cpu > $O/odometry.jsonl 2>> $O/cfg.err
timeout 200 python bench_configs.py --config prefilter --cpu > $O/prefilter.jsonl 2>> $O/cfg.err
timeout 300 python bench_configs.py --config submap --steps 3 --warmup 1 > $O/submap.jsonl 2>> $O/cfg.err
tail -3 $O/cfg.err
ls $O
