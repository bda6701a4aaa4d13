set -u
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_v7_final.jsonl 2> gpurun_out/bench_v7_final.err; tail -c 300 gpurun_out/bench_v7_final.jsonl
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_v7_reference_arm.jsonl 2>> gpurun_out/bench_v7_final.err
KERNELS="knn_cov vgicp_eval" bash tools/ncu_bench.sh r1v7
mkdir -p gpurun_out/configs_v7
timeout 200 python bench_configs.py --config ndt_vlp16 --cpu > gpurun_out/configs_v7/ndt_vlp16.jsonl 2>> gpurun_out/cfg.err
timeout 300 python bench_configs.py --config odometry --cpu > gpurun_out/configs_v7/odometry.jsonl 2>> gpurun_out/cfg.err
timeout 200 python bench_configs.py --config prefilter --cpu > gpurun_out/configs_v7/prefilter.jsonl 2>> gpurun_out/cfg.err
for M in FAST_GICP FAST_VGICP NDT_OMP; do timeout 200 python bench_configs.py --config loop_closure --method $M --steps 3 --warmup 1 --cpu > gpurun_out/configs_v7/lc_$M.jsonl 2>> gpurun_out/cfg.err; done
timeout 200 python bench_configs.py --config submap --steps 3 --warmup 1 > gpurun_out/configs_v7/submap.jsonl 2>> gpurun_out/cfg.err
tail -3 gpurun_out/cfg.err
ls gpurun_out/configs_v7
