#!/bin/bash
# Builds libb2r.so with extra nvcc flags into mrg_slam_b200/csrc/build/var_<name>/libb2r.so (kernel A/B experiments; select with
# B2R_LIB_PATH).   tools/build_variant.sh <name> "<flags>" [files to recompile with the flags, default: all]
set -e
NAME=$1; FLAGS=$2; shift 2
cd "$(dirname "$0")/../mrg_slam_b200/csrc"
OUT=build/var_$NAME; mkdir -p $OUT
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="$ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcudafe --diag_suppress=550"
ALL="api cloud lsq ndt filters gicp_pcl loop shard"
SEL=${@:-$ALL}
make -j8 >/dev/null
pids=()
for f in $ALL; do
  if [[ " $SEL " == *" $f "* ]]; then nvcc $COMMON $FLAGS -c $f.cu -o $OUT/$f.o & pids+=($!); else cp build/$f.o $OUT/$f.o; fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc $ARCH -shared -o $OUT/libb2r.so $OUT/*.o -lcudart -ldl
echo $OUT/libb2r.so
