#!/bin/bash
# Build an A/B variant of libgtop_b200.so with extra -D flags for the evaluation kernels:
#   tools/build_variant.sh <tag> [-DGTOP_MMA_SPL=1 -DGTOP_MMA_MINB=5 ...]   ->  build/variants/libgtop_<tag>.so
# Run it with GTOP_B200_LIB=build/variants/libgtop_<tag>.so python tools/kbench.py
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/grad_traj_optimization_b200/csrc
TAG=$1; shift
OUT=$ROOT/build/variants; mkdir -p $OUT/$TAG
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -ffp-contract=off"
for f in eval_kernel eval_mma_kernel; do
  nvcc $FLAGS "$@" -Xptxas -v -c $SRC/$f.cu -o $OUT/$TAG/$f.o 2> $OUT/$TAG/$f.ptxas.log
done
grep -h "Used\|spill" $OUT/$TAG/eval_mma_kernel.ptxas.log | tr '\n' ' '; echo
nvcc $ARCH -shared -o $OUT/libgtop_$TAG.so $OUT/$TAG/eval_kernel.o $OUT/$TAG/eval_mma_kernel.o $SRC/eval_one_kernel.o $SRC/sdf_kernels.o $SRC/solve_kernels.o $SRC/setup_kernels.o $SRC/traj_metrics_kernels.o $SRC/gtop_capi.o $SRC/problem_setup.o -cudart static
