#!/bin/bash
# Builds libhvi_b200.so (sm_100a only) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math=false -Xcompiler -fPIC -Xcompiler -O2"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
for f in hvi_kernels hvi_capi hvi_dense_tc hvi_dense_tma hvi_mlp_mma hvi_point hvi_predict; do
  if [ ! -f $f.o ] || [ $f.cu -nt $f.o ] || [ hvi_internal.cuh -nt $f.o ] || [ hvi_tc_common.cuh -nt $f.o ] || [ hvi_rng.cuh -nt $f.o ] || [ ../../include/hvi_b200.h -nt $f.o ]; then
    $NVCC $FLAGS "$@" -c $f.cu -o $f.o &
  fi
done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libhvi_b200.so hvi_kernels.o hvi_capi.o hvi_dense_tc.o hvi_dense_tma.o hvi_mlp_mma.o hvi_point.o hvi_predict.o
