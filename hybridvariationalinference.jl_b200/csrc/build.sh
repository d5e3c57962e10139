#!/bin/bash
# Builds libhvi_b200.so (sm_100a only) in-tree.  Usage: ./build.sh [extra nvcc flags]
# HVI_REBUILD=1 forces every object to be recompiled.  A failed compile fails the script (and never leaves a
# stale object behind for the link step).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
SRCS="hvi_kernels hvi_stream hvi_capi hvi_dense_tc hvi_dense_tma hvi_mlp_mma hvi_mlp_tc hvi_point hvi_predict hvi_generic"
HDRS="hvi_internal.cuh hvi_dev.cuh hvi_stream.cuh hvi_tc_common.cuh hvi_rng.cuh hvi_pbm.cuh ../../include/hvi_b200.h"
# the text of hvi_pbm.cuh travels inside the library: NVRTC compiles it together with a user's process model
if [ ! -f hvi_pbm_src.inc ] || [ hvi_pbm.cuh -nt hvi_pbm_src.inc ]; then
  { printf 'static const char hvi_pbm_header[] = R"HVIPBMSRC('; cat hvi_pbm.cuh; printf ')HVIPBMSRC";\n'; } > hvi_pbm_src.inc
fi
pids=()
for f in $SRCS; do
  stale=0
  if [ -n "$HVI_REBUILD" ] || [ ! -f $f.o ] || [ $f.cu -nt $f.o ]; then stale=1; fi
  for h in $HDRS; do if [ -f $h ] && [ $h -nt $f.o ]; then stale=1; fi; done
  if [ $stale = 1 ]; then
    rm -f $f.o
    $NVCC $FLAGS "$@" -c $f.cu -o $f.o &
    pids+=($!)
  fi
done
for pid in "${pids[@]}"; do wait $pid || { echo "build.sh: a compile failed" >&2; exit 1; }; done
OBJS=""
for f in $SRCS; do OBJS="$OBJS $f.o"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libhvi_b200.so $OBJS -ldl
