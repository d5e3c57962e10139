#!/bin/bash
# stream launch-shape variants (HVI_STREAM_VARIANT): bench lines only
mkdir -p gpurun_out
for v in "$@"; do
  HVI_STREAM_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2h_bench_v$v.log 2>&1
done
echo done
