#!/bin/bash
# round-2: k_stream with cross-tile prefetch + two-draw observation loop: parity, then variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
for v in 0 4 2 3; do
  HVI_STREAM_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_v$v.log 2>&1
done
echo done
