#!/bin/bash
# round-2 second GPU pass: GPU suite (generic PBM seam, sqrt(w)-prescaled obs loop), bench, launch-shape variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_bench.log
for v in 2 3; do
  HVI_STREAM_VARIANT=$v timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_v$v.log 2>&1
done
echo done
