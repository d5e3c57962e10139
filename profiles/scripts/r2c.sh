#!/bin/bash
# round-2 third GPU pass: single-launch small-batch iteration (k_small_iter)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
timeout 400 python profiles/bench_small_batch.py > gpurun_out/r2c_small.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c_bench.log 2>&1
echo done
