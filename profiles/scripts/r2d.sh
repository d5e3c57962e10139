#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_adam.py tests/test_dataset.py tests/test_prefetch.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
timeout 400 python profiles/bench_small_batch.py > gpurun_out/r2d_small.log 2>&1
echo done
