#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_adam.py tests/test_dataset.py -q -m gpu -x > gpurun_out/r2t_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2t_pytest.log
timeout 600 python profiles/bench_small_batch.py > gpurun_out/r2t_small.log 2>&1
echo done
