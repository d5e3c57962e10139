#!/bin/bash
# config 5 (4 x 512 hidden, covarK2): parity tests of the wide chain, then bench lines; $1 = tag
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_bf3.py tests/test_dense_tc.py -q -m gpu -x > gpurun_out/r2q_pytest_$1.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity_fullsize.py -q -m gpu -x -k C5 >> gpurun_out/r2q_pytest_$1.log 2>&1
timeout 300 python bench.py --sites 100000 --n-mc 16 --scenario covarK2 --hidden 512,512,512,512 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2q_bench_c5small_$1.log 2>&1
timeout 600 python bench.py --sites 1000000 --n-mc 64 --scenario covarK2 --hidden 512,512,512,512 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2q_bench_c5_$1.log 2>&1
echo done
