#!/bin/bash
# fused last-layer backward of the wide chain: tests + config-5 bench lines (fused / unfused); $1 = tag
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_bf3.py tests/test_dense_tc.py -q -m gpu -x > gpurun_out/$1_pytest.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity_fullsize.py tests/test_gpu_parity.py -q -m gpu -x -k "C5 or blocked or wide or hidden" >> gpurun_out/$1_pytest.log 2>&1
timeout 300 python bench.py --sites 100000 --n-mc 16 --scenario covarK2 --hidden 512,512,512,512 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/$1_bench_c5small.log 2>&1
HVI_WIDE_LAST_UNFUSED=1 timeout 300 python bench.py --sites 100000 --n-mc 16 --scenario covarK2 --hidden 512,512,512,512 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/$1_bench_c5small_unfused.log 2>&1
timeout 900 python bench.py --sites 1000000 --n-mc 64 --scenario covarK2 --hidden 512,512,512,512 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/$1_bench_c5.log 2>&1
echo done
