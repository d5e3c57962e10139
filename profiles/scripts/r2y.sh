#!/bin/bash
# site-blocked wide evaluation: tests of the wide chain + config 5 (small and full) bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_bf3.py tests/test_dense_tc.py -q -m gpu -x > gpurun_out/r2y_pytest.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity_fullsize.py -q -m gpu -x -k "C5 or blocked" >> gpurun_out/r2y_pytest.log 2>&1
timeout 300 python bench.py --sites 100000 --n-mc 16 --scenario covarK2 --hidden 512,512,512,512 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2y_bench_c5small.log 2>&1
HVI_WIDE_BLOCK_SITES=32768 timeout 300 python bench.py --sites 100000 --n-mc 16 --scenario covarK2 --hidden 512,512,512,512 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2y_bench_c5small_blk32k.log 2>&1
timeout 900 python bench.py --sites 1000000 --n-mc 64 --scenario covarK2 --hidden 512,512,512,512 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2y_bench_c5.log 2>&1
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> gpurun_out/r2y_bench_c5.log
echo done
