#!/bin/bash
# cp.async-ring fused last-layer backward: tests, ncu --set full of the kernel (262144 rows), config 5 full
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_bf3.py tests/test_dense_tc.py -q -m gpu -x > gpurun_out/r2ah_pytest.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity_fullsize.py -q -m gpu -x -k "C5 or blocked or fused" >> gpurun_out/r2ah_pytest.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_last_bwd_fused -c 1 -s 4 -f -o gpurun_out/r2ah_last_fused python bench.py --sites 262144 --n-mc 4 --scenario covarK2 --hidden 512,512,512,512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2ah_last_fused.log 2>&1
timeout 900 python bench.py --sites 1000000 --n-mc 64 --scenario covarK2 --hidden 512,512,512,512 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2ah_c5.log 2>&1
echo done
