#!/bin/bash
# tanh with one reciprocal per pair in the tcgen05 forward kernel: tests + A/B (C3' and default workloads)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_tc.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r2aj_pytest.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity_fullsize.py -q -m gpu -x -k "C2 or C3 or C4" >> gpurun_out/r2aj_pytest.log 2>&1
for v in 0 1; do
HVI_TANH_SHARED=$v timeout 300 python bench.py --sites 1000000 --n-mc 32 --scenario covarK2 --nan-frac 0.1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2aj_c3p_$v.log 2>&1
HVI_TANH_SHARED=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2aj_default_$v.log 2>&1
done
echo done
