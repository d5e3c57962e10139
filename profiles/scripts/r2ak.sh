#!/bin/bash
# tanh: one reciprocal per four (HVI_TANH_SHARED=2) against per pair (1): tests + A/B
mkdir -p gpurun_out
HVI_TANH_SHARED=2 timeout 600 python -m pytest tests/test_mlp_tc.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r2ak_pytest.log 2>&1
HVI_TANH_SHARED=2 timeout 600 python -m pytest tests/test_gpu_parity_fullsize.py -q -m gpu -x -k "C2 or C3 or C4" >> gpurun_out/r2ak_pytest.log 2>&1
for v in 1 2 1 2; do
HVI_TANH_SHARED=$v timeout 300 python bench.py --sites 1000000 --n-mc 32 --scenario covarK2 --nan-frac 0.1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/r2ak_c3p_$v.log 2>&1
done
HVI_TANH_SHARED=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ak_default_2.log 2>&1
echo done
