#!/bin/bash
# bias gradients as column sums in the dgrad epilogue: tests + same-box A/B at config 5 full
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_bf3.py tests/test_dense_tc.py -q -m gpu -x > gpurun_out/r2ae_pytest.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity_fullsize.py -q -m gpu -x -k "C5 or blocked or fused" >> gpurun_out/r2ae_pytest.log 2>&1
for i in 1 2; do
HVI_WIDE_BIAS_SKINNY=1 timeout 900 python bench.py --sites 1000000 --n-mc 64 --scenario covarK2 --hidden 512,512,512,512 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2ae_c5_skinny_$i.log 2>&1
timeout 900 python bench.py --sites 1000000 --n-mc 64 --scenario covarK2 --hidden 512,512,512,512 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2ae_c5_colsum_$i.log 2>&1
done
echo done
