#!/bin/bash
# tcgen05 small applicator: correctness against the mma.sync and Float64 paths, then bench lines
mkdir -p gpurun_out
timeout 300 python profiles/scripts/tc_check.py 1000 4 > gpurun_out/r2k_check_small.log 2>&1
timeout 300 python profiles/scripts/tc_check.py 200000 32 > gpurun_out/r2k_check.log 2>&1
for mode in 0 2; do
HVI_MLP_TC=$mode timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2k_bench_default_tc$mode.log 2>&1
HVI_MLP_TC=$mode timeout 300 python bench.py --sites 1000000 --n-mc 32 --scenario covarK2 --nan-frac 0.1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2k_bench_c3p_tc$mode.log 2>&1
done
echo done
