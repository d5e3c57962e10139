#!/bin/bash
# quick bench lines (default + C3') with the current library; $1 = tag
mkdir -p gpurun_out
timeout 300 python profiles/scripts/tc_check.py 20000 8 > gpurun_out/r2m_check_$1.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2m_bench_default_$1.log 2>&1
timeout 300 python bench.py --sites 1000000 --n-mc 32 --scenario covarK2 --nan-frac 0.1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2m_bench_c3p_$1.log 2>&1
echo done
