#!/bin/bash
# BASELINE configs 2, 3, 4-shard with the final library (device-timed lines), tc on and off at the small sizes
mkdir -p gpurun_out
for tc in 2 0; do
HVI_MLP_TC=$tc timeout 300 python bench.py --sites 100000 --n-mc 32 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2u_c2_tc$tc.log 2>&1
HVI_MLP_TC=$tc timeout 300 python bench.py --sites 10000 --n-mc 32 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2u_1e4_tc$tc.log 2>&1
done
timeout 300 python bench.py --sites 1000000 --n-mc 32 --nan-frac 0.1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_c3.log 2>&1
timeout 300 python bench.py --sites 1250000 --n-mc 64 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_c4shard.log 2>&1
timeout 300 python bench.py --sites 100000 --n-mc 32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_c2_e2e.log 2>&1
echo done
