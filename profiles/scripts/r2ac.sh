#!/bin/bash
# A/B on one box: config 5 full with the fused last-layer backward kernel and without it
mkdir -p gpurun_out
for i in 1 2; do
HVI_WIDE_LAST_UNFUSED=1 timeout 900 python bench.py --sites 1000000 --n-mc 64 --scenario covarK2 --hidden 512,512,512,512 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2ac_c5_unfused_$i.log 2>&1
timeout 900 python bench.py --sites 1000000 --n-mc 64 --scenario covarK2 --hidden 512,512,512,512 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2ac_c5_fused_$i.log 2>&1
done
echo done
