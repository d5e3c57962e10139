#!/bin/bash
# ncu --set full of one kernel (regex $1) in the C3' workload; report name $2
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -c 1 -s ${3:-2} -f -o gpurun_out/$2 python bench.py --sites 1000000 --n-mc 32 --scenario covarK2 --nan-frac 0.1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/$2.log 2>&1
echo done
