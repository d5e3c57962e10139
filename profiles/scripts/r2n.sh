#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --sites 1000000 --n-mc 32 --scenario covarK2 --nan-frac 0.1 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2n_dbg.log 2>&1
echo done
