#!/bin/bash
# ncu --set full of the fused last-layer backward kernel and of one dgrad GEMM with the column-sum epilogue (config-5 shape)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_last_bwd_fused -c 1 -s 4 -f -o gpurun_out/r2af_last_fused python bench.py --sites 262144 --n-mc 4 --scenario covarK2 --hidden 512,512,512,512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2af_last_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_bf3<256, 1>" -c 1 -s 4 -f -o gpurun_out/r2af_dgrad python bench.py --sites 262144 --n-mc 4 --scenario covarK2 --hidden 512,512,512,512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2af_dgrad.log 2>&1
echo done
