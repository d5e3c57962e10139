#!/bin/bash
mkdir -p gpurun_out
HVI_STREAM_VARIANT=${1:-0} timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_stream -c 1 -s 2 -f -o gpurun_out/r2g_k_stream_v${1:-0} python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2g_ncu.log 2>&1
echo done
