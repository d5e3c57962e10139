#!/bin/bash
# ncu --set full of the minibatch gather kernel and of the streaming kernel with in-kernel draws (device-dataset leg of bench.py)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gather_preprocess -c 1 -s 1 -f -o gpurun_out/r2ap_gather python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2ap_gather.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_stream.*1, 0, 1>" -c 1 -s 1 -f -o gpurun_out/r2ap_stream_rng python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2ap_stream_rng.log 2>&1
echo done
