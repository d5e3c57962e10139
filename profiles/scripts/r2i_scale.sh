#!/bin/bash
# strong scaling of BASELINE config 4 (1e7 sites in total x 64 draws) and the weak-scaling line at N GPUs
N=$1
mkdir -p gpurun_out
if [ "$N" = 1 ]; then CMD=(python bench.py --gpus 1)
else CMD=(python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N); fi
timeout 600 "${CMD[@]}" --steps 20 --warmup 3 --scaling strong --no-cpu-baseline > gpurun_out/r2i_strong_${N}gpu.log 2> gpurun_out/r2i_strong_${N}gpu.err
timeout 600 "${CMD[@]}" --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_weak_${N}gpu.log 2> gpurun_out/r2i_weak_${N}gpu.err
echo done
