#!/bin/bash
# the C-ABI collective on two GPUs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nccl_capi.py -q -m gpu -x > gpurun_out/r2r_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_nccl.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29534 tests/nccl_two_ranks.py > gpurun_out/r2r_two_ranks.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_two_ranks.log
echo done
