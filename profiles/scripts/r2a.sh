#!/bin/bash
# round-2 first GPU pass: full GPU suite, default bench, small-batch latencies, ncu --set full of the shipped k_stream
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_bench.log
timeout 300 python profiles/bench_small_batch.py > gpurun_out/r2a_small.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_stream -c 1 -s 2 -f -o gpurun_out/r2a_k_stream python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2a_ncu.log 2>&1
echo done
