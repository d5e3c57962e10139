#!/bin/bash
# final tcgen05 applicator (issuer warp): ncu --set full of both kernels in the C3' workload, launch lists, small-batch latencies
mkdir -p gpurun_out
bash profiles/scripts/r2l.sh k_mlp3_tc_bwd r2p_tc_bwd_v4 3
bash profiles/scripts/r2l.sh k_mlp3_tc_fwd r2p_tc_fwd_v4 3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2p_launches_c3p.csv python bench.py --sites 1000000 --n-mc 32 --scenario covarK2 --nan-frac 0.1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2p_ncu_c3p.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2p_launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2p_ncu_default.log 2>&1
timeout 300 python bench.py --sites 1000000 --n-mc 32 --scenario covarK2 --nan-frac 0.1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2p_bench_c3p.log 2>&1
timeout 400 python profiles/bench_small_batch.py > gpurun_out/r2p_small.log 2>&1
echo done
