#!/bin/bash
# round-2 re-entry pass: GPU suite, default bench line (with e2e + cpu baseline), C3' (pbm covariates) bench line + launch list
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
timeout 600 python bench.py > gpurun_out/r2j_bench_default.log 2> gpurun_out/r2j_bench_default.err
timeout 300 python bench.py --sites 1000000 --n-mc 32 --scenario covarK2 --nan-frac 0.1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2j_bench_c3p.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2j_launches_c3p.csv python bench.py --sites 1000000 --n-mc 32 --scenario covarK2 --nan-frac 0.1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2j_ncu_c3p.log 2>&1
echo done
