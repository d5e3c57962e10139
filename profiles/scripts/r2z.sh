#!/bin/bash
# launch list (ncu, durations only) of one config-5-shaped evaluation + the blocked-schedule tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_fullsize.py -q -m gpu -x -k "C5 or blocked" > gpurun_out/r2z_pytest.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3400 --csv --log-file gpurun_out/r2z_launches_c5small.csv \
  python bench.py --sites 100000 --n-mc 16 --scenario covarK2 --hidden 512,512,512,512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2z_ncu.log 2>&1
echo done
