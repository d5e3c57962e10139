#!/bin/bash
# round-2 fifth GPU pass: whole GPU suite + the default bench line (with e2e, broadcast e2e and cpu_baseline)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
timeout 600 python bench.py > gpurun_out/r2e_bench.log 2> gpurun_out/r2e_bench.err
echo done
