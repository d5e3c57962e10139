#!/bin/bash
# final verification: GPU suite, smoke, default bench line (with e2e and cpu baseline), reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_smoke.log
timeout 600 python bench.py > gpurun_out/r2s_bench_default.log 2> gpurun_out/r2s_bench_default.err
timeout 600 python bench.py --impl reference > gpurun_out/r2s_bench_reference.log 2> gpurun_out/r2s_bench_reference.err
echo done
