#!/bin/bash
# full GPU suite + smoke + default bench (both arms); $1 = tag
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x ) > gpurun_out/$1_pytest.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/$1_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/$1_smoke.log
timeout 600 python bench.py > gpurun_out/$1_bench_default.log 2> gpurun_out/$1_bench_default.err
echo done
