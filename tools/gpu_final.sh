#!/bin/bash
# final pass: full suite, smoke, full bench (all extras)
mkdir -p gpurun_out
T="timeout -k 10"
$T 500 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/final_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/final_pytest_all.log
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/final_smoke.log
$T 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "rc=$?" >> gpurun_out/final_bench.err
grep "passed\|failed\|FAILED" gpurun_out/final_pytest_all.log | tail -4; tail -2 gpurun_out/final_smoke.log; tail -1 gpurun_out/final_bench.err
