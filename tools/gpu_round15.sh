#!/bin/bash
# fifteenth GPU pass: full suite + bench after the raw-operand dx GEMM and the gate / chained split
mkdir -p gpurun_out
T="timeout -k 10"
$T 500 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r15_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/r15_pytest_all.log
$T 420 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r15_bench.json 2> gpurun_out/r15_bench.err; echo "rc=$?" >> gpurun_out/r15_bench.err
grep "passed\|failed\|FAILED" gpurun_out/r15_pytest_all.log | tail -5; tail -2 gpurun_out/r15_bench.err
