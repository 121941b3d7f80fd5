#!/bin/bash
# ninth GPU pass: short-chain NT GEMM (ReLU-feeding products off cuBLAS), full suite, bench
mkdir -p gpurun_out
T="timeout -k 10"
$T 120 python tools/gemm_nt_time.py > gpurun_out/r9_gemm_nt_time.log 2>&1; echo "rc=$?" >> gpurun_out/r9_gemm_nt_time.log
$T 500 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r9_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/r9_pytest_all.log
$T 420 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err; echo "rc=$?" >> gpurun_out/r9_bench.err
cat gpurun_out/r9_gemm_nt_time.log; grep "chained\|pinned\|passed\|failed\|FAILED" gpurun_out/r9_pytest_all.log | tail -20
