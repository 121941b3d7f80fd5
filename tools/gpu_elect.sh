#!/bin/bash
# after the warp-uniform MMA issuers: GEMM checks + timings, then the whole suite and the bench
cd /root/repo; mkdir -p gpurun_out
T="timeout -k 10"
$T 200 python tools/gemm_pair_check.py > gpurun_out/el_pair_check.log 2>&1; echo "rc=$?" >> gpurun_out/el_pair_check.log
$T 200 python tools/gemm_raw_check.py > gpurun_out/el_raw_check.log 2>&1; echo "rc=$?" >> gpurun_out/el_raw_check.log
$T 120 python tools/gemm_nt_time.py > gpurun_out/el_gemm_nt_time.log 2>&1; echo "rc=$?" >> gpurun_out/el_gemm_nt_time.log
$T 600 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/el_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/el_pytest_all.log
$T 400 python bench.py --steps 20 --warmup 3 > gpurun_out/el_bench.json 2> gpurun_out/el_bench.err; echo "rc=$?" >> gpurun_out/el_bench.err
tail -5 gpurun_out/el_pair_check.log; tail -5 gpurun_out/el_raw_check.log; cat gpurun_out/el_gemm_nt_time.log; tail -4 gpurun_out/el_pytest_all.log; tail -1 gpurun_out/el_bench.err; cut -c1-400 gpurun_out/el_bench.json
