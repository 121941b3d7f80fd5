#!/bin/bash
# quick pass: tile kernel timing (shipping instantiation, then the per-phase timers) + Sinkhorn tests + NT GEMM timing
mkdir -p gpurun_out
T="timeout -k 10"
TAG=${1:-q}
$T 90 python tools/t2_debug.py > gpurun_out/${TAG}_t2_plain.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_t2_plain.log
EG_PERSIST_TIMING=1 $T 90 python tools/t2_debug.py > gpurun_out/${TAG}_t2_debug.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_t2_debug.log
$T 300 python -m pytest tests/test_gpu_variants.py tests/test_gpu_sinkhorn_benched.py -m gpu -q --tb=short -x > gpurun_out/${TAG}_pytest_sk.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_sk.log
$T 120 python tools/gemm_nt_time.py > gpurun_out/${TAG}_gemm_nt_time.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_gemm_nt_time.log
$T 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "gemm or hgcn or layers" > gpurun_out/${TAG}_pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_gemm.log
grep "iters 1000" gpurun_out/${TAG}_t2_plain.log; grep "iters 1000" -B1 gpurun_out/${TAG}_t2_debug.log; tail -3 gpurun_out/${TAG}_pytest_sk.log; cat gpurun_out/${TAG}_gemm_nt_time.log; tail -3 gpurun_out/${TAG}_pytest_gemm.log
