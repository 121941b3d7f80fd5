#!/bin/bash
# quick pass: tile kernel timing + Sinkhorn tests
mkdir -p gpurun_out
T="timeout -k 10"
TAG=${1:-q}
EG_PERSIST_TIMING=1 $T 90 python tools/t2_debug.py > gpurun_out/${TAG}_t2_debug.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_t2_debug.log
$T 300 python -m pytest tests/test_gpu_variants.py tests/test_gpu_sinkhorn_benched.py -m gpu -q --tb=short -x > gpurun_out/${TAG}_pytest_sk.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_sk.log
grep "iters 1000" -B1 gpurun_out/${TAG}_t2_debug.log; tail -3 gpurun_out/${TAG}_pytest_sk.log
