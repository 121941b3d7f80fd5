#!/bin/bash
# fifth GPU pass: compile-time cluster dims (ncu), constant reduction strides, fp64 fold accumulators in the row-block kernel
mkdir -p gpurun_out
T="timeout -k 10"
EG_PERSIST_TIMING=1 $T 90 python tools/t2_debug.py > gpurun_out/r5_t2_debug.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/r5_t2_debug.log
if [ $rc -ne 0 ]; then export EG_TILE2D=0; echo "tile2d disabled for the rest of this pass" >> gpurun_out/r5_t2_debug.log; fi
$T 400 python -m pytest tests/test_gpu_variants.py tests/test_gpu_sinkhorn_benched.py -m gpu -q --tb=short -s > gpurun_out/r5_pytest_sk.log 2>&1; echo "rc=$?" >> gpurun_out/r5_pytest_sk.log
EG_PERSIST_TIMING=1 $T 300 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_tile2d_kernel -c 1 -o gpurun_out/r02_tile2d python tools/sk_time.py > gpurun_out/r5_ncu_tile2d.log 2>&1; echo "rc=$?" >> gpurun_out/r5_ncu_tile2d.log
$T 420 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err; echo "rc=$?" >> gpurun_out/r5_bench.err
tail -4 gpurun_out/r5_t2_debug.log; tail -3 gpurun_out/r5_pytest_sk.log; tail -5 gpurun_out/r5_ncu_tile2d.log
