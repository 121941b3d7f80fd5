#!/bin/bash
# second GPU pass: tile kernel debug first (short), then new kernels, the full suite, bench
mkdir -p gpurun_out
T="timeout -k 10"
EG_PERSIST_TIMING=1 $T 90 python tools/t2_debug.py > gpurun_out/r2_t2_debug.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/r2_t2_debug.log
if [ $rc -ne 0 ]; then export EG_TILE2D=0; echo "tile2d disabled for the rest of this pass" >> gpurun_out/r2_t2_debug.log; fi
EG_PERSIST_TIMING=1 $T 150 python tools/t2_check.py > gpurun_out/r2_t2_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t2_check.log
$T 200 python tools/eval_filter_check.py > gpurun_out/r2_eval_filter.log 2>&1; echo "rc=$?" >> gpurun_out/r2_eval_filter.log
$T 120 python tools/spmm_slab_check.py > gpurun_out/r2_spmm_slab.log 2>&1; echo "rc=$?" >> gpurun_out/r2_spmm_slab.log
$T 800 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r2_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_all.log
$T 420 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "rc=$?" >> gpurun_out/r2_bench.err
tail -4 gpurun_out/r2_t2_debug.log; tail -3 gpurun_out/r2_pytest_all.log
