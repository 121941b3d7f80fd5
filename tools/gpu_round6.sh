#!/bin/bash
# sixth GPU pass: st.async + mbarrier row exchange, pipelined shared-memory rows; TMEM vs SMEM read bandwidth; issue peaks
mkdir -p gpurun_out
T="timeout -k 10"
$T 60 tools/tmem_bw > gpurun_out/r6_tmem_bw.log 2>&1; echo "rc=$?" >> gpurun_out/r6_tmem_bw.log
EG_PERSIST_TIMING=1 $T 90 python tools/t2_debug.py > gpurun_out/r6_t2_debug.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/r6_t2_debug.log
if [ $rc -ne 0 ]; then export EG_TILE2D=0; echo "tile2d disabled for the rest of this pass" >> gpurun_out/r6_t2_debug.log; fi
$T 400 python -m pytest tests/test_gpu_variants.py tests/test_gpu_sinkhorn_benched.py -m gpu -q --tb=short -s > gpurun_out/r6_pytest_sk.log 2>&1; echo "rc=$?" >> gpurun_out/r6_pytest_sk.log
$T 200 python tools/measure_peaks.py > gpurun_out/r6_peaks.log 2>&1; echo "rc=$?" >> gpurun_out/r6_peaks.log
$T 420 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err; echo "rc=$?" >> gpurun_out/r6_bench.err
cat gpurun_out/r6_tmem_bw.log; tail -4 gpurun_out/r6_t2_debug.log; tail -3 gpurun_out/r6_pytest_sk.log; tail -3 gpurun_out/r6_peaks.log
