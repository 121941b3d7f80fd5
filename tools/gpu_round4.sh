#!/bin/bash
# fourth GPU pass: sign-tagged exchange in the tile kernel; ncu diagnostics for the cluster + cooperative launch
mkdir -p gpurun_out
T="timeout -k 10"
EG_PERSIST_TIMING=1 $T 90 python tools/t2_debug.py > gpurun_out/r4_t2_debug.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/r4_t2_debug.log
if [ $rc -ne 0 ]; then export EG_TILE2D=0; echo "tile2d disabled for the rest of this pass" >> gpurun_out/r4_t2_debug.log; fi
$T 400 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r4_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/r4_pytest_all.log
EG_PERSIST_TIMING=1 $T 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sinkhorn_tile2d -c 2 python tools/sk_time.py > gpurun_out/r4_ncu_a.log 2>&1; echo "rc=$?" >> gpurun_out/r4_ncu_a.log
EG_PERSIST_TIMING=1 $T 120 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:sinkhorn_tile2d -c 2 python tools/sk_time.py > gpurun_out/r4_ncu_b.log 2>&1; echo "rc=$?" >> gpurun_out/r4_ncu_b.log
EG_PERSIST_TIMING=1 $T 300 ncu --set full --clock-control none --replay-mode application --import-source on -k regex:sinkhorn_tile2d_kernel -c 1 -o gpurun_out/r02_tile2d_app python tools/sk_time.py > gpurun_out/r4_ncu_c.log 2>&1; echo "rc=$?" >> gpurun_out/r4_ncu_c.log
$T 420 python bench.py --steps 10 --warmup 3 > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err; echo "rc=$?" >> gpurun_out/r4_bench.err
tail -4 gpurun_out/r4_t2_debug.log; tail -3 gpurun_out/r4_pytest_all.log; tail -3 gpurun_out/r4_ncu_a.log; tail -3 gpurun_out/r4_ncu_b.log; tail -3 gpurun_out/r4_ncu_c.log
