#!/bin/bash
# third GPU pass: tile kernel after the load hoist, full suite, ncu captures (tile kernel, rank filter, launch list), bench
mkdir -p gpurun_out
T="timeout -k 10"
EG_PERSIST_TIMING=1 $T 90 python tools/t2_debug.py > gpurun_out/r3_t2_debug.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/r3_t2_debug.log
if [ $rc -ne 0 ]; then export EG_TILE2D=0; echo "tile2d disabled for the rest of this pass" >> gpurun_out/r3_t2_debug.log; fi
$T 400 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r3_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/r3_pytest_all.log
$T 400 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_tile2d_kernel -c 1 -o gpurun_out/r02_tile2d python tools/sk_time.py > gpurun_out/r3_ncu_tile2d.log 2>&1; echo "rc=$?" >> gpurun_out/r3_ncu_tile2d.log
$T 400 ncu --set full --clock-control none --import-source on -k regex:l1_rank_filter_kernel -s 1 -c 1 -o gpurun_out/r02_rank_filter python tools/rank_filter_one.py > gpurun_out/r3_ncu_filter.log 2>&1; echo "rc=$?" >> gpurun_out/r3_ncu_filter.log
$T 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r3_ncu_launches.log 2>&1; echo "rc=$?" >> gpurun_out/r3_ncu_launches.log
$T 420 python bench.py --steps 10 --warmup 3 > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; echo "rc=$?" >> gpurun_out/r3_bench.err
tail -4 gpurun_out/r3_t2_debug.log; tail -3 gpurun_out/r3_pytest_all.log; tail -2 gpurun_out/r3_ncu_tile2d.log; tail -2 gpurun_out/r3_ncu_filter.log; tail -2 gpurun_out/r3_ncu_launches.log
