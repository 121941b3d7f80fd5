#!/bin/bash
# tenth GPU pass: launch list of the bench step, full captures of the tile Sinkhorn kernel and the short-chain GEMM, issue peaks
mkdir -p gpurun_out
T="timeout -k 10"
$T 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r10_ncu_launches.log 2>&1; echo "rc=$?" >> gpurun_out/r10_ncu_launches.log
$T 300 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_tile2d_kernel -c 1 -o gpurun_out/r02_tile2d_tmem python tools/sk_time.py > gpurun_out/r10_ncu_tile2d.log 2>&1; echo "rc=$?" >> gpurun_out/r10_ncu_tile2d.log
$T 300 ncu --set full --clock-control none --import-source on -k regex:lse_tc_kernel -s 8 -c 4 -o gpurun_out/r02_gemm_nt_chained python tools/gemm_nt_time.py > gpurun_out/r10_ncu_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/r10_ncu_gemm.log
$T 200 python tools/measure_peaks.py > gpurun_out/r10_peaks.log 2>&1; echo "rc=$?" >> gpurun_out/r10_peaks.log
tail -2 gpurun_out/r10_ncu_launches.log; tail -2 gpurun_out/r10_ncu_tile2d.log; tail -2 gpurun_out/r10_ncu_gemm.log; tail -2 gpurun_out/r10_peaks.log
