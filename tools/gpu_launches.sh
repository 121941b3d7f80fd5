#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
T="timeout -k 10"
$T 300 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ll_bench_plain.json 2> gpurun_out/ll_bench_plain.err; echo "rc=$?" >> gpurun_out/ll_bench_plain.err
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02b_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ll_ncu_launches.log 2>&1; echo "rc=$?" >> gpurun_out/ll_ncu_launches.log
tail -1 gpurun_out/ll_bench_plain.err; tail -2 gpurun_out/ll_ncu_launches.log | cut -c1-200; wc -l gpurun_out/r02b_launches_bench.csv
