#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_tile2d_kernel -c 1 -o gpurun_out/r02_tile2d_final python tools/sk_time.py > gpurun_out/ncu_t2_final.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_t2_final.log
tail -3 gpurun_out/ncu_t2_final.log
