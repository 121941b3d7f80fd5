#!/bin/bash
# one ncu --set full capture of the fused Sinkhorn half-sweep kernel (CTA-pair instantiation), 30000 x 30000 x 300
cd /root/repo; mkdir -p gpurun_out
timeout -k 10 120 python tools/tc_one.py 30000 > gpurun_out/ncu_lse_plain.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_lse_plain.log
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:lse_tc_kernel -s 1 -c 1 -o gpurun_out/r02_lse_tc_pair python tools/tc_one.py 30000 > gpurun_out/ncu_lse.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_lse.log
tail -2 gpurun_out/ncu_lse_plain.log; tail -3 gpurun_out/ncu_lse.log
