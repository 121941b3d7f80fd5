#!/bin/bash
cd /root/repo
timeout 240 python tools/gemm_pair_check.py > gpurun_out/pair_check.log 2>&1; echo "rc=$?" >> gpurun_out/pair_check.log
tail -5 gpurun_out/pair_check.log
EG_GEMM_RAW_DEBUG=1 timeout 240 python tools/gemm_pair_one.py 2>&1 | grep eagraft | tail -2
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_raw2_kernel -s 2 -c 1 -o gpurun_out/r02_gemm_nt_raw2 python tools/gemm_pair_one.py > gpurun_out/ncu_raw2.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_raw2.log
tail -2 gpurun_out/ncu_raw2.log
