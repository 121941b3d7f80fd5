#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_raw_kernel -s 2 -c 1 -o gpurun_out/r02_gemm_nt_raw python tools/gemm_raw_one.py > gpurun_out/ncu_raw.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_raw.log
tail -3 gpurun_out/ncu_raw.log
