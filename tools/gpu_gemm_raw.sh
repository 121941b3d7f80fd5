#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 100 python tools/gemm_raw_dbg.py > gpurun_out/gemm_raw_dbg.log 2>&1; echo "rc=$?" >> gpurun_out/gemm_raw_dbg.log
tail -8 gpurun_out/gemm_raw_dbg.log
