#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 200 python tools/spmm_dyn_check.py > gpurun_out/spmm_dyn_check.log 2>&1; echo "rc=$?" >> gpurun_out/spmm_dyn_check.log
tail -30 gpurun_out/spmm_dyn_check.log
