#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 200 python tools/spmm_unroll_check.py > gpurun_out/spmm_unroll_check.log 2>&1; echo "rc=$?" >> gpurun_out/spmm_unroll_check.log
tail -14 gpurun_out/spmm_unroll_check.log
