#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 200 python tools/spmm_bulk_check.py > gpurun_out/spmm_bulk_check.log 2>&1; echo "rc=$?" >> gpurun_out/spmm_bulk_check.log
tail -12 gpurun_out/spmm_bulk_check.log
