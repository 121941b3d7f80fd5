#!/bin/bash
cd /root/repo
timeout 240 python tools/gemm_pair_check.py > gpurun_out/pair_check.log 2>&1; echo "rc=$?" >> gpurun_out/pair_check.log
tail -14 gpurun_out/pair_check.log
EG_GEMM_RAW_DEBUG=1 timeout 240 python tools/gemm_pair_check.py 2>&1 | grep eagraft | tail -6
