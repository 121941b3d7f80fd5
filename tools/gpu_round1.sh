#!/bin/bash
# first GPU pass of the round: tile kernel sanity + timing, the GPU suite, a bench line
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r1_smi.txt 2>&1
EG_PERSIST_TIMING=1 timeout 300 python tools/t2_check.py > gpurun_out/r1_t2_check.log 2>&1; echo "t2_check rc=$?" >> gpurun_out/r1_t2_check.log
timeout 1500 python -m pytest tests -m gpu -x -q -s -k "sinkhorn" > gpurun_out/r1_pytest_sinkhorn.log 2>&1; echo "rc=$?" >> gpurun_out/r1_pytest_sinkhorn.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r1_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/r1_pytest_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "rc=$?" >> gpurun_out/r1_bench.err
tail -5 gpurun_out/r1_t2_check.log; tail -3 gpurun_out/r1_pytest_sinkhorn.log; tail -3 gpurun_out/r1_pytest_all.log
