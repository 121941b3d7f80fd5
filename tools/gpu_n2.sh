#!/bin/bash
# 2-GPU pass: sharded-path parity (incl. the needed-rows exchange) and the config-4 SpMM sweep at 2 ranks
mkdir -p gpurun_out
T="timeout -k 10"
$T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > gpurun_out/n2_mgpu_check.log 2>&1; echo "rc=$?" >> gpurun_out/n2_mgpu_check.log
$T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/c4_sharded.py 10000000 5 > gpurun_out/n2_c4.log 2>&1; echo "rc=$?" >> gpurun_out/n2_c4.log
tail -5 gpurun_out/n2_mgpu_check.log; tail -4 gpurun_out/n2_c4.log
