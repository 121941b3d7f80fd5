#!/bin/bash
# 2-GPU pass after the CTA-pair kernels: sharded-path parity, then the driver's own launch line for bench.py at N = 2
cd /root/repo; mkdir -p gpurun_out
T="timeout -k 10"
$T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > gpurun_out/n2f_mgpu_check.log 2>&1; echo "rc=$?" >> gpurun_out/n2f_mgpu_check.log
$T 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2f_bench.json 2> gpurun_out/n2f_bench.err; echo "rc=$?" >> gpurun_out/n2f_bench.err
tail -4 gpurun_out/n2f_mgpu_check.log; tail -2 gpurun_out/n2f_bench.err; head -c 400 gpurun_out/n2f_bench.json
