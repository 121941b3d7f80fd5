#!/bin/bash
# 2-GPU pass: the driver's own launch line for bench.py at N = 2 (all extras on)
mkdir -p gpurun_out
timeout -k 10 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "rc=$?" >> gpurun_out/n2_bench.err
tail -3 gpurun_out/n2_bench.err; head -c 600 gpurun_out/n2_bench.json
