#!/bin/bash
# 8-GPU pass: the driver's own launch line for bench.py at N = 8 (all extras on)
mkdir -p gpurun_out
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err; echo "rc=$?" >> gpurun_out/n8_bench.err
tail -3 gpurun_out/n8_bench.err; head -c 300 gpurun_out/n8_bench.json
