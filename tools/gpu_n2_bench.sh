#!/bin/bash
# 2-GPU pass: the driver's own launch line for bench.py at N = 2, main legs only
cd /root/repo; mkdir -p gpurun_out
timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/n2b_bench.json 2> gpurun_out/n2b_bench.err; echo "rc=$?" >> gpurun_out/n2b_bench.err
tail -1 gpurun_out/n2b_bench.err; python tools/print_bench.py gpurun_out/n2b_bench.json
