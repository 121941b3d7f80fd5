#!/bin/bash
# N-GPU pass: the driver's own launch line for bench.py, main legs only.  usage: gpu_nN_bench.sh N
N=${1:-2}
cd /root/repo; mkdir -p gpurun_out
timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/n${N}b_bench.json 2> gpurun_out/n${N}b_bench.err; echo "rc=$?" >> gpurun_out/n${N}b_bench.err
tail -1 gpurun_out/n${N}b_bench.err; python tools/print_bench.py gpurun_out/n${N}b_bench.json
