#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout -k 10 900 python bench.py > gpurun_out/bo_bench.json 2> gpurun_out/bo_bench.err; echo "rc=$?" >> gpurun_out/bo_bench.err
tail -1 gpurun_out/bo_bench.err; python tools/print_bench.py gpurun_out/bo_bench.json
