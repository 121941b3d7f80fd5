#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
T="timeout -k 10"
$T 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "gemm or hgcn or layer or stack" > gpurun_out/el2_pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/el2_pytest_gemm.log
$T 400 python bench.py --steps 20 --warmup 3 > gpurun_out/el2_bench.json 2> gpurun_out/el2_bench.err; echo "rc=$?" >> gpurun_out/el2_bench.err
tail -4 gpurun_out/el2_pytest_gemm.log; tail -1 gpurun_out/el2_bench.err; cut -c1-300 gpurun_out/el2_bench.json
