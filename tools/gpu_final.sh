#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
T="timeout -k 10"
$T 600 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/tp2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/tp2_pytest.log
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/tp2_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/tp2_smoke.log
$T 900 python bench.py > gpurun_out/tp2_bench.json 2> gpurun_out/tp2_bench.err; echo "rc=$?" >> gpurun_out/tp2_bench.err
tail -3 gpurun_out/tp2_pytest.log; tail -2 gpurun_out/tp2_smoke.log; tail -1 gpurun_out/tp2_bench.err; cut -c1-260 gpurun_out/tp2_bench.json
