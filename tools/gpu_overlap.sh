#!/bin/bash
# after moving the step's Sinkhorn solve to a side stream: suite, smoke, full bench line
cd /root/repo; mkdir -p gpurun_out
T="timeout -k 10"
$T 600 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/ov_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ov_pytest.log
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ov_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/ov_smoke.log
$T 900 python bench.py > gpurun_out/ov_bench.json 2> gpurun_out/ov_bench.err; echo "rc=$?" >> gpurun_out/ov_bench.err
tail -3 gpurun_out/ov_pytest.log; tail -2 gpurun_out/ov_smoke.log; tail -1 gpurun_out/ov_bench.err; python tools/print_bench.py gpurun_out/ov_bench.json
