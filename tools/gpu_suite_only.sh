#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
T="timeout -k 10"
$T 400 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/fin_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/fin_pytest.log
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/fin_smoke.log
tail -3 gpurun_out/fin_pytest.log; tail -2 gpurun_out/fin_smoke.log
