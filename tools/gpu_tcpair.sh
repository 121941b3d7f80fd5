#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout -k 10 300 python tools/tc_pair_check.py > gpurun_out/tcpair_check.log 2>&1; echo "rc=$?" >> gpurun_out/tcpair_check.log
grep "lse 30000\|gemm_nt n=\|rc=" gpurun_out/tcpair_check.log; grep -c "identical True" gpurun_out/tcpair_check.log; grep -c "identical False" gpurun_out/tcpair_check.log
EG_TC_PAIR=1 timeout -k 10 600 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/tcpair_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/tcpair_pytest.log
tail -4 gpurun_out/tcpair_pytest.log
