#!/bin/bash
# twelfth GPU pass: full suite, smoke, full bench (all extras), captures: short-chain GEMM, config-4 SpMM DRAM bytes
mkdir -p gpurun_out
T="timeout -k 10"
$T 500 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r12_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/r12_pytest_all.log
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r12_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r12_smoke.log
$T 600 python bench.py > gpurun_out/r12_bench.json 2> gpurun_out/r12_bench.err; echo "rc=$?" >> gpurun_out/r12_bench.err
$T 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:lse_tc_kernelILi3 -s 4 -c 2 -o gpurun_out/r02_gemm_nt_chained python tools/gemm_nt_time.py > gpurun_out/r12_ncu_gemm3.log 2>&1; echo "rc=$?" >> gpurun_out/r12_ncu_gemm3.log
for cfg in "1000000 20" "10000000 5"; do
  set -- $cfg
  $T 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:spmm_vec --csv --log-file gpurun_out/r02_c4_spmm_dram_n$1.csv python tools/c4_spmm_one.py $1 $2 > gpurun_out/r12_c4_n$1.log 2>&1; echo "rc=$?" >> gpurun_out/r12_c4_n$1.log
done
tail -3 gpurun_out/r12_pytest_all.log; tail -2 gpurun_out/r12_smoke.log; tail -1 gpurun_out/r12_bench.err; tail -2 gpurun_out/r12_ncu_gemm3.log; tail -1 gpurun_out/r12_c4_n1000000.log gpurun_out/r12_c4_n10000000.log
