#!/bin/bash
# closed-form gradient of the as-shipped Wasserstein objective: A/B of the step, then the suite
cd /root/repo; mkdir -p gpurun_out
T="timeout -k 10"
for cf in 0 1; do
  EG_COL0_GRAD=$cf $T 200 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/cf${cf}_bench.json 2> gpurun_out/cf${cf}_bench.err; echo "rc=$?" >> gpurun_out/cf${cf}_bench.err
  echo "closed form=$cf: $(python tools/print_bench.py gpurun_out/cf${cf}_bench.json 2>/dev/null | cut -c1-90)"; tail -1 gpurun_out/cf${cf}_bench.err
done
$T 300 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/cf_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/cf_pytest.log
tail -3 gpurun_out/cf_pytest.log
