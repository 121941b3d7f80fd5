#!/bin/bash
# 8-GPU pass: headline step with the overlapped and with the flat gradient all-reduce
mkdir -p gpurun_out
for flat in 0 1; do
  EG_BENCH_FLAT_ALLREDUCE=$flat timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$flat bench.py --gpus 8 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/n8_sync_flat$flat.json 2> gpurun_out/n8_sync_flat$flat.err; echo "rc=$?" >> gpurun_out/n8_sync_flat$flat.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/n8_sync_flat$flat.json') if l.startswith('{')][-1])
print('flat=$flat', d['value'], d['ms_per_step'])
PY
done
