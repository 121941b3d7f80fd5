"""BASELINE.json config 5: streamed fused Sinkhorn on a synthetic I x J pair (300-d), rows of X sharded over ranks,
cost never materialised; then sharded Hits@k on a sub-sample.  Run under torchrun; prints one JSON line on rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
from gnn_mtl_b200 import parallel as par
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
g = torch.Generator(device=dev); g.manual_seed(0)            # same data on every rank
X = torch.randn(n, 300, device=dev, generator=g) / 300 ** 0.5
perm = torch.randperm(n, device=dev, generator=g)
Y = X[perm] + 0.1 * torch.randn(n, 300, device=dev, generator=g) / 300 ** 0.5
r0, r1 = par.shard_range(n, rank, world)
a = torch.full((r1 - r0,), 1.0 / n, device=dev); b = torch.full((n,), 1.0 / n, device=dev)
Xl = X[r0:r1].clone(); del X
par.sinkhorn_fused_sharded(Xl, Y, a, b, 0.05, n, numItermax=1)      # warm-up (split, norms, first launches)
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
lu, lv, loss, info = par.sinkhorn_fused_sharded(Xl, Y, a, b, 0.05, n, numItermax=sweeps)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    per = float(ms[0]) / (sweeps + 0.5)      # sweeps full sweeps + the final plan pass (one more half-sweep of tiles)
    print(json.dumps({"config": "fused Sinkhorn %dx%d d=300 reg=0.05" % (n, n), "n_gpus": world, "sweeps": sweeps,
                      "ms_total": float(ms[0]), "ms_per_sweep": per, "sweeps_per_s": 1e3 / per,
                      "tf32_mma_tflops_aggregate": 2 * 3 * 2.0 * n * n * 300 / per / 1e9, "loss": float(loss)}))
if world > 1: dist.destroy_process_group()
