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
Xl = X[r0:r1].clone()
do_eval = len(sys.argv) > 3 and sys.argv[3] == "eval"
if do_eval:
    vec = torch.cat([X, Y])                                 # get_hits takes the stacked embedding table
    pairs = torch.stack([perm, n + torch.arange(n, device=dev)], 1).cpu().numpy()   # Y[j] aligns with X[perm[j]]
del X
par.sinkhorn_fused_sharded(Xl, Y, a, b, 0.05, n, numItermax=2)      # warm-up incl. the marginal-error test of sweep 1
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
if do_eval:
    # config 5, second half: L1 ranks (Hits@k both directions) and per-row top-10 over the same n x n pair, rows sharded
    from gnn_mtl_b200 import ops
    ne = int(sys.argv[4]) if len(sys.argv) > 4 else n
    pairs = pairs[:ne]
    def timed(fn):
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); out = fn(); t1.record(); torch.cuda.synchronize()
        t = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t[0])
    par.get_hits_sharded(vec, pairs[:4096], top_k=(1, 10))          # warm-up: kernel module load, allocator
    ops.l1_topk(vec[:256], vec[n:], 0, 10)
    hits, ms_hits = timed(lambda: par.get_hits_sharded(vec, pairs, top_k=(1, 10)))
    q0, q1 = par.shard_range(ne, rank, world)
    Lq = vec[torch.as_tensor(pairs[q0:q1, 0], device=dev)]
    top, ms_top = timed(lambda: ops.l1_topk(Lq, vec[n:], 0, 10))
    want = torch.as_tensor(pairs[q0:q1, 1] - n, device=dev)
    top1_ok = float((top[:, 0] == want).float().mean()) if q1 > q0 else 1.0
    if rank == 0:
        print(json.dumps({"config": "L1 eval %d x %d d=300 (rows sharded)" % (ne, n), "n_gpus": world,
                          "get_hits_ms": ms_hits, "hits": hits, "top10_ms": ms_top, "top1_matches_link_frac_rank0": top1_ok,
                          "dadd_per_s_aggregate_T": ne * float(n) * 300 / ms_hits / 1e9}))
if world > 1: dist.destroy_process_group()
