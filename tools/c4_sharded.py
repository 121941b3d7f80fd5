"""BASELINE.json config 4 at N GPUs: row-partitioned SpMM on a power-law graph, feature rows all-gathered
over NCCL before each aggregation (forward and transposed backward).  Run under torchrun."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
from gnn_mtl_b200 import ops, parallel as par
from gnn_mtl_b200.adjacency import DeviceAdjacency
from gnn_mtl_b200.synth import make_powerlaw_graph
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 20
h, t = make_powerlaw_graph(n, deg, seed=1)
full = DeviceAdjacency.from_heads_tails(n, torch.from_numpy(h).to(dev), torch.from_numpy(t).to(dev))
nnz = full.nnz
sh = par.ShardedAdjacency(full) if world > 1 else None
shh = par.ShardedAdjacency(full, halo=True) if world > 1 else None
if shh and rank == 0:
    print("halo: rank 0 fetches %d of %d remote rows; max fraction over ranks %.3f" % (shh.plan.n_need, shh.plan.n_remote_total, shh.remote_fraction), flush=True)
del h, t
res = []
for d in (128, 300):
    H_local = torch.randn((sh.r1 - sh.r0) if sh else n, d, device=dev)
    def fwd():
        Hf = sh.gather(H_local) if sh else H_local
        return ops.spmm(sh.csr if sh else full.csr, Hf)[0]
    def bwd():
        Hf = sh.gather(H_local) if sh else H_local
        return ops.spmm(sh.csr_t if sh else full.csr_t, Hf)[0]
    def gather_only():
        return sh.gather(H_local) if sh else H_local
    def fwd_overlap():
        return sh.aggregate_overlapped(H_local, n_chunks=4) if sh else fwd()
    def fwd_halo():
        return ops.spmm(shh.csr, shh.gather(H_local))[0] if shh else fwd()
    def halo_exchange_only():
        return shh.gather(H_local) if shh else H_local
    if shh:
        assert float((fwd_halo() - fwd()).abs().max()) < 1e-4     # needed-rows exchange: same result to fp32 rounding
    if sh:
        assert torch.equal(fwd_overlap(), fwd())          # column chunks do not change a single bit
    out = {}
    for name, f in (("fwd", fwd), ("bwd", bwd), ("allgather", gather_only), ("fwd_overlap", fwd_overlap),
                    ("fwd_halo", fwd_halo), ("halo_exchange", halo_exchange_only)):
        for _ in range(2): f()
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): f()
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 5], device=dev, dtype=torch.float64)
        if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[name] = float(ms[0])
    byt = nnz * 8 + (n + 1) * 4 + nnz * d * 4 + n * d * 4
    res.append({"d": d, "fwd_ms": out["fwd"], "bwd_ms": out["bwd"], "allgather_ms": out["allgather"],
                "fwd_overlap_ms": out["fwd_overlap"], "fwd_halo_ms": out["fwd_halo"], "halo_exchange_ms": out["halo_exchange"], "fwd_overlap_gbs_aggregate": byt / out["fwd_overlap"] / 1e6,
                "fwd_gbs_aggregate": byt / out["fwd"] / 1e6, "bwd_gbs_aggregate": byt / out["bwd"] / 1e6,
                "spmm_only_gbs_aggregate": byt / max(out["fwd"] - out["allgather"], 1e-6) / 1e6})
if rank == 0:
    print(json.dumps({"config": "row-partitioned SpMM, power-law graph n=%d target degree %d, nnz=%d" % (n, deg, nnz),
                      "n_gpus": world, "results": res}))
if world > 1: dist.destroy_process_group()
