"""Multi-GPU parity check (run under torchrun on N real GPUs; not collected by pytest):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Each sharded path must reproduce the single-GPU result computed on the same rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from gnn_mtl_b200 import parallel as par
    from gnn_mtl_b200.adjacency import DeviceAdjacency
    from gnn_mtl_b200.layers.layers import HighWayGraphConvolution
    from gnn_mtl_b200.synth import make_kg_pair
    from gnn_mtl_b200.utils.eval_utils import get_hits
    from gnn_mtl_b200.utils.ot_loss import sinkhorn_fused

    torch.manual_seed(0)                       # same data on every rank
    # ---- (b) row-sharded fused Sinkhorn vs the single-GPU solve -------------------------------
    I, J, d = 5000, 4100, 300
    X = (torch.randn(I, d) * 0.06).to(dev); Y = (torch.randn(J, d) * 0.06).to(dev)
    a = torch.full((I,), 1.0, device=dev); b = torch.full((J,), I / J, device=dev)
    info = {}
    _, loss_1 = sinkhorn_fused(X, Y, a, b, 0.05, numItermax=30, stopThr=0.0, info=info)
    r0, r1 = par.shard_range(I, rank, world)
    lu, lv, loss_n, inf2 = par.sinkhorn_fused_sharded(X[r0:r1], Y, a[r0:r1], b, 0.05, I, numItermax=30)
    e_u = float((lu - info["log_u"][r0:r1]).abs().max()); e_v = float((lv - info["log_v"]).abs().max())
    e_l = abs(float(loss_n) - float(loss_1)) / abs(float(loss_1))
    assert e_u < 5e-5 and e_v < 5e-5 and e_l < 1e-5, (e_u, e_v, e_l)
    # ---- (c) row-sharded Hits@k == single-GPU dict ----------------------------------------------
    kg = make_kg_pair("dbp15k")
    vec = torch.from_numpy(kg["x"]).to(dev)
    pairs = kg["test"][:4001]
    assert par.get_hits_sharded(vec, pairs, top_k=(1, 10)) == get_hits(vec, pairs, top_k=(1, 10))
    # ---- (a) row-partitioned highway layer, forward + backward -------------------------------------
    kg = make_kg_pair("tiny", dim=300)
    full = DeviceAdjacency.from_triples(kg["n"], kg["triples"], device=dev)
    sh = par.ShardedAdjacency(full)
    layer = HighWayGraphConvolution(300, 300, 0.0, F.relu, True, local, dev).to(dev)
    x = torch.from_numpy(kg["x"]).to(dev)
    xf = x.clone().requires_grad_(True)
    y_full, _ = layer((xf, full))
    seed = torch.randn_like(y_full)
    (y_full * seed).sum().backward()
    gW = layer.linear.weight.grad.clone(); layer.zero_grad()
    xl = x[sh.r0:sh.r1].clone().requires_grad_(True)
    y_loc, _ = layer((xl, sh))
    (y_loc * seed[sh.r0:sh.r1]).sum().backward()
    gWl = layer.linear.weight.grad.clone()
    dist.all_reduce(gWl)
    assert float((y_loc - y_full[sh.r0:sh.r1]).abs().max()) < 1e-5
    assert float((xl.grad - xf.grad[sh.r0:sh.r1]).abs().max()) < 1e-5
    assert float((gWl - gW).abs().max() / gW.abs().max()) < 1e-4
    # ---- (a') the same layer over the needed-rows exchange (HaloPlan): the all-gather route's result to fp32 rounding ----
    shh = par.ShardedAdjacency(full, halo=True)
    layer.zero_grad()
    xh = x[shh.r0:shh.r1].clone().requires_grad_(True)
    y_h, _ = layer((xh, shh))
    (y_h * seed[shh.r0:shh.r1]).sum().backward()
    assert float((y_h - y_loc).abs().max()) < 1e-5 and float((xh.grad - xl.grad).abs().max()) < 1e-5
    assert shh.plan.n_need <= full.n - (shh.r1 - shh.r0)
    # ---- (d) all-gather pipelined against the SpMM over column chunks: bit-identical to gather-then-SpMM ----
    from gnn_mtl_b200 import ops
    Hl = x[sh.r0:sh.r1].contiguous()
    plain = ops.spmm(sh.csr, sh.gather(Hl))[0]
    for chunks in (2, 4, 5):
        assert torch.equal(sh.aggregate_overlapped(Hl, n_chunks=chunks), plain)
    assert torch.equal(sh.aggregate_overlapped(Hl, transposed=True), ops.spmm(sh.csr_t, sh.gather(Hl))[0])
    dist.barrier()
    if rank == 0:
        print("mgpu_check ok: world %d | sinkhorn du %.1e dv %.1e dloss %.1e" % (world, e_u, e_v, e_l))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
