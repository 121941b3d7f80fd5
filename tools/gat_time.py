"""Time the GAT aggregation kernels on the dbp100k-shaped adjacency against the reference's formulation
(edge gather + exp + two torch.spmm + divide, layers/att_layers.py:38-59) run by stock PyTorch on the same GPU."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from gnn_mtl_b200 import ops, synth
from gnn_mtl_b200.adjacency import DeviceAdjacency

dev = torch.device("cuda:0")
kg = synth.make_kg_pair("dbp100k", seed=0)
tri = torch.as_tensor(kg["triples"])
n = int(kg["n"])
A = DeviceAdjacency.from_heads_tails(n, tri[:, 0].to(dev), tri[:, 2].to(dev))
nnz = A.csr.nnz
crow = A.csr.rowptr.long()
rows = torch.repeat_interleave(torch.arange(n, device=dev), crow[1:] - crow[:-1])
cols = A.csr.col.long()
edge = torch.stack([rows, cols])
A.csr_t  # build once


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for d in (75, 300):
    h = torch.randn(n, d, device=dev, requires_grad=True)
    a = (torch.randn(1, 2 * d, device=dev) / d ** 0.5).requires_grad_(True)
    seed = torch.randn(n, d, device=dev)

    def ours(bwd):
        s1, s2 = h @ a[0, :d], h @ a[0, d:]
        y = ops.gat_aggregate(h, s1, s2, A, 0.2)
        if bwd:
            h.grad = a.grad = None
            (y * seed).sum().backward()
        return y

    def stock(bwd):
        edge_h = torch.cat((h[edge[0]], h[edge[1]]), dim=1).t()
        e = torch.exp(-F.leaky_relu(a.mm(edge_h).squeeze(), 0.2))
        # torch.sparse.mm: same forward as the reference's torch.spmm, but a SPARSE gradient for the edge values
        # (torch.spmm's backward materialises a dense n x n gradient: 149 GiB at n = 200k, i.e. the reference's
        # own formulation cannot train at this size)
        rowsum = torch.sparse.mm(torch.sparse_coo_tensor(edge, e, (n, n)), torch.ones(n, 1, device=dev))
        y = torch.sparse.mm(torch.sparse_coo_tensor(edge, e, (n, n)), h).div(rowsum)
        if bwd:
            h.grad = a.grad = None
            (y * seed).sum().backward()
        return y

    with torch.no_grad():
        err = float((ours(False) - stock(False)).abs().max())
    t_of, t_ob = timed(lambda: ours(False)), timed(lambda: ours(True))
    print("d=%d ours fwd %.3f ms fwd+bwd %.3f ms" % (d, t_of, t_ob), flush=True)
    t_sf = timed(lambda: stock(False))
    try:
        t_sb = timed(lambda: stock(True), reps=5)
    except Exception as exc:   # noqa: BLE001
        print("stock backward failed:", type(exc).__name__, flush=True)
        t_sb = float("nan")
    gbytes = (nnz * (4 + 4 * d) + n * d * 4) / 1e9
    print("d=%d n=%d nnz=%d  ours fwd %.3f ms (%.0f GB/s alg)  fwd+bwd %.3f ms | stock fwd %.3f ms  fwd+bwd %.3f ms | "
          "speedup fwd %.1fx  fwd+bwd %.1fx  maxabs diff %.2e"
          % (d, n, nnz, t_of, gbytes / t_of * 1e3, t_ob, t_sf, t_sb, t_sf / t_of, t_sb / t_ob, err), flush=True)
