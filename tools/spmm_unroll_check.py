"""SpMM on the benchmark graph: neighbour rows in flight per warp (unroll) x warps per CTA of the shipping kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import _lib, ops
from gnn_mtl_b200.adjacency import DeviceAdjacency
from gnn_mtl_b200.synth import make_kg_pair
dev = torch.device("cuda:0")
dbg = _lib.lib.eg_debug_set
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def bench(f, n=10):
    for _ in range(2): f()
    ts = []
    for _ in range(n):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
kg = make_kg_pair("dbp100k", features=False)
adj = DeviceAdjacency.from_triples(kg["n"], kg["triples"], device=dev)
c = adj.csr
for d in (300, 128):
    H = torch.randn(kg["n"], d, device=dev); g = torch.randn_like(H); xr = torch.randn_like(H)
    for unroll, warps in ((2, 4), (4, 4), (8, 4), (1, 8), (2, 8), (4, 8)):
        dbg(0, unroll); dbg(1, warps)
        t1 = bench(lambda: ops.spmm(c, H)); t2 = bench(lambda: ops.spmm(c, H, _lib.ACT_RELU, g, xr, True))
        print("d=%d unroll %d warps %d: plain %.3f ms fused+save %.3f ms" % (d, unroll, warps, t1, t2), flush=True)
    dbg(0, 2); dbg(1, 4)
