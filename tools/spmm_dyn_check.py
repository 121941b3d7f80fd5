"""SpMM on the benchmark graph: one-launch-per-row-group kernel (default) against the persistent kernel with static
round-robin and with an atomic work counter (eg_debug_set(6, n CTAs per SM), (17, 1)): identical bits, time."""
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
    ref = ops.spmm(c, H)[0]; ref2 = ops.spmm(c, H, _lib.ACT_RELU, g, xr, True)
    t1 = bench(lambda: ops.spmm(c, H)); t2 = bench(lambda: ops.spmm(c, H, _lib.ACT_RELU, g, xr, True))
    print("d=%d default: plain %.3f ms fused+save %.3f ms" % (d, t1, t2), flush=True)
    for persist in (2, 3, 4):
        for dyn in (0, 1):
            for unroll in (2, 4):
                dbg(6, persist); dbg(17, dyn); dbg(0, unroll)
                o = ops.spmm(c, H)[0]; o2 = ops.spmm(c, H, _lib.ACT_RELU, g, xr, True)
                same = torch.equal(o, ref) and torch.equal(o2[0], ref2[0]) and torch.equal(o2[1], ref2[1])
                t1 = bench(lambda: ops.spmm(c, H)); t2 = bench(lambda: ops.spmm(c, H, _lib.ACT_RELU, g, xr, True))
                print("d=%d persist %d CTAs/SM dynamic %d unroll %d: plain %.3f ms fused+save %.3f ms identical %s" %
                      (d, persist, dyn, unroll, t1, t2, same), flush=True)
    dbg(6, 0); dbg(17, 0); dbg(0, 2)
