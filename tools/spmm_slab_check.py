"""SpMM on the benchmark graph: feature-slab width 0 (whole row) / 64 / 32 float4 per launch — time, gather-model GB/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import _lib, ops
from gnn_mtl_b200.adjacency import DeviceAdjacency
from gnn_mtl_b200.synth import make_kg_pair
dev = torch.device("cuda:0")
kg = make_kg_pair("dbp100k", features=False)
adj = DeviceAdjacency.from_triples(kg["n"], kg["triples"], device=dev)
c = adj.csr
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def bench(f, n=10, do_flush=True):
    for _ in range(2): f()
    ts = []
    for _ in range(n):
        if do_flush: flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
for d in (300, 128):
    H = torch.randn(kg["n"], d, device=dev); g = torch.randn_like(H); xr = torch.randn_like(H)
    byt = c.nnz * 8 + (c.n_rows + 1) * 4 + c.nnz * d * 4 + c.n_rows * d * 4
    ref = None
    for slab in (0, 64, 32):
        _lib.lib.eg_debug_set(14, slab)
        out = ops.spmm(c, H)[0]
        if ref is None: ref = out
        same = bool(torch.equal(out, ref))
        t1 = bench(lambda: ops.spmm(c, H)); t2 = bench(lambda: ops.spmm(c, H, _lib.ACT_RELU, g, xr, True))
        t3 = bench(lambda: ops.spmm(c, H), do_flush=False)
        print("d=%d slab %2d: plain %.3f ms %5.0f GB/s | fused+save %.3f ms %5.0f GB/s | plain warm-L2 %.3f ms | bit-identical %s" %
              (d, slab, t1, byt / t1 / 1e6, t2, (byt + 3 * c.n_rows * d * 4) / t2 / 1e6, t3, same), flush=True)
    _lib.lib.eg_debug_set(14, 0)
