"""SpMM tuning sweep on the benchmark graph: GB/s by (unroll, warps/CTA, L2 hints)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import _lib, ops
from gnn_mtl_b200.adjacency import DeviceAdjacency
from gnn_mtl_b200.synth import make_kg_pair
shape = sys.argv[1] if len(sys.argv) > 1 else "dbp100k"
dev = torch.device("cuda:0")
kg = make_kg_pair(shape, features=False)
adj = DeviceAdjacency.from_triples(kg["n"], kg["triples"], device=dev)
c = adj.csr
d = 300
H = torch.randn(kg["n"], d, device=dev); g = torch.randn_like(H); xr = torch.randn_like(H)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
byt = c.nnz * 8 + (c.n_rows + 1) * 4 + c.nnz * d * 4 + c.n_rows * d * 4
def bench(f, n=10):
    for _ in range(2): f()
    ts = []
    for _ in range(n):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
dbg = _lib.lib.eg_debug_set
print("n=%d nnz=%d alg bytes plain %.2f GB fused %.2f GB" % (c.n_rows, c.nnz, byt / 1e9, (byt + 3 * c.n_rows * d * 4) / 1e9))
for persist in (0, 2, 3, 4, 6):
  for hints in (0,):
    for unroll, warps in ((2, 4), (4, 4), (1, 8)) if persist else ((2, 4),):
        dbg(0, unroll); dbg(1, warps); dbg(2, hints); dbg(6, persist)
        t1 = bench(lambda: ops.spmm(c, H))
        t2 = bench(lambda: ops.spmm(c, H, _lib.ACT_RELU, g, xr, True))
        print("persist %d hints %d unroll %d warps %d | plain %.3f ms %5.0f GB/s | fused+save %.3f ms %5.0f GB/s" %
              (persist, hints, unroll, warps, t1, byt / t1 / 1e6, t2, (byt + 3 * c.n_rows * d * 4) / t2 / 1e6))
