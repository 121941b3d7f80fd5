"""SpMM on the benchmark graph and on a power-law graph with hub rows: register-gather kernel (eg_debug_set(16, 0))
against the cp.async.bulk + shared-memory kernel (16, 1) — identical bits, time, gather-model GB/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import _lib, ops
from gnn_mtl_b200.adjacency import DeviceAdjacency
from gnn_mtl_b200.synth import make_kg_pair, make_powerlaw_graph
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def bench(f, n=10, do_flush=True):
    for _ in range(2): f()
    ts = []
    for _ in range(n):
        if do_flush: flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
def run(name, c, ct, n):
    for d in (300, 128, 52):
        H = torch.randn(n, d, device=dev); g = torch.randn_like(H); xr = torch.randn_like(H)
        byt = c.nnz * 8 + (c.n_rows + 1) * 4 + c.nnz * d * 4 + c.n_rows * d * 4
        res = {}
        for bulk in (0, 1):
            _lib.lib.eg_debug_set(16, bulk)
            o1 = ops.spmm(c, H)[0]; o2, a2 = ops.spmm(c, H, _lib.ACT_RELU, g, xr, True); o3 = ops.spmm(ct, H)[0]
            t1 = bench(lambda: ops.spmm(c, H)); t2 = bench(lambda: ops.spmm(c, H, _lib.ACT_RELU, g, xr, True))
            res[bulk] = (o1, o2, a2, o3, t1, t2)
        _lib.lib.eg_debug_set(16, 0)
        same = all(torch.equal(res[0][i], res[1][i]) for i in range(4))
        print("%s d=%d: gather plain %.3f ms %5.0f GB/s fused+save %.3f ms | bulk plain %.3f ms %5.0f GB/s fused+save %.3f ms | identical %s" %
              (name, d, res[0][4], byt / res[0][4] / 1e6, res[0][5], res[1][4], byt / res[1][4] / 1e6, res[1][5], same), flush=True)
        assert same
kg = make_kg_pair("dbp100k", features=False)
adj = DeviceAdjacency.from_triples(kg["n"], kg["triples"], device=dev)
run("dbp100k", adj.csr, adj.csr_t, kg["n"])
h, t = make_powerlaw_graph(1000000, 20, seed=1)
adj = DeviceAdjacency.from_heads_tails(1000000, torch.from_numpy(h).to(dev), torch.from_numpy(t).to(dev))
print("power-law: n_long %d n_seg %d" % (adj.csr.n_long, adj.csr.n_seg))
run("powerlaw1M", adj.csr, adj.csr_t, 1000000)
