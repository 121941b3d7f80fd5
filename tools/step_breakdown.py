"""Phase-level CUDA-event timing of the benchmark step (diagnostic, not the bench)."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from gnn_mtl_b200 import _lib, ops
from gnn_mtl_b200.adjacency import DeviceAdjacency
from gnn_mtl_b200.models.models_ea import UEAModel
from gnn_mtl_b200.synth import make_kg_pair
from gnn_mtl_b200.utils.ot_loss import sinkhorn
import bench

shape = sys.argv[1] if len(sys.argv) > 1 else "dbp100k"
dev = torch.device("cuda:0")
kg = make_kg_pair(shape)
adj_obj = DeviceAdjacency.from_triples(kg["n"], kg["triples"], device=dev)
adj = adj_obj.to_torch_coo(); _ = adj_obj.csr_t
x = torch.from_numpy(kg["x"]).to(dev)
torch.manual_seed(0)
model = UEAModel(bench.model_args(kg["n"], dev, 0)).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
bsz = 3000

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

def run():
    L = torch.randperm(kg["e1"], device=dev)[:bsz]; R = torch.randperm(kg["e2"], device=dev)[:bsz] + kg["e1"]
    t = [ev()]
    opt.zero_grad(set_to_none=True)
    emb = model.encode(x, adj); out = model.decode(emb, adj); t.append(ev())
    X, Y = out[L], out[R]; M = torch.cdist(X, Y, p=2); t.append(ev())
    a = torch.ones(bsz, device=dev)
    sinkhorn(a, a, M.detach(), 0.01, return_plan=False); t.append(ev())
    loss = torch.sum(M[:, 0].double()); loss.backward(); t.append(ev())
    opt.step(); t.append(ev())
    torch.cuda.synchronize()
    return [t[i].elapsed_time(t[i + 1]) for i in range(len(t) - 1)]

for _ in range(3): run()
r = np.mean([run() for _ in range(5)], 0)
print("fwd %.2f ms | gather+cdist %.2f | sinkhorn(1000) %.2f | bwd %.2f | adam %.2f | total %.2f" % (*r, r.sum()))
# isolated kernels
H = torch.randn(kg["n"], 300, device=dev)
def timeit(f, n=20):
    for _ in range(3): f()
    a = ev()
    for _ in range(n): f()
    b = ev(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
c = adj_obj.csr
byt = c.nnz * 8 + (c.n_rows + 1) * 4 + c.nnz * 300 * 4 + c.n_rows * 300 * 4
t = timeit(lambda: ops.spmm(c, H)); print("spmm plain   %.3f ms  %.0f GB/s (alg)" % (t, byt / t / 1e6))
g = torch.randn_like(H)
t = timeit(lambda: ops.spmm(c, H, _lib.ACT_RELU, g, H, True)); print("spmm fused+save %.3f ms  %.0f GB/s (alg)" % (t, (byt + 3 * c.n_rows * 1200) / t / 1e6))
W = torch.randn(300, 300, device=dev)
t = timeit(lambda: H @ W); print("gemm 200k x300x300 fp32 %.3f ms  %.1f TFLOP/s" % (t, 2 * kg["n"] * 300 * 300 / t / 1e9))
M = torch.cdist(H[:3000], H[3000:6000])
Mt = M.t().contiguous(); pot = torch.zeros(3000, device=dev)
t = timeit(lambda: ops.lse_dense(M, 100.0, pot), 200); print("lse_dense 3000^2 %.2f us" % (t * 1e3))
t = timeit(lambda: torch.cuda.current_stream().synchronize(), 50); print("sync %.2f us" % (t * 1e3))
import time
t0 = time.perf_counter(); sinkhorn(torch.ones(3000, device=dev), torch.ones(3000, device=dev), M, 0.01, return_plan=False); torch.cuda.synchronize(); print("sinkhorn wall %.2f ms" % ((time.perf_counter() - t0) * 1e3))
