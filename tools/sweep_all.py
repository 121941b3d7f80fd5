"""BASELINE.json configs 4/5 + eval: SpMM sweep on power-law graphs, large fused Sinkhorn passes, L1 eval.
Writes gpurun_out/sweeps.json.  Timing: CUDA events, median of n after warm-up, L2 flushed between SpMM runs."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_mtl_b200 import _lib, ops
from gnn_mtl_b200.adjacency import DeviceAdjacency
from gnn_mtl_b200.synth import make_powerlaw_graph, make_kg_pair
dev = torch.device("cuda:0")
out = {"spmm": [], "sinkhorn_fused": [], "eval": []}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def med(f, n=7, fl=True):
    for _ in range(2): f()
    ts = []
    for _ in range(n):
        if fl: flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
cfgs = [(1_000_000, 5), (1_000_000, 20), (1_000_000, 50), (3_000_000, 20)] + ([] if quick else [(10_000_000, 5), (10_000_000, 20)])
for n, deg in cfgs:
    t0 = time.time()
    h, t = make_powerlaw_graph(n, deg, seed=1)
    adj = DeviceAdjacency.from_heads_tails(n, torch.from_numpy(h).to(dev), torch.from_numpy(t).to(dev))
    c, ct = adj.csr, adj.csr_t
    build_s = time.time() - t0
    for d in (128, 300):
        H = torch.randn(n, d, device=dev)
        byt = c.nnz * 8 + (n + 1) * 4 + c.nnz * d * 4 + n * d * 4
        tf = med(lambda: ops.spmm(c, H)); tb = med(lambda: ops.spmm(ct, H))
        g = torch.randn_like(H)
        tfu = med(lambda: ops.spmm(c, H, _lib.ACT_RELU, g, H, True))
        rec = {"n": n, "avg_degree_target": deg, "nnz": c.nnz, "max_row": int((c.rowptr[1:] - c.rowptr[:-1]).max()),
               "n_long_rows": c.n_long, "d": d, "fwd_ms": tf, "bwd_ms": tb, "fused_fwd_ms": tfu,
               "fwd_gbs": byt / tf / 1e6, "bwd_gbs": byt / tb / 1e6, "fused_gbs": (byt + 3 * n * d * 4) / tfu / 1e6,
               "graph_plus_csr_build_s": build_s}
        out["spmm"].append(rec); print(rec, flush=True)
        del H, g
    del adj, c, ct
    torch.cuda.empty_cache()
# fused Sinkhorn half-sweeps (tcgen05), cost never materialised
for n in ([100_000, 300_000] if quick else [100_000, 300_000, 1_000_000]):
    X = torch.randn(n, 300, device=dev) / 300 ** 0.5
    Y = X[torch.randperm(n, device=dev)] + 0.1 * torch.randn(n, 300, device=dev) / 300 ** 0.5
    A = ops.FusedOperand(X, _lib.COST_L2, _lib.ALGO_TCGEN05); B = ops.FusedOperand(Y, _lib.COST_L2, _lib.ALGO_TCGEN05)
    pot = torch.zeros(n, device=dev)
    tm = med(lambda: ops.lse_fused(A, B, _lib.COST_L2, 20.0, pot, None, _lib.ALGO_TCGEN05), n=3, fl=False)
    rec = {"I": n, "J": n, "d": 300, "half_sweep_ms": tm, "tf32_mma_tflops": 3 * 2.0 * n * n * 300 / tm / 1e9,
           "fp32_equiv_tflops": 2.0 * n * n * 300 / tm / 1e9, "sweeps50_s": 2 * 50 * tm / 1e3}
    out["sinkhorn_fused"].append(rec); print(rec, flush=True)
    del X, Y, A, B
    torch.cuda.empty_cache()
# eval: exact fp64 L1 + ranks
kg = make_kg_pair("dbp15k")
vec = torch.from_numpy(kg["x"]).to(dev)
for npairs in (4500, 10500):
    L = vec[torch.from_numpy(kg["test"][:npairs, 0]).to(dev)]; R = vec[torch.from_numpy(kg["test"][:npairs, 1]).to(dev)]
    tm = med(lambda: ops.l1_ranks(L, R), n=5, fl=False)
    rec = {"op": "get_hits ranks", "N": npairs, "d": 300, "ms": tm, "dadd_per_s": 2.0 * npairs * npairs * 300 / tm / 1e-3 / 1e12}
    out["eval"].append(rec); print(rec, flush=True)
anchors = vec[torch.from_numpy(kg["train"][:, 0]).to(dev)]
tm = med(lambda: ops.l1_topk(anchors, vec, 1, 125), n=3, fl=False)
out["eval"].append({"op": "get_neg top-125", "t": int(anchors.shape[0]), "n": int(vec.shape[0]), "ms": tm}); print(out["eval"][-1], flush=True)
L = torch.randn(70000, 300, device=dev); R = torch.randn(70000, 300, device=dev)
tm = med(lambda: ops.l1_ranks(L, R), n=2, fl=False)
out["eval"].append({"op": "get_hits ranks", "N": 70000, "d": 300, "ms": tm, "dadd_per_s": 2.0 * 70000 ** 2 * 300 / tm / 1e-3 / 1e12}); print(out["eval"][-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweeps.json", "w"), indent=1)
