"""tcgen05 fused LSE vs the SIMT fused LSE and vs torch fp64 (diagnostic)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import _lib, ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
def run(nA, nB, d, cost, scale=0.06, inv_reg=20.0, timing=False):
    X = torch.randn(nA, d, device=dev) * scale
    Y = torch.randn(nB, d, device=dev) * scale
    pot = torch.randn(nB, device=dev)
    A_t = ops.FusedOperand(X, cost, _lib.ALGO_TCGEN05); B_t = ops.FusedOperand(Y, cost, _lib.ALGO_TCGEN05)
    _, l_tc = ops.lse_fused(A_t, B_t, cost, inv_reg, pot, None, _lib.ALGO_TCGEN05, want_pot=False, want_lse=True)
    torch.cuda.synchronize()
    _, l_simt = ops.lse_fused(A_t, B_t, cost, inv_reg, pot, None, _lib.ALGO_SIMT, want_pot=False, want_lse=True)
    if nA * nB <= 4e7:
        Xd, Yd = X.double(), Y.double()
        if cost == _lib.COST_COSINE:
            C = 1 - (Xd / Xd.norm(dim=1, keepdim=True)) @ (Yd / Yd.norm(dim=1, keepdim=True)).t()
        else:
            C = torch.cdist(Xd, Yd)
            if cost == _lib.COST_SQEUCLID: C = C * C
        ref = torch.logsumexp(pot.double()[None, :] - C * inv_reg, 1)
        e_tc = float((l_tc.double() - ref).abs().max()); e_simt = float((l_simt.double() - ref).abs().max())
    else:
        e_tc = float((l_tc - l_simt).abs().max()); e_simt = float('nan')
    msg = "nA %6d nB %6d d %3d cost %d | abs err tc %.2e simt %.2e" % (nA, nB, d, cost, e_tc, e_simt)
    if timing:
        def t(algo, n=5):
            for _ in range(2): ops.lse_fused(A_t, B_t, cost, inv_reg, pot, None, algo)
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n): ops.lse_fused(A_t, B_t, cost, inv_reg, pot, None, algo)
            b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
        t_tc, t_si = t(_lib.ALGO_TCGEN05), t(_lib.ALGO_SIMT)
        fl = 2.0 * nA * nB * d
        msg += " | tc %.3f ms (%.1f TF fp32-eq, %.1f TF tf32-mma) simt %.3f ms (%.1f TF)" % (
            t_tc, fl / t_tc / 1e9, 3 * fl / t_tc / 1e9, t_si, fl / t_si / 1e9)
    print(msg, flush=True)
for args in [(128, 256, 32, 0), (128, 256, 300, 0), (333, 270, 300, 0), (333, 270, 300, 1), (333, 270, 300, 2),
             (1000, 3000, 300, 0), (3000, 3000, 128, 0), (5000, 7000, 300, 0)]:
    run(*args)
run(20000, 20000, 300, 0, timing=True)
run(100000, 100000, 300, 0, timing=True)
