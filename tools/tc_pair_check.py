"""CTA-pair (cta_group::2) instantiations of the kernels in sinkhorn_tc.cu (debug knob 19) against the single-CTA ones:
fused LSE half-sweep, plan statistics, the NT GEMM and its short-chain variant — same results, and the time of both.
Run under `timeout` (a protocol slip hangs, it does not fail)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import _lib, ops
from gnn_mtl_b200._lib import lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
def pair(on): assert lib.eg_debug_set(19, int(on)) == 0
def timed(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
# --- fused LSE, several shapes (ragged rows / columns, odd tile counts)
for nA, nB in ((100, 300), (129, 257), (1000, 5000), (3001, 2999), (30000, 30000)):
    X = torch.randn(nA, 300, device=dev) * 0.06
    Y = torch.randn(nB, 300, device=dev) * 0.06
    if nA == nB: Y = X + 0.006 * torch.randn(nA, 300, device=dev)      # aligned pairs: the exact re-evaluation branch runs
    pot = torch.randn(nB, device=dev)
    res = []
    for on in (0, 1):
        pair(on)
        A = ops.FusedOperand(X, _lib.COST_L2, _lib.ALGO_TCGEN05); B = ops.FusedOperand(Y, _lib.COST_L2, _lib.ALGO_TCGEN05)
        out = ops.lse_fused(A, B, _lib.COST_L2, 20.0, pot, None, _lib.ALGO_TCGEN05)
        torch.cuda.synchronize()
        res.append(out)
    outs = [r if torch.is_tensor(r) else r[0] for r in res]
    ref = torch.logsumexp(pot[None, :].double() - 20.0 * torch.cdist(X.double(), Y.double()), 1) if nA * nB <= 2.5e7 else None
    print("lse %d x %d: max |pair - single| %.2e%s" % (nA, nB, float((outs[0] - outs[1]).abs().max()),
          "" if ref is None else ", vs fp64 %.2e / %.2e" % (float((outs[0].double() - ref).abs().max()), float((outs[1].double() - ref).abs().max()))), flush=True)
nf = 30000
X = torch.randn(nf, 300, device=dev) * 0.06; Y = X + 0.006 * torch.randn(nf, 300, device=dev); pot = torch.zeros(nf, device=dev)
for on in (0, 1):
    pair(on)
    A = ops.FusedOperand(X, _lib.COST_L2, _lib.ALGO_TCGEN05); B = ops.FusedOperand(Y, _lib.COST_L2, _lib.ALGO_TCGEN05)
    t = timed(lambda: ops.lse_fused(A, B, _lib.COST_L2, 20.0, pot, None, _lib.ALGO_TCGEN05), 5)
    print("lse 30000^2 pair=%d: %.3f ms  %.0f TF/s" % (on, t, 3 * 2.0 * nf * nf * 300 / t / 1e9), flush=True)
# --- GEMMs
m, k = 200000, 300
x = torch.randn(m, k, device=dev)
sp = [ops.split_tf32(x, ops._pad16(k))]
for n in (600, 300):
    W = torch.randn(n, k, device=dev) / 17; b = torch.randn(n, device=dev)
    n1 = 300 if n == 600 else None
    cat = lambda r: torch.cat(r, 1) if isinstance(r, tuple) else r
    for chained in (False, True):
        pair(0); r0 = cat(ops.gemm_nt([x], W, b, n1=n1, a_splits=sp, chained=chained)); t0 = timed(lambda: ops.gemm_nt([x], W, b, n1=n1, a_splits=sp, chained=chained))
        pair(1); r1 = cat(ops.gemm_nt([x], W, b, n1=n1, a_splits=sp, chained=chained)); t1 = timed(lambda: ops.gemm_nt([x], W, b, n1=n1, a_splits=sp, chained=chained))
        print("gemm_nt n=%d chained=%d: single %.3f ms, pair %.3f ms, identical %s (max diff %.2e)" % (n, chained, t0, t1, torch.equal(r0, r1), float((r0 - r1).abs().max())), flush=True)
for (mm, kk, nn) in ((130, 52, 340), (1, 16, 4), (4099, 300, 300), (128 * 149 + 5, 300, 600)):
    a = torch.randn(mm, kk, device=dev); w = torch.randn(nn, kk, device=dev) * 0.1; bb = torch.randn(nn, device=dev)
    for chained in (False, True):
        pair(0); r0 = ops.gemm_nt([a], w, bb, chained=chained)
        pair(1); r1 = ops.gemm_nt([a], w, bb, chained=chained)
        torch.cuda.synchronize()
        print("gemm_nt %dx%dx%d chained=%d: identical %s (max diff %.2e)" % (mm, kk, nn, chained, torch.equal(r0, r1), float((r0 - r1).abs().max())), flush=True)
pair(0)
