"""CTA-pair (tcgen05.mma.cta_group::2) raw-operand NT GEMM (debug knob 18) against the single-SM kernels: identical
bits, and the time of both.  Small shapes first; run under `timeout` (a protocol slip hangs, it does not fail)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import ops
from gnn_mtl_b200._lib import lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
def pair(on): assert lib.eg_debug_set(18, int(on)) == 0
shapes = ((130, 52, 340, 20), (257, 128, 128, 128), (1, 16, 4, 4), (4099, 300, 600, 300), (128 * 149, 300, 300, 300),
          (777, 36, 260, 128), (200000, 300, 600, 300))
for m, k, n, n1 in shapes:
    A = torch.randn(m, k, device=dev); B = torch.randn(n, k, device=dev) * 0.1; bias = torch.randn(n, device=dev)
    A2 = torch.randn(m, 36, device=dev); B2 = torch.randn(n, k + 36, device=dev) * 0.1
    add = torch.randn(m, n, device=dev)
    pair(0)
    r0 = ops.gemm_nt_raw([A], B, bias, n1=n1); d0 = ops.gemm_nt_raw([A, A2], B2)
    pair(1)
    r1 = ops.gemm_nt_raw([A], B, bias, n1=n1); d1 = ops.gemm_nt_raw([A, A2], B2)
    e1 = ops.gemm_nt_raw([A, A2], B2, addend=add)
    torch.cuda.synchronize()
    c0 = torch.cat(r0, 1) if isinstance(r0, tuple) else r0
    c1 = torch.cat(r1, 1) if isinstance(r1, tuple) else r1
    ref = A.double() @ B.double().t() + bias.double()
    print("m=%d k=%d n=%d: identical %s / %s, addend exact %s, err vs fp64 %.2e, max |pair - single| %.2e" %
          (m, k, n, torch.equal(c0, c1), torch.equal(d0, d1), torch.equal(e1, d0 + add),
           float((c1.double() - ref).abs().max() / ref.abs().max()), float((c1 - c0).abs().max())), flush=True)
m, k = 200000, 300
x = torch.randn(m, k, device=dev)
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n in (600, 300):
    W = torch.randn(n, k, device=dev) / 17; b = torch.randn(n, device=dev)
    n1 = 300 if n == 600 else None
    pair(0); t1 = timed(lambda: ops.gemm_nt_raw([x], W, b, n1=n1))
    pair(1); t2 = timed(lambda: ops.gemm_nt_raw([x], W, b, n1=n1))
    print("n=%d: single-SM %.3f ms, CTA pair %.3f ms" % (n, t1, t2), flush=True)
dh = torch.randn(m, k, device=dev); dg = torch.randn(m, k, device=dev)
Wd = torch.randn(300, 2 * k, device=dev) / 17
pair(0); t1 = timed(lambda: ops.gemm_nt_raw([dh, dg], Wd))
pair(1); t2 = timed(lambda: ops.gemm_nt_raw([dh, dg], Wd))
print("dx GEMM K=2x300 n=300: single-SM %.3f ms, CTA pair %.3f ms" % (t1, t2), flush=True)
if os.environ.get("EG_GEMM_RAW_DEBUG"):
    pass
