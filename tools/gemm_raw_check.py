"""NT GEMM from raw fp32 operands (in-kernel hi/lo split) against eg_split_tf32 + eg_gemm_nt_3xtf32: identical bits,
and the time of both (the split-operand time is quoted with and without its split launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
for m, k, n, n1 in ((130, 52, 340, 20), (257, 128, 128, 128), (1, 16, 4, 4), (4099, 300, 600, 300), (128 * 149, 300, 300, 300),
                    (777, 36, 260, 128), (200000, 300, 600, 300)):
    A = torch.randn(m, k, device=dev); B = torch.randn(n, k, device=dev) * 0.1; bias = torch.randn(n, device=dev)
    A2 = torch.randn(m, 36, device=dev); B2 = torch.randn(n, k + 36, device=dev) * 0.1
    add = torch.randn(m, n, device=dev)
    r0 = ops.gemm_nt([A], B, bias, n1=n1); r1 = ops.gemm_nt_raw([A], B, bias, n1=n1)
    c0 = torch.cat(r0, 1) if isinstance(r0, tuple) else r0
    c1 = torch.cat(r1, 1) if isinstance(r1, tuple) else r1
    d0 = ops.gemm_nt([A, A2], B2); d1 = ops.gemm_nt_raw([A, A2], B2)
    e1 = ops.gemm_nt_raw([A, A2], B2, addend=add)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t() + bias.double()
    print("m=%d k=%d n=%d: identical %s / %s, addend exact %s, err vs fp64 %.2e" %
          (m, k, n, torch.equal(c0, c1), torch.equal(d0, d1), torch.equal(e1, d0 + add),
           float((c1.double() - ref).abs().max() / ref.abs().max())), flush=True)
    assert torch.equal(c0, c1) and torch.equal(d0, d1) and torch.equal(e1, d0 + add)
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
sp = [ops.split_tf32(x, ops._pad16(k))]
for n in (600, 300):
    W = torch.randn(n, k, device=dev) / 17; b = torch.randn(n, device=dev)
    n1 = 300 if n == 600 else None
    t_split = timed(lambda: ops.gemm_nt([x], W, b, n1=n1, a_splits=sp))
    t_full = timed(lambda: ops.gemm_nt([x], W, b, n1=n1))
    t_raw = timed(lambda: ops.gemm_nt_raw([x], W, b, n1=n1))
    print("n=%d: split operands %.3f ms (%.3f ms with the split launches), raw operands %.3f ms" % (n, t_split, t_full, t_raw), flush=True)
dh = torch.randn(m, k, device=dev); dg = torch.randn(m, k, device=dev); add = torch.randn(m, k, device=dev)
Wd = torch.randn(300, 2 * k, device=dev) / 17
sps = [ops.split_tf32(dh, ops._pad16(k)), ops.split_tf32(dg, ops._pad16(k))]
t_split = timed(lambda: ops.gemm_nt([dh, dg], Wd, a_splits=sps))
t_raw = timed(lambda: ops.gemm_nt_raw([dh, dg], Wd))
t_raw_add = timed(lambda: ops.gemm_nt_raw([dh, dg], Wd, addend=add))
print("dx GEMM K=2x300 n=300: split operands %.3f ms, raw %.3f ms, raw + addend %.3f ms" % (t_split, t_raw, t_raw_add), flush=True)
