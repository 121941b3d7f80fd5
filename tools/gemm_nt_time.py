"""Time / profile the NT 3xTF32 GEMM at the layer shapes: [200k, 300] x [600, 300]^T (hidden + gate) and n = 300."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import ops
dev = torch.device("cuda:0")
m, k = 200000, 300
x = torch.randn(m, k, device=dev)
sp = [ops.split_tf32(x, ops._pad16(k))]
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
    t = timed(lambda: ops.gemm_nt([x], W, b, n1=300 if n == 600 else None, a_splits=sp))
    print("gemm_nt m=%d k=%d n=%d: %.3f ms  (%.0f TF/s TF32-MMA, %.0f fp32-equivalent; writes %.0f MB)"
          % (m, k, n, t, 3 * 2.0 * m * n * 304 / t / 1e9, 2.0 * m * n * k / t / 1e9, m * n * 4 / 1e6), flush=True)

# short-chain (ReLU-feeding) variant and the cuBLAS fp32 product it replaces
for n in (600, 300):
    W = torch.randn(n, k, device=dev) / 17; b = torch.randn(n, device=dev)
    t = timed(lambda: ops.gemm_nt([x], W, b, n1=300 if n == 600 else None, a_splits=sp, chained=True))
    print("gemm_nt CHAINED m=%d k=%d n=%d: %.3f ms  (%.0f TF/s TF32-MMA)" % (m, k, n, t, 3 * 2.0 * m * n * 304 / t / 1e9), flush=True)
W = torch.randn(300, k, device=dev) / 17; b = torch.randn(300, device=dev)
torch.backends.cuda.matmul.allow_tf32 = False
t = timed(lambda: torch.mm(x, W.t()).add_(b))
print("cuBLAS fp32 mm + bias m=%d k=%d n=300: %.3f ms" % (m, k, t), flush=True)
