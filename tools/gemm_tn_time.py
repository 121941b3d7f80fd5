import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import ops
from gnn_mtl_b200.layers.layers import _weight_grad
dev = torch.device("cuda:0")
K, m, n = 200000, 300, 300
A = torch.randn(K, m, device=dev); B = torch.randn(K, n, device=dev)
sa, sb = ops.split_tf32(A, ops._pad16(m)), ops.split_tf32(B, ops._pad16(n))
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_tc = timed(lambda: ops.gemm_tn(sa, m, sb, n))
t_cb = timed(lambda: _weight_grad(A, B))
t_sp = timed(lambda: ops.split_tf32(A, ops._pad16(m)))
print("dW 200000x300x300: tcgen05 MN-major split-K %.3f ms (%.0f TF/s TF32-MMA incl. padding to 384x320; %.0f fp32-equivalent useful) | cuBLAS bmm split-K %.3f ms | one hi/lo split %.3f ms"
      % (t_tc, 3 * 2 * 384 * 320 * K / t_tc / 1e9, 2.0 * m * n * K / t_tc / 1e9, t_cb, t_sp))
