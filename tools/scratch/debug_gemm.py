import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import ops
dev = torch.device("cuda:0")
def rel(a, b): return float((a.double() - b).abs().max() / b.abs().max())
for m, k, n, k2 in [(130, 300, 600, 36), (130, 300, 600, 300), (3000, 300, 300, 300), (777, 52, 260, 36), (38960, 300, 300, 300), (777, 304, 260, 48), (777, 300, 260, 0)]:
    torch.manual_seed(1)
    A = torch.randn(m, k, device=dev); B = torch.randn(n, k, device=dev) * 0.1; bias = torch.randn(n, device=dev)
    got = ops.gemm_nt([A], B, bias)
    e1 = rel(got, A.double() @ B.double().t() + bias.double())
    e2 = float('nan')
    if k2:
        A2 = torch.randn(m, k2, device=dev); B2 = torch.randn(n, k + k2, device=dev) * 0.1
        got2 = ops.gemm_nt([A, A2], B2)
        ref2 = torch.cat([A, A2], 1).double() @ B2.double().t()
        e2 = rel(got2, ref2)
        # which K part is wrong?
        g1 = ops.gemm_nt([A, torch.zeros_like(A2)], B2); r1 = A.double() @ B2[:, :k].double().t()
        g2 = ops.gemm_nt([torch.zeros_like(A), A2], B2); r2 = A2.double() @ B2[:, k:].double().t()
        print("   part1 %.2e part2 %.2e" % (rel(g1, r1), rel(g2, r2)))
    print("m %d k %d n %d k2 %d | single %.2e concat %.2e" % (m, k, n, k2, e1, e2))
