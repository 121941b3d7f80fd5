import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import ops
dev = torch.device("cuda:0")
for K, m, n in [(1, 4, 4), (8, 32, 32), (16, 128, 160), (37, 300, 300), (5000, 300, 300), (200000, 300, 300)]:
    gen = torch.Generator().manual_seed(K + m)
    A = torch.randn(K, m, generator=gen); B = torch.randn(K, n, generator=gen)
    want = A.double().t() @ B.double()
    sa, sb = ops.split_tf32(A.to(dev), ops._pad16(m)), ops.split_tf32(B.to(dev), ops._pad16(n))
    got = ops.gemm_tn(sa, m, sb, n).double().cpu()
    err = (got - want).abs()
    print(K, m, n, "max err %.3e scale %.3e nonzero frac %.3f" % (float(err.max()), float(want.abs().max()), float((got != 0).float().mean())), flush=True)
    if K <= 16 and float(err.max()) > 1e-3:
        # which (k, m, n) pairing does the hardware see?  probe with one-hot operands
        for (ka, ma, kb, nb) in [(0, 0, 0, 0), (0, 1, 0, 0), (0, 0, 0, 1), (1 % K, 0, 1 % K, 0), (0, 5 % m, 0, 3 % n)]:
            A1 = torch.zeros(K, m); B1 = torch.zeros(K, n); A1[ka, ma] = 1.0; B1[kb, nb] = 2.0
            g = ops.gemm_tn(ops.split_tf32(A1.to(dev), ops._pad16(m)), m, ops.split_tf32(B1.to(dev), ops._pad16(n)), n).cpu()
            nz = torch.nonzero(g)
            print("   one-hot A[%d,%d] B[%d,%d] -> nonzeros %s values %s" % (ka, ma, kb, nb, nz.tolist()[:6], g[g != 0].tolist()[:6]), flush=True)
