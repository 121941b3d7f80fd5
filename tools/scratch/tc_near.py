import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import _lib, ops
dev = torch.device("cuda:0")
n = 100_000
g = torch.Generator(device=dev); g.manual_seed(3)
X = torch.randn(n, 300, device=dev, generator=g) / 300 ** 0.5
Y = X[torch.randperm(n, device=dev, generator=g)] + 0.1 * torch.randn(n, 300, device=dev, generator=g) / 300 ** 0.5
A = ops.FusedOperand(X, 0, 1); B = ops.FusedOperand(Y, 0, 1)
pot = torch.zeros(n, device=dev)
for _ in range(2): ops.lse_fused(A, B, 0, 20.0, pot, None, 1)
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): ops.lse_fused(A, B, 0, 20.0, pot, None, 1)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print("aligned-pair data (one close pair per row): %.2f ms per half-sweep, %.1f TF/s tf32-mma" % (ms, 3 * 2.0 * n * n * 300 / ms / 1e9))
