import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import _lib, ops
dev = torch.device("cuda:0"); torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
X = torch.randn(n, 300, device=dev) * 0.06; Y = torch.randn(n, 300, device=dev) * 0.06
pot = torch.randn(n, device=dev)
A = ops.FusedOperand(X, 0, 1); B = ops.FusedOperand(Y, 0, 1)
for _ in range(3):
    ops.lse_fused(A, B, 0, 20.0, pot, None, 1)
torch.cuda.synchronize()
print("ok")
