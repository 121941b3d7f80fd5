import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import ops
dev = torch.device("cuda:0"); torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10500
L = torch.randn(n, 300, device=dev); R = torch.randn(n, 300, device=dev)
for _ in range(3):
    D = ops.l1_matrix(L, R)
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); D = ops.l1_matrix(L, R); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
print("l1_matrix %d^2 x300: %.3f ms, %.2f T DADD/s" % (n, ms, 2.0 * n * n * 300 / ms / 1e9))
