import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import ops
dev = torch.device("cuda:0"); torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
L = torch.randn(n, 300, device=dev); R = L + 0.5 * torch.randn(n, 300, device=dev)
for _ in range(2):
    ops.l1_ranks(L, R)
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); ops.l1_ranks(L, R); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
print("l1_ranks (streamed) %d^2 x300: %.3f ms, %.2f T fp64 op/s" % (n, ms, 2.0 * n * n * 300 / ms / 1e9))
