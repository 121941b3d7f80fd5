import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200.utils.ot_loss import sinkhorn
dev = torch.device("cuda:0"); torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
X = torch.randn(n, 300, device=dev) * 0.06; Y = torch.randn(n, 300, device=dev) * 0.06
M = torch.cdist(X, Y); a = torch.ones(n, device=dev)
for _ in range(2):
    sinkhorn(a, a, M, 0.01, return_plan=False)
torch.cuda.synchronize(); t0 = time.perf_counter()
sinkhorn(a, a, M, 0.01, return_plan=False)
torch.cuda.synchronize(); print("sinkhorn %d^2 x1000 sweeps: %.2f ms" % (n, (time.perf_counter() - t0) * 1e3))
