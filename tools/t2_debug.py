"""Smallest possible exercise of the 2-D tiled Sinkhorn kernel: a few sweeps past the first marginal-error check."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import _lib
from gnn_mtl_b200.utils.ot_loss import sinkhorn
dev = torch.device("cuda:0"); torch.manual_seed(0)
X = torch.randn(3000, 300, device=dev) * 0.06; Y = torch.randn(3000, 300, device=dev) * 0.06
M = torch.cdist(X, Y); a = torch.ones(3000, device=dev)
for iters in (12, 13, 30, 200, 1000):
    info = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _, loss = sinkhorn(a, a, M, 0.01, numItermax=iters, stopThr=-1.0, return_plan=False, info=info)
    torch.cuda.synchronize()
    print("iters %d: %.2f ms loss %.6f redos %d" % (iters, (time.perf_counter() - t0) * 1e3, float(loss), _lib.lib.eg_debug_set(8, 0)), flush=True)
# with the stop rule active (host reads sweeps / err back)
for thr in (1e-9, 1e-3):
    info = {}
    _, loss = sinkhorn(a / 3000, a / 3000, M, 0.05, numItermax=300, stopThr=thr, return_plan=False, info=info)
    print("thr %g: sweeps %d err %.3e loss %.6f" % (thr, info["sweeps"], info["err"], float(loss)), flush=True)
