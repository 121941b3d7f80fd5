"""Scaling-domain on-chip Sinkhorn vs the log-domain kernel: potentials, loss, sweeps, time (3000 x 3000, reg 0.01)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import _lib
from gnn_mtl_b200.synth import make_kg_pair
from gnn_mtl_b200.utils.ot_loss import sinkhorn
dev = torch.device("cuda:0"); torch.manual_seed(0)
kg = make_kg_pair("dbp100k")
x = torch.from_numpy(kg["x"]).to(dev)
L = torch.randperm(kg["e1"], device=dev)[:3000]; R = torch.randperm(kg["e2"], device=dev)[:3000] + kg["e1"]
cases = {"embeddings cdist, a=b=1": (torch.cdist(x[L], x[R]), torch.ones(3000, device=dev), 0.01, 1000, -1.0),
         "uniform random cost [0,2], a=b=1/n": (2 * torch.rand(3000, 3000, device=dev), torch.ones(3000, device=dev) / 3000, 0.01, 1000, -1.0),
         "random cost, reg 0.05, stop 1e-9": (torch.rand(2900, 3000, device=dev), torch.ones(3000, device=dev) / 3000, 0.05, 1000, 1e-9),
         "clustered cost (outlier columns)": (torch.cat([torch.rand(3000, 2900, device=dev), 3 + torch.rand(3000, 100, device=dev)], 1), torch.ones(3000, device=dev) / 3000, 0.01, 300, -1.0)}
for name, (M, w, reg, iters, thr) in cases.items():
    a = w[:M.shape[0]] * (w.sum() / w[:M.shape[0]].sum())
    out = {}
    for mode in (0, 1):
        _lib.lib.eg_debug_set(7, mode)
        info = {}
        for _ in range(2):
            sinkhorn(a, w, M, reg, numItermax=iters, stopThr=thr, return_plan=False, info=info)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _, loss = sinkhorn(a, w, M, reg, numItermax=iters, stopThr=thr, return_plan=False, info=info)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
        out[mode] = (info["log_u"].double(), info["log_v"].double(), float(loss), info["sweeps"], info["err"], ms)
    d_u = float((out[0][0] - out[1][0]).abs().max()); d_v = float((out[0][1] - out[1][1]).abs().max())
    # the potentials are defined up to a constant shift (u*c, v/c): compare the shift-invariant sum too
    shift = float(((out[0][0] - out[1][0]).mean() + (out[0][1] - out[1][1]).mean()))
    print("%s: log %d sweeps %.2f ms loss %.9g err %.3e | scaling %d sweeps %.2f ms loss %.9g err %.3e | max|dlog u| %.2e max|dlog v| %.2e (mean shift sum %.2e) rel dloss %.2e | fallbacks %d absorbs %d"
          % (name, out[0][3], out[0][5], out[0][2], out[0][4], out[1][3], out[1][5], out[1][2], out[1][4], d_u, d_v, shift,
             abs(out[0][2] - out[1][2]) / abs(out[0][2]), _lib.lib.eg_debug_set(8, 0), _lib.lib.eg_debug_set(9, 0)), flush=True)
