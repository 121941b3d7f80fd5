"""2-D tiled scaling Sinkhorn (default) vs the row-block kernel vs the log-domain kernel: agreement and time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import _lib
from gnn_mtl_b200.utils.ot_loss import sinkhorn
dev = torch.device("cuda:0"); torch.manual_seed(0)
dbg = _lib.lib.eg_debug_set
shapes = [(3000, 3000, 0.01, 1000), (3000, 3000, 0.05, 200), (2500, 3000, 0.02, 64), (1200, 1600, 0.05, 40), (3072, 3072, 0.01, 100),
          (700, 1000, 0.1, 30)]
for I, J, reg, iters in shapes:
    X = torch.randn(I, 300, device=dev) * 0.06; Y = torch.randn(J, 300, device=dev) * 0.06
    M = torch.cdist(X, Y); a = torch.ones(I, device=dev); b = torch.ones(J, device=dev) * (I / J)
    res = {}
    for name, knobs in (("tile2d", {}), ("rowblock", {12: 0}), ("log", {7: 0})):
        for k, v in knobs.items(): dbg(k, v)
        info = {}
        for _ in range(2):
            sinkhorn(a, b, M, reg, numItermax=iters, stopThr=-1.0, return_plan=False, info=info)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _, loss = sinkhorn(a, b, M, reg, numItermax=iters, stopThr=-1.0, return_plan=False, info=info)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
        res[name] = (info["log_u"].double(), info["log_v"].double(), float(loss), ms)
        for k in knobs: dbg(k, 1)
    ref = res["log"]
    line = "%dx%d reg %.2f %d sweeps:" % (I, J, reg, iters)
    for name in ("tile2d", "rowblock", "log"):
        r = res[name]
        d = (r[0][:, None] + r[1][None, :256]) - (ref[0][:, None] + ref[1][None, :256])
        line += "  %s %.2f ms (%.2f us/sweep) dpot %.1e dloss %.1e" % (name, r[3], r[3] * 1e3 / iters, float(d.abs().max()),
                                                                     abs(r[2] - ref[2]) / abs(ref[2]))
    print(line, "redos", dbg(8, 0), "folds", dbg(9, 0), flush=True)
