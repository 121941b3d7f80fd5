import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200.synth import make_kg_pair
from gnn_mtl_b200.utils.ot_loss import sinkhorn
dev = torch.device("cuda:0"); torch.manual_seed(0)
kg = make_kg_pair("dbp100k")
x = torch.from_numpy(kg["x"]).to(dev)
L = torch.randperm(kg["e1"], device=dev)[:3000]; R = torch.randperm(kg["e2"], device=dev)[:3000] + kg["e1"]
M = torch.cdist(x[L], x[R]); a = torch.ones(3000, device=dev)
info = {}
for _ in range(2): sinkhorn(a, a, M, 0.01, return_plan=False, info=info)
torch.cuda.synchronize(); t0 = time.perf_counter()
sinkhorn(a, a, M, 0.01, return_plan=False, info=info)
torch.cuda.synchronize(); print("sinkhorn real 3000^2: %d sweeps err %.3e: %.2f ms" % (info["sweeps"], info["err"], (time.perf_counter() - t0) * 1e3))
