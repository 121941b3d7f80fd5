"""One get_hits-sized rank evaluation (10,500 pairs, d = 300) through the fp32-filter kernels: the ncu target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import ops
from gnn_mtl_b200.synth import make_kg_pair
dev = torch.device("cuda:0")
kg = make_kg_pair("dbp15k")
x = torch.from_numpy(kg["x"]).to(dev)
t = torch.from_numpy(kg["test"]).to(dev)
L, R = x[t[:, 0]].contiguous(), x[t[:, 1]].contiguous()
for _ in range(3):
    r, c = ops.l1_ranks(L, R, filtered=True)
torch.cuda.synchronize()
print("hits@1 l %.2f" % (100.0 * float((r == 0).float().mean())))
