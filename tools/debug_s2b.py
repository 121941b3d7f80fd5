import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_mtl_b200 import ops
from gnn_mtl_b200.SinkhornOT import sinkhorn_iteration
g = np.load('tests/golden/sinkhorn.npz')
dev = torch.device('cuda:0')
C3 = torch.from_numpy(g['C_cos']).to(dev)
C = C3[0]
I, J = C.shape
mu = torch.full((I,), 1 / I, dtype=torch.float64, device=dev)
nu = torch.full((J,), 1 / J, dtype=torch.float64, device=dev)
inv = 100.0
Ct = ops.transpose(C)
al = torch.zeros(I, dtype=torch.float64, device=dev); be = torch.zeros(J, dtype=torch.float64, device=dev)
al_r, be_r = al.clone(), be.clone()
for ii in range(12):
    al, _ = ops.lse_dense(C, inv, be, torch.log(mu))
    be, _ = ops.lse_dense(Ct, inv, al, torch.log(nu))
    al_r = torch.log(mu) - torch.logsumexp(be_r[None, :] - C * inv, 1)
    be_r = torch.log(nu) - torch.logsumexp(al_r[None, :] - C.t() * inv, 1)
    print(ii, float((al - al_r).abs().max()), float((be - be_r).abs().max()))
w, k1, k2, K = sinkhorn_iteration(C3, mu.view(1, I, 1), nu.view(1, 1, J), 1e-2, numIterMax=100)
want = torch.from_numpy(g['S2_K_e2']).to(dev)
print('K err', float((K - want).abs().max() / want.max()), float(w), float(g['S2_w_e2']))
