import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_mtl_b200 import ops
g = np.load('tests/golden/sinkhorn.npz')
dev = torch.device('cuda:0')
C = torch.from_numpy(g['C_cos'])[0].to(dev)
I, J = C.shape
inv = 100.0
beta = torch.randn(J, dtype=torch.float64, device=dev)
alpha = torch.randn(I, dtype=torch.float64, device=dev)
lw = torch.randn(I, dtype=torch.float64, device=dev)
p, l = ops.lse_dense(C, inv, beta, lw, want_lse=True)
ref = torch.logsumexp(beta[None, :] - C * inv, 1)
print('lse row err', float((l - ref).abs().max()), 'pot err', float((p - (lw - ref)).abs().max()))
Ct = ops.transpose(C)
print('transpose err', float((Ct - C.t()).abs().max()))
lw2 = torch.randn(J, dtype=torch.float64, device=dev)
p2, l2 = ops.lse_dense(Ct, inv, alpha, lw2, want_lse=True)
ref2 = torch.logsumexp(alpha[None, :] - C.t() * inv, 1)
print('lse col err', float((l2 - ref2).abs().max()))
P, loss, rs, cs = ops.plan_dense(C, inv, alpha, beta, want_plan=True, want_rows=True, want_cols=True)
Pr = torch.exp(alpha[:, None] + beta[None, :] - C * inv)
print('plan err', float((P - Pr).abs().max() / Pr.max()), 'loss', float(loss), float((Pr * C).sum()),
      'rs', float((rs - Pr.sum(1)).abs().max() / Pr.sum(1).max()), 'cs', float((cs - Pr.sum(0)).abs().max() / Pr.sum(0).max()))
_, loss2, _, _ = ops.plan_dense(C, inv, alpha, beta, want_plan=False)
print('loss noplan', float(loss2))
