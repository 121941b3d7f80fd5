import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_mtl_b200 import ops
dev = torch.device("cuda:0"); rng = np.random.default_rng(0)
n, d, t, k = 38960, 300, 4500, 125
x = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32) * 0.05).to(dev).requires_grad_(True)
ILL = np.stack([rng.permutation(19000)[:t], rng.permutation(19000)[:t] + 19388], 1)
ix = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int64), device=dev)
L, R = ix(ILL[:, 0]), ix(ILL[:, 1])
nl, n2r = ix(np.repeat(ILL[:, 0], k)), ix(np.repeat(ILL[:, 1], k))
nr, n2l = ix(rng.integers(0, n, t * k)), ix(rng.integers(0, n, t * k))
def run():
    x.grad = None
    loss = ops.margin_loss(x, L, R, nl, nr, n2l, n2r, k); loss.backward(); return loss
def ref():
    x.grad = None
    A = (x[L] - x[R]).abs().sum(1); D = (A + 1.0).reshape(t, 1)
    B1 = (x[nl] - x[nr]).abs().sum(1).reshape(t, k); B2 = (x[n2l] - x[n2r]).abs().sum(1).reshape(t, k)
    loss = (torch.relu(D - B1).sum() + torch.relu(D - B2).sum()) / (2.0 * t * k); loss.backward(); return loss
for f, name in ((run, "fused kernel"), (ref, "torch ops on the same GPU")):
    for _ in range(2): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): l = f()
    torch.cuda.synchronize(); print("%s: fwd+bwd %.2f ms, loss %.6f" % (name, (time.perf_counter() - t0) / 5 * 1e3, float(l)))
