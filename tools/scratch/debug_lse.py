import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_mtl_b200 import _lib, ops
dev = torch.device("cuda:0")
n = 100_000
g = torch.Generator(device=dev); g.manual_seed(3)
X = torch.randn(n, 300, device=dev, generator=g) / 300 ** 0.5
Y = X[torch.randperm(n, device=dev, generator=g)] + 0.1 * torch.randn(n, 300, device=dev, generator=g) / 300 ** 0.5
B = ops.FusedOperand(Y, 0, 1)
for pot_kind in ("zeros", "randn", "spread"):
    pot = {"zeros": torch.zeros(n, device=dev), "randn": torch.randn(n, device=dev),
           "spread": torch.randn(n, device=dev) * 3 - 11.5}[pot_kind]
    for nA in (2000, 256, 20000):
        sl = ops.FusedOperand(X[:nA], 0, 1)
        _, l_tc = ops.lse_fused(sl, B, 0, 20.0, pot, None, 1, want_pot=False, want_lse=True)
        _, l_si = ops.lse_fused(sl, B, 0, 20.0, pot, None, 0, want_pot=False, want_lse=True)
        C = torch.cdist(X[:64].double(), Y.double())
        ref = torch.logsumexp(pot.double()[None, :] - C * 20.0, 1)
        print("pot %-6s nA %5d | tc-ref %.2e  simt-ref %.2e  tc-simt(all rows) %.2e" % (
            pot_kind, nA, float((l_tc[:64].double() - ref).abs().max()), float((l_si[:64].double() - ref).abs().max()),
            float((l_tc - l_si).abs().max())))
