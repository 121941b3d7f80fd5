"""Which rows of the CTA-pair raw-operand GEMM differ from the single-SM kernel, and why: for every wrong row the 160
wrong outputs are solved (least squares per k-block) for the 16 A values that would produce them, and those values are
searched for in A — this is how the early release of the raw A stage was identified (the stale values were the same
row's k-block ten stages ahead = the ring depth)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import ops
from gnn_mtl_b200._lib import lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
m, n, k = 200000, 160, 300
A = torch.randn(m, k, device=dev); B = torch.randn(n, k, device=dev) * 0.1
lib.eg_debug_set(18, 0); r0 = ops.gemm_nt_raw([A], B)
for rep in range(3):
    lib.eg_debug_set(18, 1); r1 = torch.zeros_like(r0); r1 = ops.gemm_nt_raw([A], B)
    torch.cuda.synchronize()
    bad = (r0 != r1)
    rows = bad.any(1).nonzero().reshape(-1)
    print("rep %d: %d bad rows: %s" % (rep, len(rows), [(int(r) // 256, (int(r) // 128) & 1, int(r) % 128) for r in rows.tolist()]), flush=True)
    Ap = torch.nn.functional.pad(A, (0, 4)).double(); Bp = torch.nn.functional.pad(B, (0, 4)).double()
    for r in rows.tolist()[:12]:
        diff = (r1[r] - r0[r]).double()                     # [160]
        best = None
        for kb in range(19):
            Bk = Bp[:, kb * 16:(kb + 1) * 16]               # [160,16]
            sol = torch.linalg.lstsq(Bk, diff.unsqueeze(1)).solution.squeeze(1)
            res = float((Bk @ sol - diff).norm() / diff.norm())
            if best is None or res < best[0]: best = (res, kb, sol)
        res, kb, delta = best
        stale = Ap[r, kb * 16:(kb + 1) * 16] + delta
        # where does the stale k-block come from?  search same kb column block across rows, and other kbs of nearby rows
        src = None
        for kb2 in range(19):
            d = (Ap[:, kb2 * 16:(kb2 + 1) * 16] - stale).abs().max(1).values
            j = int(d.argmin())
            if float(d[j]) < 1e-3: src = (j, kb2, float(d[j])); break
        print("  row %d (pair tile %d, cta %d, lane %d): best kb %d residual %.2e, stale source (row, kb, err) %s, |stale| %.3f"
              % (r, r // 256, (r // 128) & 1, r % 128, kb, res, src, float(stale.abs().max())), flush=True)
