"""x·Wᵀ+b for the ReLU layers ([200k, 300] x [300, 300]): cuBLAS entry points vs the tcgen05 3xTF32 kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from gnn_mtl_b200 import ops
dev = torch.device("cuda:0")
n, d = 200000, 300
x = torch.randn(n, d, device=dev); W = torch.randn(d, d, device=dev) / 17; b = torch.randn(d, device=dev)
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ref = (x.double() @ W.double().t() + b.double())
Wt = W.t().contiguous()
out = torch.empty(n, d, device=dev)
variants = {
    "F.linear(x, W, b)": lambda: F.linear(x, W, b),
    "x @ W.t()": lambda: x @ W.t(),
    "x @ Wt (pre-transposed)": lambda: x @ Wt,
    "torch.addmm(b, x, W.t())": lambda: torch.addmm(b, x, W.t()),
    "mm(out=) then add_(b)": lambda: torch.mm(x, W.t(), out=out).add_(b),
    "tcgen05 gemm_nt([x], W, b)": lambda: ops.gemm_nt([x], W, b),
}
for name, fn in variants.items():
    y = fn()
    if "W.t()" in name and "addmm" not in name and "linear" not in name and "add_" not in name or "Wt" in name:
        y = y + b
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    print("%-32s %.3f ms   max rel err %.2e" % (name, timed(fn), err), flush=True)
