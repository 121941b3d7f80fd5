"""Per-role wait / work cycles of the raw-operand NT GEMM (EG_GEMM_RAW_DEBUG): which stage of the TMA -> converter ->
MMA -> epilogue pipeline the k-block period comes from."""
import os, sys
os.environ["EG_GEMM_RAW_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import ops
dev = torch.device("cuda:0")
x = torch.randn(200000, 300, device=dev)
for n in (300, 600):
    W = torch.randn(n, 300, device=dev) / 17; b = torch.randn(n, device=dev)
    for _ in range(3):
        ops.gemm_nt_raw([x], W, b, n1=300 if n == 600 else None)
    torch.cuda.synchronize()
