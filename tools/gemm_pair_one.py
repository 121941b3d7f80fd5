"""One shape of the CTA-pair raw-operand NT GEMM ([200k, 300] x [600, 300]^T, two outputs) for an ncu capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import ops
dev = torch.device("cuda:0")
x = torch.randn(200000, 300, device=dev); W = torch.randn(600, 300, device=dev) / 17; b = torch.randn(600, device=dev)
for _ in range(4):
    ops.gemm_nt_raw([x], W, b, n1=300)
torch.cuda.synchronize()
