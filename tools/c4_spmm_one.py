"""One configuration of the BASELINE.json config-4 SpMM sweep (power-law graph, forward launches), for ncu:
   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:spmm_vec -s 2 -c 1 ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import ops
from gnn_mtl_b200.adjacency import DeviceAdjacency
from gnn_mtl_b200.synth import make_powerlaw_graph
n, deg = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
h, t = make_powerlaw_graph(n, deg, seed=1)
full = DeviceAdjacency.from_heads_tails(n, torch.from_numpy(h).to(dev), torch.from_numpy(t).to(dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for d in (128, 300):
    H = torch.randn(n, d, device=dev)
    for _ in range(3):
        flush.zero_()
        ops.spmm(full.csr, H)
    torch.cuda.synchronize()
    del H
print("done n=%d nnz=%d" % (n, full.nnz))
