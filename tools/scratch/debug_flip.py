import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import ea_oracle as orc
from gnn_mtl_b200.adjacency import DeviceAdjacency
import gnn_mtl_b200.layers.layers as LL
from gnn_mtl_b200.synth import make_kg_pair
dev = torch.device("cuda:0")
def rel(a, b): return float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max())
torch.manual_seed(1)
kg = make_kg_pair("dbp15k", dim=300)
x = torch.from_numpy(kg["x"])
acts = [F.relu, F.relu, (lambda z: z)]
layers = [LL.HighWayGraphConvolution(300, 300, 0.0, a, True, -1, "cpu") for a in acts]
params = [(l.linear.weight.detach(), l.linear.bias.detach(), l.kernel_gate, l.bias_gate) for l in layers]
tri = kg["triples"]
adjc = orc.adjacency_torch_coo(kg["n"], tri[:, 0], tri[:, 2])
xs = x.clone().requires_grad_(True)
ps = [tuple(p.clone().requires_grad_(i < 2) for i, p in enumerate(q)) for q in params]
# oracle forward keeping layer-1 pre-activation
h = xs; pre = []
for (w, b, g, c), a in zip(ps, ["relu", "relu", "identity"]):
    S = torch.sparse.mm(adjc, F.linear(h, w, b)); pre.append(S.detach())
    h = orc.highway_layer(h, adjc, w, b, g, c, a)
seed = torch.randn_like(h); (h * seed).sum().backward()
adj = DeviceAdjacency.from_triples(kg["n"], kg["triples"], device=dev).to_torch_coo()
for l in layers: l.to(dev)
for flag in (False, True):
    LL.USE_TCGEN05_GEMM = flag
    for l in layers: l.zero_grad()
    xg = x.to(dev).requires_grad_(True)
    hg = xg
    for l in layers: hg, _ = l((hg, adj))
    (hg * seed.to(dev)).sum().backward()
    print("tcgen05 gemm %s | out %.2e | dx %.2e | dW0 %.2e" % (flag, rel(hg.detach(), h.detach()), rel(xg.grad, xs.grad), rel(layers[0].linear.weight.grad, ps[0][0].grad)))
    d = (xg.grad.cpu() - xs.grad).abs()
    print("   dx abs err: max %.2e, 99.99pct %.2e, median %.2e ; max|dx| %.2e" % (float(d.max()), float(d.flatten().kthvalue(int(d.numel() * 0.9999)).values), float(d.median()), float(xs.grad.abs().max())))
print("smallest |S| in relu layers (oracle):", [float(p.abs().min()) for p in pre[:2]], "count |S|<1e-5:", [int((p.abs() < 1e-5).sum()) for p in pre[:2]])
