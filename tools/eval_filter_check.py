"""L1 rank kernels: fp32 candidate filter vs the all-fp64 streamed kernel — identical ranks, time, candidate share."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_mtl_b200 import ops
from gnn_mtl_b200.synth import make_kg_pair
dev = torch.device("cuda:0")
def timed(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = f()
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / reps
kg = make_kg_pair("dbp15k")
x = torch.from_numpy(kg["x"]).to(dev)
for name, pairs in (("dbp15k test 10500", kg["test"]), ("dbp15k train 4500", kg["train"])):
    t = torch.from_numpy(pairs).to(dev)
    L, R = x[t[:, 0]].contiguous(), x[t[:, 1]].contiguous()
    (rf, cf), ms_f = timed(lambda: ops.l1_ranks(L, R, filtered=True))
    (rx, cx), ms_x = timed(lambda: ops.l1_ranks(L, R, filtered=False))
    n = L.shape[0]
    print("%s: filtered %.3f ms  exact %.3f ms  identical %s  | fp32 lane-ops/s %.3e (of 3.7e13)" %
          (name, ms_f, ms_x, bool(torch.equal(rf, rx) and torch.equal(cf, cx)), 2.0 * n * n * 300 / ms_f * 1e3), flush=True)
g = torch.Generator(device=dev); g.manual_seed(0)
for n in (30000, 100000):
    X = torch.randn(n, 300, device=dev, generator=g) / 300 ** 0.5
    Y = X + 0.1 * torch.randn(n, 300, device=dev, generator=g) / 300 ** 0.5
    (rf, cf), ms_f = timed(lambda: ops.l1_ranks(X, Y, filtered=True), reps=2)
    (rx, cx), ms_x = timed(lambda: ops.l1_ranks(X, Y, filtered=False), reps=1)
    print("n=%d: filtered %.1f ms  exact %.1f ms  identical %s | fp32 lane-ops/s %.3e" %
          (n, ms_f, ms_x, bool(torch.equal(rf, rx) and torch.equal(cf, cx)), 2.0 * n * n * 300 / ms_f * 1e3), flush=True)
    Z = torch.randn(n, 300, device=dev, generator=g) / 300 ** 0.5        # unaligned: the diagonal sits in the bulk
    (rf, cf), ms_f = timed(lambda: ops.l1_ranks(X, Z, filtered=True), reps=2)
    (rx, cx), ms_x = timed(lambda: ops.l1_ranks(X, Z, filtered=False), reps=1)
    print("n=%d unaligned: filtered %.1f ms  exact %.1f ms  identical %s" %
          (n, ms_f, ms_x, bool(torch.equal(rf, rx) and torch.equal(cf, cx))), flush=True)
