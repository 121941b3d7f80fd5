"""Streamed vs stored-matrix L1 evaluation: get_hits ranks, get_neg top-125, per-row top-10 (CUDA events, median)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_mtl_b200 import ops

dev = torch.device("cuda:0")


def med(fn, n=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


g = torch.Generator(device=dev).manual_seed(0)
for n in (4500, 10500, 70000):
    L = torch.randn(n, 300, device=dev, generator=g)
    R = L + 0.5 * torch.randn(n, 300, device=dev, generator=g)
    ts, tm = med(lambda: ops.l1_ranks(L, R, streamed=True)), med(lambda: ops.l1_ranks(L, R, streamed=False), n=3)
    print("ranks n=%d: streamed %.2f ms (%.2f T DADD/s)  stored %.2f ms" % (n, ts, n * n * 300 / ts / 1e9, tm), flush=True)
vec = torch.randn(38960, 300, device=dev, generator=g)
anchors = vec[:4500]
ts, tm = med(lambda: ops.l1_topk(anchors, vec, 1, 125, streamed="force")), med(lambda: ops.l1_topk(anchors, vec, 1, 125, streamed=False), n=3)
print("get_neg top-125 4500 x 38960: streamed %.2f ms  stored %.2f ms" % (ts, tm), flush=True)
vec = torch.randn(200000, 300, device=dev, generator=g)
anchors = vec[:30000]
ts, tm = med(lambda: ops.l1_topk(anchors, vec, 1, 125, streamed="force"), n=3), med(lambda: ops.l1_topk(anchors, vec, 1, 125, streamed=False), n=2)
print("get_neg top-125 30000 x 200000: streamed %.2f ms (%.2f T DADD/s)  stored %.2f ms" % (ts, 30000 * 200000 * 300 / ts / 1e9, tm), flush=True)
ts, tm = med(lambda: ops.l1_topk(anchors, vec, 0, 10, streamed=True), n=3), med(lambda: ops.l1_topk(anchors, vec, 0, 10, streamed=False), n=2)
print("top-10 30000 x 200000: streamed %.2f ms (%.2f T DADD/s)  stored %.2f ms" % (ts, 30000 * 200000 * 300 / ts / 1e9, tm), flush=True)
