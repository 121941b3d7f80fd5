"""Alignment evaluation entry points of utils/eval_utils.py (get_hits :71-98,
eval_gw_matching_matrix :133-159, eval_at_1 :161-168, format_metrics :7-9) on the
exact-fp64 L1 kernels.  Ranks follow the stable order (ties: lower index first);
NumPy's default argsort in the reference leaves ties unspecified.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


def format_metrics(metrics, split):
    return " ".join(["{}_{}: {:.4f}".format(split, name, val) for name, val in metrics.items()])


def _to_cuda(t, device=None):
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.asarray(t))
    if t.is_cuda:
        return t
    # the reference hands get_hits a CPU copy (models/models_ea.py:64-65); the kernels need HBM
    return t.to(device or "cuda")


def _pair_index(test_pair, device):
    pairs = torch.as_tensor(np.asarray(test_pair, dtype=np.int64).reshape(-1, 2), device=device)
    return pairs[:, 0].contiguous(), pairs[:, 1].contiguous()


def _hits_dict(rank_row, rank_col, top_k, n):
    ks = torch.as_tensor(list(top_k), device=rank_row.device, dtype=torch.int32)
    lr = (rank_row[None, :] < ks[:, None]).sum(1).tolist()
    rl = (rank_col[None, :] < ks[:, None]).sum(1).tolist()
    metrics = {}
    for k, c in zip(top_k, lr):
        metrics['Hits@{}_l'.format(k)] = c / n * 100
    for k, c in zip(top_k, rl):
        metrics['Hits@{}_r'.format(k)] = c / n * 100
    return metrics


def mean_reciprocal_rank(rank_row, rank_col):
    """MRR in both directions from the 0-based ranks (the reference has no MRR — BASELINE.json's north_star asks
    for it; it falls out of the same rank computation): mean of 1 / (rank + 1)."""
    return {"MRR_l": float((1.0 / (rank_row.double() + 1.0)).mean()), "MRR_r": float((1.0 / (rank_col.double() + 1.0)).mean())}


def get_hits(vec, test_pair, top_k=(1, 10, 50, 100), *, return_ranks=False, mrr=False):
    """Hits@k in both directions over the fp64 L1 matrix of the test pairs (keyword-only extras:
    ``mrr=True`` appends MRR_l / MRR_r, ``return_ranks=True`` also returns the rank tensors)."""
    vec = _to_cuda(vec.detach())
    left, right = _pair_index(test_pair, vec.device)
    n = int(left.numel())
    L = vec.index_select(0, left).to(torch.float32)
    R = vec.index_select(0, right).to(torch.float32)
    rank_row, rank_col = ops.l1_ranks(L, R)
    metrics = _hits_dict(rank_row, rank_col, top_k, n)
    if mrr:
        metrics.update(mean_reciprocal_rank(rank_row, rank_col))
    if return_ranks:
        return metrics, rank_row, rank_col
    return metrics


def eval_gw_matching_matrix(T, test_pair, index1_R, index2_R, top_k=(1, 10, 50, 100)):
    """Same ranking on the sub-block T[L][:, R] of a given score matrix (:133-159)."""
    T = _to_cuda(T.detach())
    rows = torch.as_tensor([index1_R[l] for l, r in test_pair], device=T.device)
    cols = torch.as_tensor([index2_R[r] for l, r in test_pair], device=T.device)
    sim = T.index_select(0, rows).index_select(1, cols)
    rank_row, rank_col = ops.matrix_ranks(sim)
    return _hits_dict(rank_row, rank_col, top_k, len(test_pair))


def eval_at_1(outputs, data):
    """Share of test pairs whose nearest right entity (L1) is the true match, in %."""
    outputs = _to_cuda(outputs.detach())
    left, right = _pair_index(data['test'], outputs.device)
    L = outputs.index_select(0, left).to(torch.float32)
    R = outputs.index_select(0, right).to(torch.float32)
    row_min, row_arg, _, _ = ops.l1_argmins(L, R)
    cnt = (row_arg == torch.arange(left.numel(), device=outputs.device)).to(torch.float32)
    return torch.sum(cnt) / len(cnt) * 100
