"""Adjacency construction entry points of the reference's utils/data_utils.py
(get_matrix :296-321, get_sparse_tensor :325-336, get_sparse_tensor_for_one_graph
:339-350, sparse_mx_to_torch_sparse_tensor :51-57), executed on the GPU.

Only the adjacency/normalisation part of that file is on the hot path; the
DBP15K file loaders are out of scope.
"""
from __future__ import annotations

import numpy as np
import torch

from ..adjacency import DeviceAdjacency


def get_sparse_tensor(e, KG, device=None):
    """Degree-normalised adjacency of the union graph.  Returns a DeviceAdjacency
    (HBM-resident CSR); ``.tocoo()`` gives the scipy matrix the reference returns."""
    return DeviceAdjacency.from_triples(e, KG, device=device)


def get_sparse_tensor_for_one_graph(e, KG, index_R, device=None):
    """Same for one KG with ids remapped through ``index_R`` (:339-350)."""
    arr = np.asarray(KG, dtype=np.int64).reshape(-1, 3)
    if arr.size:
        lut = np.vectorize(index_R.__getitem__, otypes=[np.int64])
        arr = np.stack([lut(arr[:, 0]), arr[:, 1], lut(arr[:, 2])], 1)
    return DeviceAdjacency.from_triples(e, arr, device=device)


def get_matrix(e, KG, device=None):
    """(edge dict, degree dict) like the reference (:296-321), read back from the
    device-built adjacency.  Edge order is row-major, not dict-insertion order."""
    adj = DeviceAdjacency.from_triples(e, KG, device=device)
    crow = adj.crow.cpu().numpy()
    rows = np.repeat(np.arange(adj.n), np.diff(crow))
    cols = adj.col64.cpu().numpy()
    M = {(int(r), int(c)): 1 for r, c in zip(rows, cols)}
    arr = np.asarray(KG, dtype=np.int64).reshape(-1, 3)
    off = arr[:, 0] != arr[:, 2]
    cnt = np.bincount(arr[off, 0], minlength=e) + np.bincount(arr[off, 2], minlength=e)
    seen = np.zeros(e, dtype=bool)
    seen[arr[:, 0]] = True
    seen[arr[:, 2]] = True
    degree = {int(v): int(1 + cnt[v]) for v in np.nonzero(seen)[0]}
    return M, degree


def sparse_mx_to_torch_sparse_tensor(sparse_mx):
    """:51-57.  A DeviceAdjacency (or a scipy matrix made from one) becomes a
    coalesced CUDA sparse COO tensor with the kernel-format CSR attached; any
    other scipy matrix is converted on the host exactly as the reference does."""
    if isinstance(sparse_mx, DeviceAdjacency):
        return sparse_mx.to_torch_coo()
    attached = getattr(sparse_mx, "_eg_adj", None)
    if attached is not None:
        return attached.to_torch_coo()
    sparse_mx = sparse_mx.tocoo()
    indices = torch.from_numpy(np.vstack((sparse_mx.row, sparse_mx.col)).astype(np.int64))
    values = torch.from_numpy(sparse_mx.data.astype(np.float32))
    return torch.sparse_coo_tensor(indices, values, torch.Size(sparse_mx.shape))
