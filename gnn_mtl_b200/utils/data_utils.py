"""Adjacency construction entry points of the reference's utils/data_utils.py
(get_matrix :296-321, get_sparse_tensor :325-336, get_sparse_tensor_for_one_graph
:339-350, sparse_mx_to_torch_sparse_tensor :51-57), executed on the GPU.

The adjacency/normalisation part of that file is on the hot path.  The DBP15K
on-disk format (ent_ids_{1,2}, rel_ids_{1,2}, triples_{1,2},
ref_ent_ids, ref_r_ids, <lang>_vectorList.json) is parsed on the host by
read_dbp15k() and fed straight to the device CSR builder by load_data_ea() /
load_seperate_data_ea(), which keep the reference's dict keys (:375-455).
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch
import torch.nn.functional as F

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


# --------------------------------------------------------------------------- #
# DBP15K on-disk format -> device tensors (utils/data_utils.py:362-455)        #
# --------------------------------------------------------------------------- #

DATA_ROOT = "data/dbp15k"


def loadfile(fn, num=1):
    """Tab-separated integer columns -> list of tuples (:362-372)."""
    rows = []
    with open(fn, encoding="utf-8") as f:
        for line in f:
            cells = line.rstrip("\n").split("\t")
            rows.append(tuple(int(c) for c in cells[:num]))
    return rows


def read_dbp15k(lang, root=DATA_ROOT):
    """Host-side parse of one language pair directory; no device work.  Returns a dict of plain
    Python / NumPy objects: ent1, ent2, rel1, rel2 (id lists), kg1, kg2 (triples), ill, ill_r, x."""
    d = os.path.join(root, lang)
    col0 = lambda name: [r[0] for r in loadfile(os.path.join(d, name), 1)]
    with open(os.path.join(d, lang[0:2] + "_vectorList.json"), encoding="utf-8") as f:
        vectors = torch.tensor(json.load(f), dtype=torch.float32)
    return {"ent1": col0("ent_ids_1"), "ent2": col0("ent_ids_2"), "rel1": col0("rel_ids_1"), "rel2": col0("rel_ids_2"),
            "kg1": loadfile(os.path.join(d, "triples_1"), 3), "kg2": loadfile(os.path.join(d, "triples_2"), 3),
            "ill": loadfile(os.path.join(d, "ref_ent_ids"), 2), "ill_r": loadfile(os.path.join(d, "ref_r_ids"), 2),
            "x": F.normalize(vectors, 2, 1)}                       # get_features (:352-358)


def _split_links(ill):
    """Shuffle with the GLOBAL NumPy RNG, first len//10*3 links train, the rest test (:389-393)."""
    ill = list(ill)
    np.random.shuffle(ill)
    cut = len(ill) // 10 * 3
    return np.array(ill[:cut]), np.array(ill[cut:])


def rfunc(e, KG):
    """Relation bookkeeping of :272-292.  head/tail: relation -> entity lists (first-seen relation order);
    head_r/tail_r: [e, r] 0/1 indicator matrices, returned as scipy CSR instead of dense float64 arrays
    (38,960 x 3,024 dense is 0.9 GB each and nothing on the EA path reads them)."""
    import scipy.sparse as sp
    head, tail = {}, {}
    for h, r, t in KG:
        head.setdefault(r, []).append(h)
        tail.setdefault(r, []).append(t)
    arr = np.asarray(KG, dtype=np.int64).reshape(-1, 3)
    n_rel = len(head)
    ones = np.ones(len(arr))
    head_r = sp.csr_matrix((ones, (arr[:, 0], arr[:, 1])), shape=(e, n_rel))
    tail_r = sp.csr_matrix((ones, (arr[:, 2], arr[:, 1])), shape=(e, n_rel))
    head_r.data[:] = 1
    tail_r.data[:] = 1
    return head, tail, head_r, tail_r


def load_data_ea(args, root=DATA_ROOT):
    """:375-413 with the same keys.  'x' is a dense CUDA tensor (the reference wraps the same dense matrix
    in a sparse COO tensor), 'adj' a CUDA sparse COO tensor carrying the kernel-format CSR."""
    raw = read_dbp15k(args.dataset, root)
    dev = torch.device(getattr(args, "device", "cuda"))
    e = len(set(raw["ent1"]) | set(raw["ent2"]))
    r = len(set(raw["rel1"]) | set(raw["rel2"]))
    train, test = _split_links(raw["ill"])
    KG = raw["kg1"] + raw["kg2"]
    x = raw["x"].to(dev)
    adj = sparse_mx_to_torch_sparse_tensor(get_sparse_tensor(e, KG, device=dev))
    head, tail, head_r, tail_r = rfunc(e, KG)
    feat_r = torch.zeros(r, x.shape[1], device=dev)
    for rel in range(r):                                            # :402-404
        hs = torch.as_tensor(head[rel], device=dev)
        ts = torch.as_tensor(tail[rel], device=dev)
        feat_r[rel] = (x[ts].sum(0) - x[hs].sum(0)) / len(head[rel])
    return {"x": x, "adj": adj, "r": feat_r.to_sparse(), "train": train, "test": test, "test_r": raw["ill_r"],
            "triple": KG, "head": head, "tail": tail, "head_r": head_r, "tail_r": tail_r,
            "idx_x": torch.arange(x.shape[0]), "idx_r": torch.arange(r)}


def load_seperate_data_ea(args, root=DATA_ROOT):
    """:416-455 with the same keys (union-graph adjacency 'adj', per-KG 'adj1'/'adj2', index maps)."""
    raw = read_dbp15k(args.dataset, root)
    dev = torch.device(getattr(args, "device", "cuda"))
    index1 = dict(enumerate(raw["ent1"]))
    index2 = dict(enumerate(raw["ent2"]))
    index1_R = {v: i for i, v in index1.items()}
    index2_R = {v: i for i, v in index2.items()}
    E1, E2 = len(index1), len(index2)
    to_t = sparse_mx_to_torch_sparse_tensor
    M1 = to_t(get_sparse_tensor_for_one_graph(E1, raw["kg1"], index1_R, device=dev))
    M2 = to_t(get_sparse_tensor_for_one_graph(E2, raw["kg2"], index2_R, device=dev))
    M = to_t(get_sparse_tensor(E1 + E2, raw["kg1"] + raw["kg2"], device=dev))
    train, test = _split_links(raw["ill"])
    return {"x": raw["x"].to(dev), "adj": M, "e1": E1, "e2": E2, "adj1": M1, "adj2": M2, "train": train, "test": test,
            "index1": index1, "index2": index2, "index1_R": index1_R, "index2_R": index2_R}
