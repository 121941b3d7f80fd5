"""CPU restatement of the GNN-MTL entity-alignment hot path (test oracle).

TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.  Every function cites the
reference ``file:line`` (relative to the HestiaSky/GNN-MTL checkout) whose
arithmetic it restates.  The third-party arithmetic the reference leans on
(``torch.spmm``, ``torch.cdist``, ``scipy.spatial.distance.cdist``,
``numpy.argsort``; all unpinned upstream — SURVEY.md §8c) is called through the
same libraries here, so timing this module on host cores is a fair "port" CPU
baseline for ``bench.py``.

Pin status: checked against the live reference in ``tests/test_oracle_vs_reference.py``
(build container only) and against ``tests/golden/*.npz`` everywhere.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.spatial.distance as _ssd
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- #
# A1–A3  adjacency: triples -> normalised sparse matrix                        #
# --------------------------------------------------------------------------- #


def adjacency_insertion_order(triples):
    """Pure-Python walk, small inputs only.  utils/data_utils.py:296-336.

    Returns (rows, cols, vals_fp64) in the order the reference's dict would be
    iterated: symmetric edges in first-sight order, then one self-loop per node
    in degree-dict order (:319-320).  Degree rule (:298-305): a node starts at 1
    the first time it is seen at either end of any triple; every triple whose
    ends differ adds 1 to both ends (duplicates and both directions count).
    """
    deg = {}
    for h, _r, t in triples:
        deg.setdefault(h, 1)
        deg.setdefault(t, 1)
        if h != t:
            deg[h] += 1
            deg[t] += 1
    seen = {}
    for h, _r, t in triples:
        if h == t:
            continue
        seen.setdefault((h, t), 1)
        seen.setdefault((t, h), 1)
    for node in deg:
        seen[(node, node)] = 1
    rows, cols, vals = [], [], []
    for (i, j), w in seen.items():
        rows.append(i)
        cols.append(j)
        # left-to-right float64: (w / sqrt(di)) / sqrt(dj)   data_utils.py:334
        vals.append(w / math.sqrt(deg[i]) / math.sqrt(deg[j]))
    return (np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64),
            np.asarray(vals, dtype=np.float64))


def adjacency_degrees(n_ent, heads, tails):
    """Vectorised degree rule of utils/data_utils.py:298-305 (0 = never seen)."""
    heads = np.asarray(heads, dtype=np.int64)
    tails = np.asarray(tails, dtype=np.int64)
    deg = np.zeros(n_ent, dtype=np.int64)
    touched = np.zeros(n_ent, dtype=bool)
    touched[heads] = True
    touched[tails] = True
    deg[touched] = 1
    off = heads != tails
    deg += np.bincount(heads[off], minlength=n_ent)
    deg += np.bincount(tails[off], minlength=n_ent)
    return deg


def adjacency_csr(n_ent, heads, tails):
    """Sorted, duplicate-free CSR of the reference adjacency.

    Contract: equals
    ``sparse_mx_to_torch_sparse_tensor(get_sparse_tensor(e, KG)).coalesce().to_sparse_csr()``
    — utils/data_utils.py:296-336 then :51-57 (fp64 -> fp32 value cast at :55).
    Returns (crow int64 [n+1], col int64 [nnz], val float32 [nnz]).
    """
    heads = np.asarray(heads, dtype=np.int64)
    tails = np.asarray(tails, dtype=np.int64)
    deg = adjacency_degrees(n_ent, heads, tails)
    off = heads != tails
    h, t = heads[off], tails[off]
    nodes = np.nonzero(deg)[0]
    key = np.concatenate([h * n_ent + t, t * n_ent + h, nodes * n_ent + nodes])
    key = np.unique(key)
    row = key // n_ent
    col = key % n_ent
    val64 = (1.0 / np.sqrt(deg[row].astype(np.float64))) / np.sqrt(deg[col].astype(np.float64))
    crow = np.zeros(n_ent + 1, dtype=np.int64)
    np.cumsum(np.bincount(row, minlength=n_ent), out=crow[1:])
    return crow, col.astype(np.int64), val64.astype(np.float32)


def adjacency_torch_coo(n_ent, heads, tails):
    """The tensor the reference hands to its layers (utils/data_utils.py:51-57),
    already coalesced.  fp32 values, int64 indices."""
    crow, col, val = adjacency_csr(n_ent, heads, tails)
    row = np.repeat(np.arange(n_ent, dtype=np.int64), np.diff(crow))
    idx = torch.from_numpy(np.stack([row, col]))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(val), (n_ent, n_ent)).coalesce()


# --------------------------------------------------------------------------- #
# L1–L3  GCN / highway-GCN layer                                              #
# --------------------------------------------------------------------------- #


def _act(name_or_fn):
    if name_or_fn is None or name_or_fn == "identity":
        return lambda z: z
    if isinstance(name_or_fn, str):
        return getattr(F, name_or_fn)
    return name_or_fn


def gcn_layer(x, adj, weight, bias, act="relu"):
    """layers/layers.py:30-39 (dropout p=0): act(adj @ (x Wᵀ + b))."""
    hidden = F.linear(x, weight, bias)
    agg = torch.sparse.mm(adj, hidden) if adj.is_sparse else adj @ hidden
    return _act(act)(agg)


def highway_layer(x, adj, weight, bias, gate_w, gate_b, act="relu"):
    """layers/layers.py:59-77 (dropout p=0).

    t = sigmoid(x G + c);  out = t * act(adj @ (x Wᵀ + b)) + (1 - t) * x.
    """
    agg = gcn_layer(x, adj, weight, bias, act)
    t = torch.sigmoid(x @ gate_w + gate_b)
    return t * agg + (1.0 - t) * x


def hgcn_stack(x, adj, params, acts):
    """models/encoders.py:53-66 + models/decoders.py:40-47: highway layers chained.
    ``params`` = list of (W, b, G, c); ``acts`` = list of activation names."""
    h = x
    for (w, b, g, c), a in zip(params, acts):
        h = highway_layer(h, adj, w, b, g, c, a)
    return h


def gat_layer(x, adj, W, a, alpha=0.2, act="identity"):
    """layers/att_layers.py:29-61 (dropout p=0): sparse single-head graph attention.

    h = x W;  e_ij = exp(-leakyrelu_alpha(a · [h_i ‖ h_j])) on the stored (i, j) of adj (its values are not used);
    out = act((sum_j e_ij h_j) / sum_j e_ij).  Differentiable torch ops in the input dtype.
    """
    edge = adj.coalesce().indices()
    h = x @ W
    edge_h = torch.cat((h[edge[0]], h[edge[1]]), dim=1)                 # [E, 2D]
    edge_e = torch.exp(-F.leaky_relu(edge_h @ a.reshape(-1), alpha))    # [E]
    n = x.shape[0]
    rowsum = torch.zeros(n, dtype=h.dtype).index_add(0, edge[0], edge_e)
    agg = torch.zeros(n, h.shape[1], dtype=h.dtype).index_add(0, edge[0], edge_e[:, None] * h[edge[1]])
    return _act(act)(agg / rowsum[:, None])


def gat_multihead(x, adj, heads, alpha=0.2, act="identity", concat=True):
    """layers/att_layers.py:81-93: ``heads`` = list of (W, a); concatenate or average the head outputs."""
    outs = [gat_layer(x, adj, W, a, alpha, act) for W, a in heads]
    return torch.cat(outs, dim=1) if concat else torch.stack(outs, dim=2).mean(dim=2)


# --------------------------------------------------------------------------- #
# S4  cost matrices                                                           #
# --------------------------------------------------------------------------- #


def cost_l2(x, y):
    """models/models_ea.py:218 — torch.cdist(X, Y, p=2)."""
    return torch.cdist(x, y, p=2)


def cost_sqeuclid(x, y):
    """SinkhornOT/cderivation.py:14-26 with p=2: sum_k (x_k - y_k)^2, no root."""
    return ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)


def cost_cosine(x, y):
    """SinkhornOT/cderivation.py:44-61: 1 - cosine_similarity (torch eps 1e-8)."""
    return 1.0 - F.cosine_similarity(x[:, None, :], y[None, :, :], dim=-1)


# --------------------------------------------------------------------------- #
# S1  scaling-form Sinkhorn in float64                                        #
# --------------------------------------------------------------------------- #


def sinkhorn_scaling(a, b, M, reg, numItermax=1000, stopThr=1e-9, return_info=False):
    """utils/ot_loss.py:26-76.

    All arithmetic in float64 (:27).  u0 = 1/I, v0 = 1/J (:38-39); Gibbs kernel
    K = exp(M / -reg) (:41-43); per sweep v = b / (Kᵀ u) then u = a / (K v)
    (:53-55, the reference writes the second as 1 / ((K / a) v)); numerical
    failure rolls back one sweep and stops (:57-62); on sweeps 0, 10, 20, … the
    column marginal v ∘ (Kᵀ u) is compared with b in the 2-norm (:64-66) and the
    loop ends when that is ≤ stopThr or after numItermax sweeps (:50).
    Returns plan P = diag(u) K diag(v) and <P, M> (:75-76).
    """
    a = a.detach().to(torch.float64).reshape(-1)
    b = b.detach().to(torch.float64).reshape(-1)
    M = M.detach().to(torch.float64)
    n_src, n_tgt = M.shape
    assert a.numel() == n_src and b.numel() == n_tgt
    u = torch.full((n_src, 1), 1.0 / n_src, dtype=torch.float64, device=M.device)
    v = torch.full((n_tgt, 1), 1.0 / n_tgt, dtype=torch.float64, device=M.device)
    K = torch.exp(M / (-reg))
    K_over_a = K / a.reshape(-1, 1)
    sweeps, err, failed = 0, 1.0, False
    while err > stopThr and sweeps < numItermax:
        u_old, v_old = u, v
        col = K.t() @ u
        v = b.reshape(-1, 1) / col
        u = 1.0 / (K_over_a @ v)
        bad = bool((col == 0).any()) or not bool(torch.isfinite(u).all()) \
            or not bool(torch.isfinite(v).all())
        if bad:
            u, v, failed = u_old, v_old, True
            break
        if sweeps % 10 == 0:
            marg = (v * (K.t() @ u)).reshape(-1)
            err = float(torch.linalg.vector_norm(marg - b))
        sweeps += 1
    P = u * K * v.reshape(1, -1)
    loss = (P * M).sum()
    if return_info:
        return P, loss, {"sweeps": sweeps, "err": err, "failed": failed,
                         "log_u": torch.log(u).reshape(-1), "log_v": torch.log(v).reshape(-1)}
    return P, loss


# --------------------------------------------------------------------------- #
# S2  stabilised (absorbing) Sinkhorn                                         #
# --------------------------------------------------------------------------- #

_CLAMP_HI = 1e30
_ABSORB_AT = 1e20
_KL_EPS = 1e-7


def _kl_terms(x, y):
    """SinkhornOT/sinkhorn_loss.py:20-30."""
    ratio = x / (y + _KL_EPS)
    return y * (ratio * torch.log(ratio + _KL_EPS) - ratio + 1)


def sinkhorn_stabilised(C, mu, nu, epsilon, numIterMax=100, tol=1e-9, return_info=False):
    """SinkhornOT/sinkhorn_loss.py:159-220, batched [B, I, J].

    Potentials (u, v) start at 0 and the scaling vector b at 1 (:190-193).  Each
    sweep: a = mu / (K b) over the last axis, b = nu / (Kᵀ a) over the middle
    axis, both clamped to [0, 1e30] (:197-201).  On sweep 0, 10, 20, …, or when a
    or b exceeds 1e20, or on the final sweep, the scalings are absorbed into the
    potentials, K is rebuilt as clamp(exp((u + v - C)/eps)), b resets to 1, and
    the primal cost <K, C> is compared with its previous value; relative change
    below ``tol`` ends the loop (:203-213).  Returns (cost, KL(row marg ‖ mu),
    KL(col marg ‖ nu), K) (:215-220).
    """
    C = C.detach()
    dt = C.dtype

    def gibbs(uu, vv):
        return torch.clamp(torch.exp((uu + vv - C) / epsilon), 0, _CLAMP_HI)

    u = torch.zeros_like(mu, dtype=dt)
    v = torch.zeros_like(nu, dtype=dt)
    scale_b = torch.ones_like(nu, dtype=dt)
    K = gibbs(u, v)
    cost_prev = (K * C).sum((-1, -2)).squeeze()
    cost_now = cost_prev
    sweeps_done = 0
    for it in range(numIterMax):
        sweeps_done = it + 1
        scale_a = torch.clamp(mu / (K * scale_b).sum(-1, keepdim=True), 0, _CLAMP_HI)
        scale_b = torch.clamp(nu / (K * scale_a).sum(-2, keepdim=True), 0, _CLAMP_HI)
        absorb = (it % 10 == 0) or bool(scale_a.max() > _ABSORB_AT) \
            or bool(scale_b.max() > _ABSORB_AT) or (it == numIterMax - 1)
        if absorb:
            u = u + epsilon * torch.log(scale_a)
            v = v + epsilon * torch.log(scale_b)
            K = gibbs(u, v)
            scale_b = torch.ones_like(nu, dtype=dt)
            cost_now = (K * C).sum((-1, -2)).squeeze()
            if abs(cost_now - cost_prev) / abs(cost_prev) < tol:
                break
            cost_prev = cost_now
    kl_row = _kl_terms(K.sum(-1, keepdim=True), mu).sum(-2).squeeze()
    kl_col = _kl_terms(K.sum(-2, keepdim=True), nu).sum(-1).squeeze()
    if return_info:
        return cost_now, kl_row, kl_col, K, {"sweeps": sweeps_done, "u": u, "v": v}
    return cost_now, kl_row, kl_col, K


# --------------------------------------------------------------------------- #
# S3  the (quirky) Wasserstein training loss                                  #
# --------------------------------------------------------------------------- #


def wasserstein_loss_as_shipped(X, Y, reg=0.01, numItermax=1000, stopThr=1e-9):
    """models/models_ea.py:214-224 for already-sampled X, Y.

    The Sinkhorn plan is computed and then ignored: the one-hot "newT" is built
    from argmax of an all-zero tensor, i.e. column 0 of every row, so the value
    is sum_i M[i, 0] with M = cdist(X, Y, p=2) carrying the gradient.
    """
    n = X.shape[0]
    M = torch.cdist(X, Y, p=2)
    sinkhorn_scaling(torch.ones(n, device=X.device), torch.ones(Y.shape[0], device=X.device), M.detach(), reg,
                     numItermax=numItermax, stopThr=stopThr)
    pick = torch.zeros_like(M)
    pick[torch.arange(n, device=X.device), torch.zeros(n, dtype=torch.long, device=X.device)] = 1
    return (pick * M).sum()


# --------------------------------------------------------------------------- #
# E1–E5  alignment evaluation, negatives, mutual nearest neighbours           #
# --------------------------------------------------------------------------- #


def l1_matrix(L, R):
    """utils/eval_utils.py:74 / models/models_ea.py:24,149 — scipy cityblock in
    float64 (fp32 inputs are widened by scipy before the loop)."""
    return _ssd.cdist(np.asarray(L), np.asarray(R), metric="cityblock")


def diagonal_ranks(sim):
    """Rank of the true match in every row and every column of ``sim``.

    utils/eval_utils.py:77-89 takes ``argsort`` of the row (column) and looks up
    where index i landed.  NumPy's default sort is unstable, so ties are
    unspecified upstream; the contract here is the stable order:
    rank = #(strictly smaller) + #(equal with a lower index).
    """
    n = sim.shape[0]
    diag = sim[np.arange(n), np.arange(n)]
    idx = np.arange(sim.shape[1])
    row_rank = (sim < diag[:, None]).sum(1) + ((sim == diag[:, None]) & (idx[None, :] < np.arange(n)[:, None])).sum(1)
    idr = np.arange(sim.shape[0])
    col_rank = (sim < diag[None, :]).sum(0) + ((sim == diag[None, :]) & (idr[:, None] < np.arange(n)[None, :])).sum(0)
    return row_rank.astype(np.int64), col_rank.astype(np.int64)


def diagonal_ranks_by_sort(sim):
    """Same as :func:`diagonal_ranks`, literally via stable argsort (small n)."""
    n = sim.shape[0]
    rr = np.array([int(np.where(np.argsort(sim[i, :], kind="stable") == i)[0][0]) for i in range(n)])
    cr = np.array([int(np.where(np.argsort(sim[:, i], kind="stable") == i)[0][0]) for i in range(n)])
    return rr, cr


def hits_from_ranks(row_rank, col_rank, top_k):
    """utils/eval_utils.py:91-98 — percentages, keys Hits@k_l then Hits@k_r."""
    n = len(row_rank)
    out = {}
    for k in top_k:
        out["Hits@{}_l".format(k)] = int((row_rank < k).sum()) / n * 100
    for k in top_k:
        out["Hits@{}_r".format(k)] = int((col_rank < k).sum()) / n * 100
    return out


def get_hits(vec, test_pair, top_k=(1, 10, 50, 100)):
    """utils/eval_utils.py:71-98."""
    vec = np.asarray(vec.detach().cpu() if torch.is_tensor(vec) else vec)
    pairs = np.asarray(test_pair)
    sim = l1_matrix(vec[pairs[:, 0]], vec[pairs[:, 1]])
    rr, cr = diagonal_ranks(sim)
    return hits_from_ranks(rr, cr, top_k)


def eval_matching_matrix(T, test_pair, index1_R, index2_R, top_k=(1, 10, 50, 100)):
    """utils/eval_utils.py:133-159: same ranking on a given score sub-block."""
    rows = [index1_R[l] for l, _ in test_pair]
    cols = [index2_R[r] for _, r in test_pair]
    sim = np.asarray(T)[rows, :][:, cols]
    rr, cr = diagonal_ranks(sim)
    return hits_from_ranks(rr, cr, top_k)


def eval_at_1(outputs, test_pair):
    """utils/eval_utils.py:161-168: fp32 torch.cdist(p=1), row argmin == i, in %."""
    pairs = np.asarray(test_pair)
    D = torch.cdist(outputs[pairs[:, 0]], outputs[pairs[:, 1]], p=1)
    hit = (torch.argmin(D, dim=1) == torch.arange(len(pairs))).float()
    return hit.sum() / len(hit) * 100


def nearest_negatives(anchor_ids, output, k):
    """models/models_ea.py:19-30: for each anchor the k entities ranked 1..k by
    fp64 L1 distance over *all* entities (rank 0, normally the anchor itself, is
    dropped).  Stable tie order.  Flat int64 [t*k]."""
    out = np.asarray(output.detach().cpu() if torch.is_tensor(output) else output)
    sim = l1_matrix(out[np.asarray(anchor_ids)], out)
    order = np.argsort(sim, axis=1, kind="stable")
    return order[:, 1:k + 1].reshape(-1).astype(np.int64)


def mutual_nearest_pairs(outputs, L_ids, R_ids, bsz):
    """models/models_ea.py:143-167: fp64 L1 between the two entity sets, row and
    column argmin, keep i where colargmin[rowargmin[i]] == i, order by distance
    (np.argsort default) and keep the first ``bsz``; positions are *local*
    (the reference stores i and v, not L[i] and R[v] — :158)."""
    out = np.asarray(outputs.detach().cpu() if torch.is_tensor(outputs) else outputs)
    M = l1_matrix(out[np.asarray(L_ids)], out[np.asarray(R_ids)])
    r_arg, r_min = M.argmin(1), M.min(1)
    c_arg = M.argmin(0)
    keep = np.nonzero(c_arg[r_arg] == np.arange(M.shape[0]))[0]
    pairs = np.stack([keep, r_arg[keep]], 1)
    order = np.argsort(r_min[keep], kind="stable")[:bsz]
    return pairs[order]


# --------------------------------------------------------------------------- #
# §8f-1  margin ranking loss with hard negatives                               #
# --------------------------------------------------------------------------- #


def margin_loss(outputs, ILL, neg_left, neg_right, neg2_left, neg2_right, k, gamma=1.0):
    """models/models_ea.py:103-123 (EAModel.get_loss) == :185-204 (UEAModel.get_loss):
    A = |x_l - x_r|_1 per pair; hinge relu(A + gamma - |x_nl - x_nr|_1) over both negative sets;
    mean over 2 t k."""
    ILL = np.asarray(ILL)
    t = len(ILL)

    def ix(a):
        return torch.as_tensor(np.asarray(a, dtype=np.int64))
    A = (outputs[ix(ILL[:, 0])] - outputs[ix(ILL[:, 1])]).abs().sum(1)
    D = (A + gamma).reshape(t, 1)
    B1 = (outputs[ix(neg_left)] - outputs[ix(neg_right)]).abs().sum(1).reshape(t, k)
    B2 = (outputs[ix(neg2_left)] - outputs[ix(neg2_right)]).abs().sum(1).reshape(t, k)
    return (F.relu(D - B1).sum() + F.relu(D - B2).sum()) / (2.0 * t * k)


# --------------------------------------------------------------------------- #
# §8f-3  Gromov-Wasserstein by iterative projection                            #
# --------------------------------------------------------------------------- #


def gw_iterative(C1, C2, mu, nu, epsilon, max_iter, tol=1e-9):
    """SinkhornOT/iterative_projection.py:8-60 with g=False (gw_iterative_1, :119-120) and
    cderivation.py:147-163,180-183: constC = ½(C1² mu 1ᵀ + 1 nuᵀ C2ᵀ²); L(T) = constC - C1 T C2ᵀ;
    T <- sinkhorn_stabilised(2 L(T), mu, nu, eps); stop when ||T_old - T||_F < tol.  Returns (T [1,I,J], gw)."""
    I, J = C1.shape[0], C2.shape[0]
    mu3, nu3 = mu.reshape(1, I, 1), nu.reshape(1, 1, J)
    constC = 0.5 * (C1 ** 2) @ mu.reshape(I, 1) + 0.5 * nu.reshape(1, J) @ (C2.t() ** 2)
    T_old = torch.full((I, J), 1.0 / (I * J), dtype=C1.dtype)
    lt = constC - C1 @ (T_old @ C2.t())
    gw = (T_old * lt).sum()
    T = T_old
    for _ in range(max_iter):
        gw, _, _, T = sinkhorn_stabilised(2 * lt.reshape(1, I, J), mu3, nu3, epsilon)
        if torch.linalg.norm((T_old - T).reshape(-1)) < tol:
            break
        T_old = T
        lt = constC - C1 @ (T_old.reshape(I, J) @ C2.t())
    return T, gw
