"""Generate tests/golden/*.npz from the LIVE reference (build container only).

Run:  python tests/golden/make_golden.py
Every array below is produced by the reference's own functions
(HestiaSky/GNN-MTL, imported read-only through oracle/ref_shim.py); nothing from
oracle/ea_oracle.py or the CUDA path is involved, so the fixtures pin both.
Seeds are fixed; torch 2.11 / numpy 2.3 / scipy 1.18 CPU.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_shim import ref  # noqa: E402


def tiny_triples(rng, n_ent, n_tri, n_rel):
    h = rng.integers(0, n_ent - 5, n_tri)          # last 5 entities stay isolated
    t = rng.integers(0, n_ent - 5, n_tri)
    r = rng.integers(0, n_rel, n_tri)
    h[:6] = t[:6]                                   # self-loop triples
    h[6:12], t[6:12], r[6:12] = h[12:18], t[12:18], r[12:18]   # exact duplicates
    h[18:24], t[18:24] = t[24:30], h[24:30]         # reversed duplicates
    return [(int(a), int(b), int(c)) for a, b, c in zip(h, r, t)]


def adjacency_fixture():
    rng = np.random.default_rng(11)
    n_ent = 64
    KG = tiny_triples(rng, n_ent, 260, 7)
    coo = ref.data_utils.get_sparse_tensor(n_ent, KG)
    adj = ref.data_utils.sparse_mx_to_torch_sparse_tensor(coo)
    csr = adj.coalesce().to_sparse_csr()
    np.savez(os.path.join(HERE, "adjacency.npz"),
             n_ent=n_ent, triples=np.array(KG, dtype=np.int64),
             coo_row=coo.row.astype(np.int64), coo_col=coo.col.astype(np.int64), coo_val=coo.data,
             crow=csr.crow_indices().numpy(), col=csr.col_indices().numpy(), val=csr.values().numpy())
    return n_ent, KG, adj


def layer_fixture(n_ent, adj):
    import torch.nn.functional as F
    torch.manual_seed(5)
    d = 20
    x = torch.randn(n_ent, d)
    gc = ref.layers.GraphConvolution(d, d, 0.0, F.relu, True)
    hw = ref.layers.HighWayGraphConvolution(d, d, 0.0, F.relu, True, -1, "cpu")
    hw_id = ref.layers.HighWayGraphConvolution(d, d, 0.0, lambda z: z, True, -1, "cpu")
    hw.bias_gate = torch.randn(d) * 0.1
    out = {}
    for name, layer in (("gc", gc), ("hw", hw), ("hwid", hw_id)):
        xin = x.clone().requires_grad_(True)
        y, _ = layer((xin, adj))
        w = torch.randn_like(y)
        (y * w).sum().backward()
        out[name + "_W"] = layer.linear.weight.detach().numpy()
        out[name + "_b"] = layer.linear.bias.detach().numpy()
        if name != "gc":
            out[name + "_G"] = layer.kernel_gate.numpy()
            out[name + "_c"] = layer.bias_gate.numpy()
        out[name + "_y"] = y.detach().numpy()
        out[name + "_seed"] = w.numpy()
        out[name + "_dx"] = xin.grad.numpy()
        out[name + "_dW"] = layer.linear.weight.grad.numpy()
        out[name + "_db"] = layer.linear.bias.grad.numpy()
    np.savez(os.path.join(HERE, "layers.npz"), x=x.numpy(), **out)


def sinkhorn_fixture():
    torch.manual_seed(3)
    X = torch.randn(40, 8) * 0.3
    Y = torch.randn(50, 8) * 0.3
    M = torch.cdist(X, Y, p=2)
    a = torch.rand(40) + 0.5
    b = torch.rand(50) + 0.5
    b = b * a.sum() / b.sum()
    out = {"X": X.numpy(), "Y": Y.numpy(), "M": M.numpy(), "a": a.numpy(), "b": b.numpy()}
    for tag, reg, iters in (("r05_i37", 0.05, 37), ("r01_i200", 0.01, 200), ("r1_conv", 0.5, 1000)):
        P, loss = ref.ot_loss.sinkhorn(a, b, M, reg, numItermax=iters)
        out["P_" + tag] = P.numpy()
        out["loss_" + tag] = loss.numpy()
    # the as-shipped training loss on the same points (plan ignored, column 0 picked)
    ones_a, ones_b = torch.ones(40), torch.ones(50)
    T, _ = ref.ot_loss.sinkhorn(ones_a, ones_b, M, reg=0.01)
    newT = torch.zeros_like(T)
    newT[torch.arange(len(newT)), torch.argmax(newT, dim=1)] = 1
    out["loss_shipped"] = (newT * M.double()).sum().numpy()

    # stabilised variant on the reference test's recipe (cosine cost, uniform vectors, fp64)
    rng = np.random.default_rng(123)
    Va = torch.from_numpy(rng.uniform(size=(30, 10)))
    Vb = torch.from_numpy(rng.uniform(size=(36, 10)))
    C = ref.cderivation.cos_dist_mat(Va, Vb).double().view(1, 30, 36)
    mu = torch.full((1, 30, 1), 1 / 30, dtype=torch.float64)
    nu = torch.full((1, 1, 36), 1 / 36, dtype=torch.float64)
    out.update({"Va": Va.numpy(), "Vb": Vb.numpy(), "C_cos": C.numpy()})
    for tag, eps, iters in (("e2", 1e-2, 100), ("e3_i25", 1e-3, 25)):
        w, k1, k2, K = ref.sinkhorn_loss.sinkhorn_iteration(C, mu, nu, eps, numIterMax=iters)
        out["S2_w_" + tag] = w.numpy()
        out["S2_kl1_" + tag] = k1.numpy()
        out["S2_kl2_" + tag] = k2.numpy()
        out["S2_K_" + tag] = K.numpy()
    out["C_sq"] = ref.cderivation.p_norm_dist_mat(Va, Vb).numpy()
    np.savez(os.path.join(HERE, "sinkhorn.npz"), **out)


def eval_fixture():
    rng = np.random.default_rng(21)
    n, d, npair = 90, 12, 32
    vec = rng.standard_normal((n, d)).astype(np.float32)
    left = rng.permutation(45)[:npair]
    right = rng.permutation(45)[:npair] + 45
    vec[right] = vec[left] + 0.6 * rng.standard_normal((npair, d)).astype(np.float32)
    pairs = np.stack([left, right], 1).astype(np.int64)
    tv = torch.from_numpy(vec)
    hits = ref.eval_utils.get_hits(tv, pairs, top_k=(1, 5, 10))
    at1 = ref.eval_utils.eval_at_1(tv, {"test": pairs})

    class _Args:
        n_nodes, device = n, "cpu"
    base = ref.models_ea.BaseModel(_Args())
    neg = base.get_neg(pairs[:, 0], tv, 7)

    class _U(ref.models_ea.UEAModel):
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.ILL = None
    um = _U()
    data = {"e1": 45, "e2": 45, "index1": {i: i for i in range(45)}, "index2": {i: i + 45 for i in range(45)}}
    um.generate_pairs(tv, data, 20)
    T = torch.from_numpy(rng.random((45, 45)))
    gw = ref.eval_utils.eval_gw_matching_matrix(-T, pairs, {i: i for i in range(45)},
                                                {i + 45: i for i in range(45)}, top_k=(1, 5))
    np.savez(os.path.join(HERE, "eval.npz"), vec=vec, pairs=pairs,
             hits_keys=np.array(list(hits.keys())), hits_vals=np.array(list(hits.values())),
             at1=at1.numpy(), neg=neg.astype(np.int64), mnn=np.asarray(um.ILL, dtype=np.int64),
             T=T.numpy(), gw_keys=np.array(list(gw.keys())), gw_vals=np.array(list(gw.values())))


def margin_fixture():
    """EAModel.get_loss of the live reference (models/models_ea.py:103-123) on integer index arrays
    (the reference builds them as float arrays of integral values, which current torch refuses as indices)."""
    rng = np.random.default_rng(33)
    n, d, t, k = 120, 20, 25, 6
    out = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32) * 0.4).requires_grad_(True)
    ILL = np.stack([rng.permutation(60)[:t], rng.permutation(60)[:t] + 60], 1).astype(np.int64)

    class _Fake:
        pass
    me = _Fake()
    me.neg_num = k
    me.neg_left = np.repeat(ILL[:, 0], k)
    me.neg2_right = np.repeat(ILL[:, 1], k)
    me.neg_right = rng.integers(0, n, t * k)
    me.neg2_left = rng.integers(0, n, t * k)
    loss = ref.models_ea.EAModel.get_loss(me, out, {"train": ILL}, "train")
    loss.backward()
    np.savez(os.path.join(HERE, "margin.npz"), out=out.detach().numpy(), ILL=ILL, k=k, neg_left=me.neg_left,
             neg_right=me.neg_right, neg2_left=me.neg2_left, neg2_right=me.neg2_right,
             loss=loss.detach().numpy(), grad=out.grad.numpy())


def gw_fixture():
    """gw_iterative_1 of the live reference (SinkhornOT/iterative_projection.py:119-120), fp64, 6 projections."""
    import contextlib
    import io
    rng = np.random.default_rng(8)
    Va = torch.from_numpy(rng.uniform(size=(18, 6)))
    Vb = torch.from_numpy(rng.uniform(size=(22, 6)))
    C1 = ref.cderivation.cos_dist_mat(Va, Va).double()
    C2 = ref.cderivation.cos_dist_mat(Vb, Vb).double()
    mu = torch.full((18,), 1 / 18, dtype=torch.float64)
    nu = torch.full((22,), 1 / 22, dtype=torch.float64)
    with contextlib.redirect_stdout(io.StringIO()):
        T, gw = ref.iterative_projection.gw_iterative_1(C1, C2, mu, nu, epsilon=0.02, max_iter=6)
    np.savez(os.path.join(HERE, "gw.npz"), C1=C1.numpy(), C2=C2.numpy(), T=T.numpy(), gw=gw.numpy())


def gat_fixture():
    """SpGraphAttentionLayer / GraphAttentionLayer (layers/att_layers.py) on a graph with no isolated entity
    (the reference asserts on the NaN an empty row would give)."""
    import torch.nn.functional as F
    rng = np.random.default_rng(23)
    n_ent = 48
    KG = tiny_triples(rng, n_ent, 150, 5) + [(i, 0, i + 1) for i in range(n_ent - 1)]
    adj = ref.data_utils.sparse_mx_to_torch_sparse_tensor(ref.data_utils.get_sparse_tensor(n_ent, KG))
    att = load_att()
    torch.manual_seed(9)
    x = torch.randn(n_ent, 24)
    out = {"n_ent": n_ent, "triples": np.array(KG, dtype=np.int64), "x": x.numpy()}
    single = att.SpGraphAttentionLayer(24, 15, 0.0, 0.2, F.elu)
    xin = x.clone().requires_grad_(True)
    y = single(xin, adj)
    seed = torch.randn_like(y)
    (y * seed).sum().backward()
    out.update(s_W=single.W.detach().numpy(), s_a=single.a.detach().numpy(), s_y=y.detach().numpy(),
               s_seed=seed.numpy(), s_dx=xin.grad.numpy(), s_dW=single.W.grad.numpy(), s_da=single.a.grad.numpy())
    multi = att.GraphAttentionLayer(24, 9, 0.0, F.relu, 0.2, 4, True)
    xin = x.clone().requires_grad_(True)
    y, _ = multi((xin, adj))
    seed = torch.randn_like(y)
    (y * seed).sum().backward()
    out.update(m_y=y.detach().numpy(), m_seed=seed.numpy(), m_dx=xin.grad.numpy())
    for i, head in enumerate(multi.attentions):
        out["m_W%d" % i] = head.W.detach().numpy()
        out["m_a%d" % i] = head.a.detach().numpy()
        out["m_dW%d" % i] = head.W.grad.numpy()
        out["m_da%d" % i] = head.a.grad.numpy()
    np.savez(os.path.join(HERE, "gat.npz"), **out)


def load_att():
    from oracle.ref_shim import load
    return load("layers.att_layers")


if __name__ == "__main__":
    gat_fixture()
    gw_fixture()
    margin_fixture()
    n_ent, KG, adj = adjacency_fixture()
    layer_fixture(n_ent, adj)
    sinkhorn_fixture()
    eval_fixture()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
