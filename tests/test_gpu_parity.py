"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the
golden fixtures written by the live reference.  Bars (BASELINE.json north_star):
CSR arrays / ranks / top-k indices bit-exact; embeddings, gradients, OT loss and
plan within 1e-4 relative (max-norm); Hits@k identical."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

REL = 1e-4   # north_star tolerance for fp32 quantities


def relerr(got, want):
    got = torch.as_tensor(got).double().cpu()
    want = torch.as_tensor(want).double().cpu()
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


# ------------------------------------------------------------------ adjacency --

def test_adjacency_golden_bit_exact(golden_dir, dev):
    from gnn_mtl_b200.adjacency import DeviceAdjacency
    g = _load(golden_dir, "adjacency.npz")
    adj = DeviceAdjacency.from_triples(int(g["n_ent"]), g["triples"], device=dev)
    assert np.array_equal(adj.crow.cpu().numpy(), g["crow"])
    assert np.array_equal(adj.col64.cpu().numpy(), g["col"])
    assert np.array_equal(adj.val.cpu().numpy().view(np.uint32), g["val"].view(np.uint32))
    assert np.array_equal(adj.csr.rowptr.cpu().numpy(), g["crow"].astype(np.int32))


@pytest.mark.parametrize("shape", ["tiny", "dbp15k"])
def test_adjacency_vs_oracle(shape, dev):
    from oracle import ea_oracle as orc
    from gnn_mtl_b200.adjacency import DeviceAdjacency
    from gnn_mtl_b200.synth import make_kg_pair
    kg = make_kg_pair(shape, features=False)
    tri = kg["triples"]
    crow, col, val = orc.adjacency_csr(kg["n"], tri[:, 0], tri[:, 2])
    adj = DeviceAdjacency.from_triples(kg["n"], tri, device=dev)
    assert adj.nnz == len(col)
    assert np.array_equal(adj.crow.cpu().numpy(), crow)
    assert np.array_equal(adj.col64.cpu().numpy(), col)
    assert np.array_equal(adj.val.cpu().numpy().view(np.uint32), val.view(np.uint32))
    # transposed CSR == scipy transpose (values bitwise, order by source row)
    import scipy.sparse as sp
    A = sp.csr_matrix((val, col, crow), shape=(kg["n"], kg["n"]))
    At = A.T.tocsr()
    At.sort_indices()
    t = adj.csr_t
    assert np.array_equal(t.rowptr.cpu().numpy(), At.indptr.astype(np.int32))
    assert np.array_equal(t.col.cpu().numpy(), At.indices.astype(np.int32))
    assert np.array_equal(t.val.cpu().numpy().view(np.uint32), At.data.astype(np.float32).view(np.uint32))


def test_adjacency_edge_cases(dev):
    from oracle import ea_oracle as orc
    from gnn_mtl_b200.adjacency import DeviceAdjacency
    # no triples at all: all-zero matrix
    adj = DeviceAdjacency.from_triples(7, np.zeros((0, 3), dtype=np.int64), device=dev)
    assert adj.nnz == 0 and adj.crow.cpu().tolist() == [0] * 8
    # only self-loop triples, duplicates, isolated tail entity
    tri = np.array([[2, 0, 2], [2, 1, 2], [0, 0, 1], [1, 0, 0], [0, 0, 1]], dtype=np.int64)
    crow, col, val = orc.adjacency_csr(5, tri[:, 0], tri[:, 2])
    adj = DeviceAdjacency.from_triples(5, tri, device=dev)
    assert np.array_equal(adj.crow.cpu().numpy(), crow) and np.array_equal(adj.col64.cpu().numpy(), col)
    assert np.array_equal(adj.val.cpu().numpy().view(np.uint32), val.view(np.uint32))
    with pytest.raises(IndexError):
        DeviceAdjacency.from_triples(3, np.array([[0, 0, 5]]), device=dev)


# --------------------------------------------------------------------- layers --

def test_layers_golden(golden_dir, dev):
    from gnn_mtl_b200.adjacency import DeviceAdjacency
    from gnn_mtl_b200.layers.layers import GraphConvolution, HighWayGraphConvolution
    ga = _load(golden_dir, "adjacency.npz")
    adj = DeviceAdjacency.from_triples(int(ga["n_ent"]), ga["triples"], device=dev).to_torch_coo()
    g = _load(golden_dir, "layers.npz")
    d = g["x"].shape[1]
    for name, act in (("gc", F.relu), ("hw", F.relu), ("hwid", lambda z: z)):
        if name == "gc":
            layer = GraphConvolution(d, d, 0.0, act, True)
        else:
            layer = HighWayGraphConvolution(d, d, 0.0, act, True, 0, dev)
            layer.kernel_gate = torch.from_numpy(g[name + "_G"]).to(dev)
            layer.bias_gate = torch.from_numpy(g[name + "_c"]).to(dev)
        layer = layer.to(dev)
        with torch.no_grad():
            layer.linear.weight.copy_(torch.from_numpy(g[name + "_W"]))
            layer.linear.bias.copy_(torch.from_numpy(g[name + "_b"]))
        x = torch.from_numpy(g["x"]).to(dev).requires_grad_(True)
        y, adj_out = layer((x, adj))
        assert adj_out is adj
        (y * torch.from_numpy(g[name + "_seed"]).to(dev)).sum().backward()
        assert relerr(y, g[name + "_y"]) < REL
        assert relerr(x.grad, g[name + "_dx"]) < REL
        assert relerr(layer.linear.weight.grad, g[name + "_dW"]) < REL
        assert relerr(layer.linear.bias.grad, g[name + "_db"]) < REL


def _oracle_stack(kg, x, params, acts):
    from oracle import ea_oracle as orc
    tri = kg["triples"]
    adj = orc.adjacency_torch_coo(kg["n"], tri[:, 0], tri[:, 2])
    xs = x.clone().requires_grad_(True)
    ps = [tuple(p.clone().requires_grad_(i < 2) for i, p in enumerate(q)) for q in params]
    y = orc.hgcn_stack(xs, adj, ps, acts)
    return xs, ps, y


def _oracle_stack_pinned(kg, x, params, acts, gpu_act, tau=1e-5):
    """The oracle's H-GCN stack (fp32, CPU — what the reference computes) with the ReLU branch PINNED to the GPU's
    decision at the pre-activations the reference itself cannot decide: entries whose magnitude is below ``tau``
    of the layer's mean |S|.  (fp32 CPU vs fp64 already disagree on a handful of those at the benchmark size —
    which of them an implementation takes depends on its BLAS's summation order — and each such flip moves the
    gradient of that row by ~5e-3 of the max-norm.)  Everywhere else the two masks must be IDENTICAL (asserted).
    Returns (xs, ps, y, number of pinned entries, number of ReLU entries)."""
    from oracle import ea_oracle as orc
    tri = kg["triples"]
    adj = orc.adjacency_torch_coo(kg["n"], tri[:, 0], tri[:, 2])
    xs = x.clone().requires_grad_(True)
    ps = [tuple(p.clone().requires_grad_(i < 2) for i, p in enumerate(q)) for q in params]
    h, pinned, total = xs, 0, 0
    for (W, b, G, c), act, ga in zip(ps, acts, gpu_act):
        S = torch.sparse.mm(adj, F.linear(h, W, b))                  # layers/layers.py:61-64 (orc.gcn_layer)
        if act == "relu":
            own = S.detach() > 0
            theirs = ga.cpu() > 0
            differ = own != theirs
            undecidable = S.detach().abs() < tau * S.detach().abs().mean()
            assert not bool((differ & ~undecidable).any()), "ReLU branch differs at a decidable pre-activation"
            pinned += int(differ.sum())
            total += S.numel()
            a = S * theirs.to(S.dtype)
        else:
            a = S
        t = torch.sigmoid(h @ G + c)                                 # :69-76
        h = t * a + (1.0 - t) * h
    return xs, ps, h, pinned, total


@pytest.mark.parametrize("shape,dim", [("tiny", 300), ("tiny", 128), ("tiny", 50), ("dbp15k", 300), ("dbp100k", 300)])
def test_hgcn_stack_vs_oracle(shape, dim, dev):
    """2 encoder + 1 decoder highway layers (models/encoders.py:53-66, decoders.py:40-47); the last case is the
    benchmark's own graph (200k nodes, BASELINE.json config 3): forward and every gradient against the oracle.
    At the two large sizes the ReLU branch of undecidable pre-activations is pinned (see _oracle_stack_pinned); the
    forward comparison never needs it."""
    from gnn_mtl_b200.adjacency import DeviceAdjacency
    from gnn_mtl_b200.layers import layers as L
    from gnn_mtl_b200.synth import make_kg_pair
    torch.manual_seed(1)
    kg = make_kg_pair(shape, dim=dim)
    x = torch.from_numpy(kg["x"])
    acts = [F.relu, F.relu, (lambda z: z)]
    layers = [L.HighWayGraphConvolution(dim, dim, 0.0, a, True, -1, "cpu") for a in acts]
    params = [(l.linear.weight.detach(), l.linear.bias.detach(), l.kernel_gate, l.bias_gate) for l in layers]
    seed = torch.randn(kg["n"], dim)

    adj = DeviceAdjacency.from_triples(kg["n"], kg["triples"], device=dev).to_torch_coo()
    xg = x.to(dev).requires_grad_(True)
    h = xg
    L.CAPTURE_ACT = []
    try:
        for l in layers:
            l.to(dev)
            h, _ = l((h, adj))
        gpu_act = L.CAPTURE_ACT
    finally:
        L.CAPTURE_ACT = None
    (h * seed.to(dev)).sum().backward()

    names = ["relu", "relu", "identity"]
    xs0, ps0, y_plain = _oracle_stack(kg, x, params, names)
    assert relerr(h, y_plain) < REL                                   # forward: the reference as it is
    if shape == "tiny":
        xs, ps, y_ref = xs0, ps0, y_plain
    else:
        xs, ps, y_ref, pinned, total = _oracle_stack_pinned(kg, x, params, names, gpu_act)
        print("%s: %d of %d ReLU branches pinned" % (shape, pinned, total))
        assert pinned <= max(4, total // 1000000)                      # a few per 10^7 (fp32 rounding), not more
        assert relerr(h, y_ref) < REL
    (y_ref * seed).sum().backward()
    assert relerr(xg.grad, xs.grad) < REL
    for l, p in zip(layers, ps):
        assert relerr(l.linear.weight.grad, p[0].grad) < REL
        assert relerr(l.linear.bias.grad, p[1].grad) < REL


def test_spmm_hub_rows_and_general_adj(dev):
    """Power-law graph with rows far above the long-row threshold; and an
    unsymmetric adjacency handed in as a plain torch sparse tensor."""
    from gnn_mtl_b200 import ops
    from gnn_mtl_b200.adjacency import DeviceAdjacency, resolve
    from gnn_mtl_b200.synth import make_powerlaw_graph
    n = 20000
    heads, tails = make_powerlaw_graph(n, 12, seed=3)
    adj = DeviceAdjacency.from_heads_tails(n, torch.from_numpy(heads).to(dev), torch.from_numpy(tails).to(dev))
    assert adj.csr.n_long > 0 and adj.csr.n_seg > adj.csr.n_long
    H = torch.randn(n, 300, device=dev)
    ref = torch.sparse.mm(adj.to_torch_coo().cpu(), H.cpu())
    out, _ = ops.spmm(adj.csr, H)
    assert relerr(out, ref) < REL
    out_t, _ = ops.spmm(adj.csr_t, H)
    assert relerr(out_t, ref) < REL   # reference-built adjacency is symmetric
    # general (unsymmetric, rectangular-valued) sparse input
    idx = torch.randint(0, 500, (2, 4000))
    vals = torch.randn(4000)
    A = torch.sparse_coo_tensor(idx, vals, (500, 500))
    Hs = torch.randn(500, 36)
    da = resolve(A.to(dev))
    o1, _ = ops.spmm(da.csr, Hs.to(dev))
    o2, _ = ops.spmm(da.csr_t, Hs.to(dev))
    assert relerr(o1, torch.sparse.mm(A.coalesce(), Hs)) < REL
    assert relerr(o2, torch.sparse.mm(A.coalesce().t(), Hs)) < REL


# ----------------------------------------------------------------------- eval --

def test_eval_golden(golden_dir, dev):
    from gnn_mtl_b200.utils.eval_utils import get_hits, eval_at_1, eval_gw_matching_matrix
    from gnn_mtl_b200.models.models_ea import BaseModel, UEAModel
    g = _load(golden_dir, "eval.npz")
    vec = torch.from_numpy(g["vec"])
    pairs = g["pairs"]
    hits = get_hits(vec, pairs, top_k=(1, 5, 10))          # CPU tensor in, as the reference's caller does
    assert list(hits.keys()) == [str(k) for k in g["hits_keys"]]
    assert list(hits.values()) == list(g["hits_vals"])
    assert float(eval_at_1(vec.to(dev), {"test": pairs})) == float(g["at1"])

    class _A:
        n_nodes, device = 90, dev
    neg = BaseModel(_A()).get_neg(pairs[:, 0], vec.to(dev), 7)
    assert neg.dtype == np.int64 and np.array_equal(neg, g["neg"])

    class _U(UEAModel):
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.ILL = None
    um = _U()
    data = {"e1": 45, "e2": 45, "index1": {i: i for i in range(45)}, "index2": {i: i + 45 for i in range(45)}}
    um.generate_pairs(vec.to(dev), data, 20)
    assert np.array_equal(um.ILL, g["mnn"])
    gw = eval_gw_matching_matrix(torch.from_numpy(-g["T"]), pairs, {i: i for i in range(45)},
                                 {i + 45: i for i in range(45)}, top_k=(1, 5))
    assert list(gw.values()) == list(g["gw_vals"])


def test_l1_matrix_bit_exact_and_ranks(dev):
    from oracle import ea_oracle as orc
    from gnn_mtl_b200 import ops
    rng = np.random.default_rng(9)
    for n, m, d in ((1, 1, 1), (70, 131, 300), (257, 64, 33), (1500, 1500, 300)):
        L = rng.standard_normal((n, d)).astype(np.float32)
        R = rng.standard_normal((m, d)).astype(np.float32)
        D = ops.l1_matrix(torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)).cpu().numpy()
        want = orc.l1_matrix(L, R)
        assert np.array_equal(D.view(np.uint64), want.view(np.uint64)), (n, m, d)
    # ranks incl. exact ties (duplicated rows) — stable order
    L = rng.standard_normal((900, 40)).astype(np.float32)
    R = L + 0.8 * rng.standard_normal((900, 40)).astype(np.float32)
    R[100:110] = R[90:100]
    L[200:205] = L[300:305]
    sim = orc.l1_matrix(L, R)
    rr, cr = orc.diagonal_ranks(sim)
    for streamed in (True, False):
        gr, gc = ops.l1_ranks(torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev), block_bytes=900 * 8 * 128,
                              streamed=streamed)
        assert np.array_equal(gr.cpu().numpy(), rr) and np.array_equal(gc.cpu().numpy(), cr), streamed


def test_hits_topk_argmin_vs_oracle_medium(dev):
    from oracle import ea_oracle as orc
    from gnn_mtl_b200 import ops
    from gnn_mtl_b200.utils.eval_utils import get_hits
    from gnn_mtl_b200.synth import make_kg_pair
    kg = make_kg_pair("dbp15k")
    vec = torch.from_numpy(kg["x"])
    pairs = kg["test"][:3000]
    assert get_hits(vec.to(dev), pairs, top_k=(1, 10)) == orc.get_hits(vec, pairs, top_k=(1, 10))
    anchors = kg["train"][:64, 0]
    want = orc.nearest_negatives(anchors, vec, 125)
    out = vec.to(dev)
    for streamed in ("force", False):
        got = ops.l1_topk(out[torch.from_numpy(anchors).to(dev)], out, 1, 125, streamed=streamed)
        assert np.array_equal(got.reshape(-1).cpu().numpy(), want), streamed
    Lx, Rx = kg["x"][:2100], kg["x"][kg["e1"]:kg["e1"] + 2300]
    M = orc.l1_matrix(Lx, Rx)
    rmin, rarg, cmin, carg = ops.l1_argmins(torch.from_numpy(Lx).to(dev), torch.from_numpy(Rx).to(dev),
                                            block_bytes=2300 * 8 * 500)
    assert np.array_equal(rarg.cpu().numpy(), M.argmin(1)) and np.array_equal(carg.cpu().numpy(), M.argmin(0))
    assert np.array_equal(rmin.cpu().numpy(), M.min(1)) and np.array_equal(cmin.cpu().numpy(), M.min(0))


@pytest.mark.parametrize("nL,nR,d,skip,k", [(1, 1, 1, 0, 1), (70, 131, 300, 0, 10), (257, 64, 33, 1, 63),
                                            (300, 5, 16, 0, 10), (1500, 2100, 40, 1, 125), (5000, 3000, 24, 0, 10),
                                            (40000, 700, 8, 2, 3)])
def test_streamed_topk_and_ranks_with_ties(nL, nR, d, skip, k, dev):
    """The streamed kernels (distance matrix never stored) against NumPy's stable argsort of the oracle's fp64
    matrix and against the stored-matrix kernels: heavy exact ties (duplicated rows, coarse-grid coordinates),
    ragged tile edges, fewer columns than skip + k, one and many column segments."""
    from oracle import ea_oracle as orc
    from gnn_mtl_b200 import ops
    rng = np.random.default_rng(nL * 7 + nR)
    L = (rng.integers(-3, 4, (nL, d)) * 0.25).astype(np.float32)         # coarse grid -> many equal distances
    R = (rng.integers(-3, 4, (nR, d)) * 0.25).astype(np.float32)
    if nR > 20:
        R[5:12] = R[13:20]                                                # exact duplicate columns
    Lg, Rg = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)
    got = ops.l1_topk(Lg, Rg, skip, k, streamed="force").cpu().numpy()
    ref = ops.l1_topk(Lg, Rg, skip, k, streamed=False).cpu().numpy()
    assert np.array_equal(got, ref)
    sub = slice(0, min(nL, 600))
    order = np.argsort(orc.l1_matrix(L[sub], R), axis=1, kind="stable")
    want = np.full((order.shape[0], k), -1, dtype=np.int64)
    take = order[:, skip:skip + k]
    want[:, :take.shape[1]] = take
    assert np.array_equal(got[sub], want)
    n = min(nL, nR)
    if n > 1:
        rs, cs = ops.l1_ranks(Lg[:n], Rg[:n], streamed=True)                 # fp32 filter (queue overflows here:
        rx, cx = ops.l1_ranks(Lg[:n], Rg[:n], streamed=True, filtered=False)  # tie-heavy data) vs all-fp64 streamed
        rm, cm = ops.l1_ranks(Lg[:n], Rg[:n], streamed=False)
        assert torch.equal(rs, rm) and torch.equal(cs, cm)
        assert torch.equal(rx, rm) and torch.equal(cx, cm)
        if n <= 3000:
            rr, cr = orc.diagonal_ranks(orc.l1_matrix(L[:n], R[:n]))
            assert np.array_equal(rs.cpu().numpy(), rr) and np.array_equal(cs.cpu().numpy(), cr)


@pytest.mark.parametrize("n,d,noise", [(1, 300, 0.5), (130, 7, 0.5), (3000, 300, 0.05), (5000, 300, 3.0),
                                       (4097, 129, 1.0), (10500, 300, 0.5)])
def test_rank_filter_identical_to_exact(n, d, noise, dev):
    """fp32 candidate filter + exact fp64 decisions == all-fp64 kernels, on continuous data (near-ties only inside
    the rounding band), with duplicated rows / columns (exact ties), row blocks offset by row0, and through the
    sharded-row entry (row0 > 0)."""
    from oracle import ea_oracle as orc
    from gnn_mtl_b200 import ops
    rng = np.random.default_rng(n + d)
    L = rng.standard_normal((n, d)).astype(np.float32)
    R = (L + noise * rng.standard_normal((n, d))).astype(np.float32)
    if n > 200:
        R[100:110] = R[90:100]
        L[120:125] = L[130:135]
        R[150] = L[150]                      # zero distance on the diagonal
    Lg, Rg = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)
    rf, cf = ops.l1_ranks(Lg, Rg, filtered=True)
    rx, cx = ops.l1_ranks(Lg, Rg, filtered=False)
    assert torch.equal(rf, rx) and torch.equal(cf, cx)
    if n <= 5000:
        rr, cr = orc.diagonal_ranks(orc.l1_matrix(L, R))
        assert np.array_equal(rf.cpu().numpy(), rr) and np.array_equal(cf.cpu().numpy(), cr)
    if n >= 3000:                            # two row blocks, as the sharded evaluation calls it
        diag = ops.l1_paired(Lg, Rg)
        r2 = torch.zeros(n, dtype=torch.int32, device=dev)
        c2 = torch.zeros(n, dtype=torch.int32, device=dev)
        cut = n // 3 + 5
        ops.l1_rank_fused(Lg[:cut], 0, Rg, diag, r2, c2, filtered=True)
        ops.l1_rank_fused(Lg[cut:], cut, Rg, diag, r2, c2, filtered=True)
        assert torch.equal(r2, rx) and torch.equal(c2, cx)


# ------------------------------------------------------------------- sinkhorn --

def test_sinkhorn_golden(golden_dir, dev):
    from oracle import ea_oracle as orc
    from gnn_mtl_b200.utils.ot_loss import sinkhorn
    g = _load(golden_dir, "sinkhorn.npz")
    a, b, M = (torch.from_numpy(g[k]).to(dev) for k in ("a", "b", "M"))
    for tag, reg, iters in (("r05_i37", 0.05, 37), ("r01_i200", 0.01, 200), ("r1_conv", 0.5, 1000)):
        # fp64 cost in -> fp64 arithmetic: matches the reference's float64 to rounding
        P, loss = sinkhorn(a.double(), b.double(), M.double(), reg, numItermax=iters)
        assert P.dtype == torch.float64 and relerr(P, g["P_" + tag]) < 1e-9
        assert abs(float(loss) - float(g["loss_" + tag])) / abs(float(g["loss_" + tag])) < 1e-9
        # fp32 cost in -> fp32 arithmetic: the north_star bar
        info = {}
        P32, loss32 = sinkhorn(a, b, M, reg, numItermax=iters, info=info)
        assert relerr(P32, g["P_" + tag]) < REL, tag
        assert abs(float(loss32) - float(g["loss_" + tag])) / abs(float(g["loss_" + tag])) < REL


def test_sinkhorn_stop_rule_matches_oracle(dev):
    from oracle import ea_oracle as orc
    from gnn_mtl_b200.utils.ot_loss import sinkhorn
    torch.manual_seed(4)
    M = torch.cdist(torch.randn(60, 5), torch.randn(45, 5)).double()
    a = torch.full((60,), 1 / 60, dtype=torch.float64)
    b = torch.full((45,), 1 / 45, dtype=torch.float64)
    for thr in (1e-3, 1e-6, 1e-9):
        _, _, ref = orc.sinkhorn_scaling(a, b, M, 0.3, stopThr=thr, return_info=True)
        info = {}
        P, loss = sinkhorn(a.to(dev), b.to(dev), M.to(dev), 0.3, stopThr=thr, info=info)
        assert info["sweeps"] == ref["sweeps"], (thr, info["sweeps"], ref["sweeps"])
        assert relerr(info["log_u"], ref["log_u"]) < 1e-9


def test_sinkhorn_iteration_golden(golden_dir, dev):
    from gnn_mtl_b200.SinkhornOT import sinkhorn_iteration
    g = _load(golden_dir, "sinkhorn.npz")
    C = torch.from_numpy(g["C_cos"]).to(dev)
    mu = torch.full((1, 30, 1), 1 / 30, dtype=torch.float64, device=dev)
    nu = torch.full((1, 1, 36), 1 / 36, dtype=torch.float64, device=dev)
    for tag, eps, iters in (("e2", 1e-2, 100), ("e3_i25", 1e-3, 25)):
        w, k1, k2, K = sinkhorn_iteration(C, mu, nu, eps, numIterMax=iters)
        assert K.shape == (1, 30, 36) and relerr(K, g["S2_K_" + tag]) < 1e-8
        assert abs(float(w) - float(g["S2_w_" + tag])) / float(g["S2_w_" + tag]) < 1e-9
        assert abs(float(k1) - float(g["S2_kl1_" + tag])) < 1e-9
        assert abs(float(k2) - float(g["S2_kl2_" + tag])) < 1e-9
        w32, _, _, K32 = sinkhorn_iteration(C.float(), mu.float(), nu.float(), eps, numIterMax=iters)
        assert relerr(K32, g["S2_K_" + tag]) < 5e-4   # eps=1e-3 amplifies fp32 cost rounding by 1e3


@pytest.mark.parametrize("algo", ["simt", "tcgen05"])
@pytest.mark.parametrize("cost,reg", [("l2", 0.05), ("sqeuclid", 0.1), ("cos", 0.02)])
def test_sinkhorn_fused_vs_oracle(cost, reg, algo, dev):
    from oracle import ea_oracle as orc
    from gnn_mtl_b200.utils.ot_loss import sinkhorn_fused
    torch.manual_seed(7)
    X, Y = torch.randn(333, 300) * 0.06, torch.randn(270, 300) * 0.06
    a = torch.rand(333) + 0.5
    b = torch.rand(270) + 0.5
    b = b * a.sum() / b.sum()
    Mfn = {"l2": orc.cost_l2, "sqeuclid": orc.cost_sqeuclid, "cos": orc.cost_cosine}[cost]
    M = Mfn(X.double(), Y.double())
    P_ref, loss_ref = orc.sinkhorn_scaling(a, b, M, reg, numItermax=40)
    P, loss = sinkhorn_fused(X.to(dev), Y.to(dev), a.to(dev), b.to(dev), reg, numItermax=40, cost=cost,
                             algo=algo, return_plan=True)
    assert relerr(P, P_ref) < REL
    assert abs(float(loss) - float(loss_ref)) / abs(float(loss_ref)) < REL


@pytest.mark.parametrize("cost,reg", [("l2", 0.05), ("sqeuclid", 0.1), ("cos", 0.02)])
@pytest.mark.parametrize("I,J,d", [(200, 180, 300), (333, 1030, 128), (70, 65, 37)])
def test_fused_ot_loss_is_differentiable(cost, reg, I, J, d, dev):
    """sinkhorn_fused_loss: value and gradients (through the cost only, plan detached — models/models_ea.py:218-224)
    against fp64 autograd on the materialised cost with the oracle's plan; aligned pairs included (close-pair path)."""
    from oracle import ea_oracle as orc
    from gnn_mtl_b200.utils.ot_loss import sinkhorn_fused_loss
    torch.manual_seed(I + d)
    X = torch.randn(I, d) * 0.06
    Y = torch.randn(J, d) * 0.06
    Y[:40] = X[:40] + 0.004 * torch.randn(40, d)
    a = torch.rand(I) + 0.5
    b = torch.rand(J) + 0.5
    b = b * a.sum() / b.sum()
    Xd, Yd = X.double().requires_grad_(True), Y.double().requires_grad_(True)
    M = {"l2": orc.cost_l2, "sqeuclid": orc.cost_sqeuclid, "cos": orc.cost_cosine}[cost](Xd, Yd)
    P, _ = orc.sinkhorn_scaling(a, b, M.detach(), reg, numItermax=60, stopThr=-1.0)
    loss_ref = (P * M).sum()
    loss_ref.backward()
    Xg, Yg = X.to(dev).requires_grad_(True), Y.to(dev).requires_grad_(True)
    loss = sinkhorn_fused_loss(Xg, Yg, a.to(dev), b.to(dev), reg, numItermax=60, stopThr=-1.0, cost=cost)
    (3.0 * loss).backward()
    assert abs(float(loss) - float(loss_ref)) / abs(float(loss_ref)) < REL
    assert relerr(Xg.grad / 3.0, Xd.grad) < 5 * REL and relerr(Yg.grad / 3.0, Yd.grad) < 5 * REL


@pytest.mark.parametrize("nA,nB,d", [(1, 1, 8), (129, 257, 300), (700, 1300, 128), (2500, 900, 52)])
def test_fused_lse_tcgen05_matches_fp64(nA, nB, d, dev):
    """TMA + tcgen05 3xTF32 cost tiles against an fp64 materialised reference, ragged edges included."""
    from gnn_mtl_b200 import _lib, ops
    torch.manual_seed(nA)
    X, Y = torch.randn(nA, d, device=dev) * 0.1, torch.randn(nB, d, device=dev) * 0.1
    pot = torch.randn(nB, device=dev)
    lw = torch.randn(nA, device=dev)
    A = ops.FusedOperand(X, _lib.COST_L2, _lib.ALGO_TCGEN05)
    B = ops.FusedOperand(Y, _lib.COST_L2, _lib.ALGO_TCGEN05)
    got_pot, got = ops.lse_fused(A, B, _lib.COST_L2, 25.0, pot, lw, _lib.ALGO_TCGEN05, want_lse=True)
    ref = torch.logsumexp(pot.double()[None, :] - torch.cdist(X.double(), Y.double()) * 25.0, 1)
    assert float((got.double() - ref).abs().max()) < 2e-5       # absolute error of a log-sum: 2e-5 ~ 2e-5 relative in P
    assert float((got_pot.double() - (lw.double() - ref)).abs().max()) < 2e-5


def test_wasserstein_loss_as_shipped(golden_dir, dev):
    from gnn_mtl_b200.models.models_ea import UEAModel
    g = _load(golden_dir, "sinkhorn.npz")
    X, Y = torch.from_numpy(g["X"]), torch.from_numpy(g["Y"])[:40]

    class _U(UEAModel):
        def __init__(self):
            torch.nn.Module.__init__(self)
    out = torch.cat([X, Y]).to(dev).requires_grad_(True)
    data = {"e1": 40, "e2": 40, "index1": {i: i for i in range(40)}, "index2": {i: i + 40 for i in range(40)}}
    np.random.seed(0)
    loss = _U().get_loss_wassertein(out, data, 40, numItermax=50)
    np.random.seed(0)
    Lp, Rp = np.random.permutation(40)[:40], np.random.permutation(40)[:40]
    want = torch.cdist(X[Lp], Y[Rp])[:, 0].sum()
    assert abs(float(loss) - float(want)) / float(want) < 1e-5
    loss.backward()
    assert out.grad is not None and float(out.grad.abs().sum()) > 0


# ------------------------------------------------------------------- C ABI ----

def test_abi_error_behaviour(dev):
    from gnn_mtl_b200 import _lib
    lib = _lib.lib
    assert lib.eg_device_check() == 0
    assert lib.eg_spmm(None, None, None, 4, None, 8, 0, None, None, None, None, 0, None, None, None, 0,
                       None, None, 0, None, None) == -1
    assert lib.eg_topk_rows(None, 10, 1, 10, 0, 4096, None, None) == -4
    assert lib.eg_l1_matrix(None, 0, None, 5, 3, None, 5, None) == 0     # empty input is a no-op
    with pytest.raises(_lib.EagraftError):
        from gnn_mtl_b200 import ops
        ops.l1_matrix(torch.zeros(2, 2), torch.zeros(2, 2))               # CPU tensors are refused


@pytest.mark.parametrize("m,k,n,n1", [(1, 4, 4, 4), (130, 300, 600, 300), (777, 52, 260, 128), (3000, 300, 300, 300),
                                      (1537, 300, 600, 300), (40001, 300, 300, 300), (20000, 52, 164, 100),
                                      (5000, 300, 1000, 300)])
def test_gemm_nt_3xtf32_matches_fp64(m, k, n, n1, dev):
    """layers' dense products on tcgen05: fp32-accurate (3xTF32) against an fp64 reference."""
    from gnn_mtl_b200 import ops
    torch.manual_seed(m)
    A = torch.randn(m, k, device=dev)
    B = torch.randn(n, k, device=dev) * 0.1
    bias = torch.randn(n, device=dev)
    res = ops.gemm_nt([A], B, bias, n1=n1)
    got = torch.cat(res, 1) if isinstance(res, tuple) else res
    ref = A.double() @ B.double().t() + bias.double()
    assert relerr(got, ref) < 1e-5      # 3xTF32: ~2e-6 at k = 300
    # K-concatenated operand: [A1 | A2] · Bᵀ
    A2 = torch.randn(m, 36, device=dev)
    B2 = torch.randn(n, k + 36, device=dev) * 0.1
    got2 = ops.gemm_nt([A, A2], B2)
    ref2 = torch.cat([A, A2], 1).double() @ B2.double().t()
    assert relerr(got2, ref2) < 1e-5


@pytest.mark.parametrize("m,k,n,n1", [(1000, 300, 300, 300), (4099, 300, 600, 300), (257, 128, 128, 128),
                                      (130, 52, 340, 20), (200000, 300, 300, 300)])
def test_gemm_nt_chained_is_fp32_accurate(m, k, n, n1, dev):
    """Short-chain variant (eg_gemm_nt_3xtf32_chained, the ReLU-feeding x·Wᵀ + b of layers/layers.py:32,61):
    error at the level of an fp32 SIMT product (cuBLAS fp32 on the same operands is the yardstick), several times
    below the long-chain kernel's."""
    from gnn_mtl_b200 import ops
    torch.manual_seed(m + n)
    A = torch.randn(m, k, device=dev)
    B = torch.randn(n, k, device=dev) * 0.1
    bias = torch.randn(n, device=dev)
    ref = A.double() @ B.double().t() + bias.double()
    res = ops.gemm_nt([A], B, bias, n1=n1, chained=True)
    got = torch.cat(res, 1) if isinstance(res, tuple) else res
    plain = ops.gemm_nt([A], B, bias, n1=n1)
    plain = torch.cat(plain, 1) if isinstance(plain, tuple) else plain
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        simt = torch.addmm(bias, A, B.t())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    e_chain, e_plain, e_simt = relerr(got, ref), relerr(plain, ref), relerr(simt, ref)
    print("m=%d k=%d n=%d: chained %.2e  long-chain %.2e  cuBLAS fp32 %.2e" % (m, k, n, e_chain, e_plain, e_simt))
    assert e_chain < 5e-7
    assert e_chain < max(2.0 * e_simt, 2e-7)
    # mean absolute error too (the long chain's error is a one-sided bias; the short chain's must not be)
    assert float((got.double() - ref).mean().abs()) < 3e-7 * float(ref.abs().mean()) + 1e-9


@pytest.mark.parametrize("m,k,k2,n,n1", [(1, 16, 0, 4, 4), (130, 52, 36, 340, 20), (4099, 300, 300, 300, 300), (777, 36, 0, 260, 128),
                                         (128 * 149 + 5, 300, 0, 600, 300)])
def test_gemm_nt_raw_operands_bit_identical(m, k, k2, n, n1, dev):
    """eg_gemm_nt_3xtf32_raw (hi/lo split inside the kernel; layers/layers.py:61,69 products and their dx) gives the
    bits of eg_split_tf32 + eg_gemm_nt_3xtf32: ragged K (no padding of the A operands), two A operands, two outputs,
    and the fused addend."""
    from gnn_mtl_b200 import ops
    torch.manual_seed(m + k)
    parts = [torch.randn(m, k, device=dev)] + ([torch.randn(m, k2, device=dev)] if k2 else [])
    B = torch.randn(n, k + k2, device=dev) * 0.1
    bias = torch.randn(n, device=dev)
    want = ops.gemm_nt(parts, B, bias, n1=n1)
    got = ops.gemm_nt_raw(parts, B, bias, n1=n1)
    cat = lambda r: torch.cat(r, 1) if isinstance(r, tuple) else r
    assert torch.equal(cat(got), cat(want))
    ref = torch.cat(parts, 1).double() @ B.double().t() + bias.double()
    assert relerr(cat(got), ref) < 1e-5
    add = torch.randn(m, n, device=dev)
    assert torch.equal(ops.gemm_nt_raw(parts, B, bias, addend=add), cat(ops.gemm_nt(parts, B, bias)) + add)


@pytest.mark.parametrize("pair", [1, 0])
def test_gemm_nt_raw_both_kernels_and_raw_stage_release(pair, dev):
    """Both kernels behind eg_gemm_nt_3xtf32_raw — CTA pairs on tcgen05.mma.cta_group::2 (debug knob 18 = 1, the default)
    and the single-SM kernel (0) — give the split-operand bits at the benched height, on every run.  One column tile per
    row tile (n = 160) at 200k rows is the shape that exposed an early release of the raw A stage (the arrive that frees
    the stage overtook the loads of it, and the TMA refill raced the read: a handful of rows per run held the k-block
    ten stages ahead); the release now sits behind the tensor-memory stores that consume the loaded registers."""
    from gnn_mtl_b200 import _lib, ops
    torch.manual_seed(3)
    m, k = 200000, 300
    A = torch.randn(m, k, device=dev)
    try:
        _lib.lib.eg_debug_set(18, pair)
        for n in (160, 600):
            B = torch.randn(n, k, device=dev) * 0.1
            want = ops.gemm_nt([A], B)
            for _ in range(3):
                got = ops.gemm_nt_raw([A], B)
                bad = int((got != want).any(1).sum())
                assert bad == 0, "%d rows differ (n = %d, pair = %d)" % (bad, n, pair)
    finally:
        _lib.lib.eg_debug_set(18, 1)


@pytest.mark.parametrize("shape", [(1031, 300), (257, 7), (64, 4)])
@pytest.mark.parametrize("act_relu,gated", [(True, True), (False, True), (True, False)])
def test_epilogue_bwd_float4_and_scalar_paths(shape, act_relu, gated, dev):
    """Backward of the SpMM epilogue (layers/layers.py:65-72: activation, then the highway mix): the float4 kernel
    (16-byte aligned operands, element count % 4 == 0) and the scalar kernel (anything else) against the formula."""
    from gnn_mtl_b200 import _lib, ops
    torch.manual_seed(shape[0] + shape[1])
    g = torch.randn(*shape, device=dev)
    a = torch.randn(*shape, device=dev)
    pre = torch.randn(*shape, device=dev) if gated else None
    x = torch.randn(*shape, device=dev) if gated else None
    act = _lib.ACT_RELU if act_relu else _lib.ACT_IDENTITY
    dS, dG, dX = ops.epilogue_bwd(g, a, pre, x, act, True, True)
    t = torch.sigmoid(pre) if gated else torch.ones_like(g)
    want_dS = g * t
    if act_relu:
        want_dS = torch.where(a > 0, want_dS, torch.zeros_like(want_dS))
    assert torch.allclose(dS, want_dS, rtol=1e-6, atol=1e-7)
    if gated:
        assert torch.allclose(dG, g * (a - x) * (t * (1 - t)), rtol=1e-5, atol=1e-7)
        assert torch.allclose(dX, g * (1 - t), rtol=1e-5, atol=1e-7)


def test_tcgen05_kernels_cta_pairs_match_single_cta(dev):
    """The kernels of sinkhorn_tc.cu run on CTA pairs (tcgen05.mma.cta_group::2, debug knob 19 = 1, the default) or on
    single CTAs (0): fused LSE half-sweep (utils/ot_loss.py:58-70 on the fly), plan statistics, the NT GEMM and its
    short-chain variant give identical bits either way, on ragged shapes and odd row-tile counts."""
    from gnn_mtl_b200 import _lib, ops
    torch.manual_seed(11)
    knob = _lib.lib.eg_debug_set
    try:
        for nA, nB in ((100, 300), (129, 257), (1000, 5000), (3001, 2999)):
            X = torch.randn(nA, 300, device=dev) * 0.06
            Y = torch.randn(nB, 300, device=dev) * 0.06
            if nA > 3000:
                Y[:nB] = X[:nB] + 0.006 * torch.randn(nB, 300, device=dev)      # close pairs: exact re-evaluation branch
            pot = torch.randn(nB, device=dev)
            got = []
            for on in (0, 1):
                knob(19, on)
                A = ops.FusedOperand(X, _lib.COST_L2, _lib.ALGO_TCGEN05)
                B = ops.FusedOperand(Y, _lib.COST_L2, _lib.ALGO_TCGEN05)
                got.append(ops.lse_fused(A, B, _lib.COST_L2, 20.0, pot, None, _lib.ALGO_TCGEN05, want_lse=True)[1].clone())
            assert torch.equal(got[0], got[1]), (nA, nB)
            ref = torch.logsumexp(pot[None, :].double() - 20.0 * torch.cdist(X.double(), Y.double()), 1)
            assert float((got[1].double() - ref).abs().max()) < 2e-4
        for m, k, n in ((130, 52, 340), (1, 16, 4), (128 * 149 + 5, 300, 600)):
            a = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev) * 0.1; b = torch.randn(n, device=dev)
            for chained in (False, True):
                knob(19, 0); r0 = ops.gemm_nt([a], w, b, chained=chained)
                knob(19, 1); r1 = ops.gemm_nt([a], w, b, chained=chained)
                assert torch.equal(r0, r1), (m, k, n, chained)
    finally:
        knob(19, 1)


def test_wasserstein_loss_solve_on_side_stream(golden_dir, dev):
    """get_loss_wassertein (models/models_ea.py:206-224) with every sweep forced (stopThr < 0) runs its Sinkhorn solve
    on a side stream next to the caller's backward pass; loss and gradient are those of the serialised call, the solve
    is joined by join_pending_solve(), and with a stop rule (host read-back) the call stays on the caller's stream."""
    from gnn_mtl_b200.models import models_ea
    g = _load(golden_dir, "sinkhorn.npz")
    X, Y = torch.from_numpy(g["X"]), torch.from_numpy(g["Y"])[:40]

    class _U(models_ea.UEAModel):
        def __init__(self):
            torch.nn.Module.__init__(self)
    data = {"e1": 40, "e2": 40, "index1": np.arange(40), "index2": np.arange(40) + 40}
    sample = (torch.arange(40, device=dev), torch.arange(40, device=dev) + 40)
    res = []
    saved = models_ea.OVERLAP_SINKHORN
    try:
        for ov in (False, True):
            models_ea.OVERLAP_SINKHORN = ov
            m = _U()
            out = torch.cat([X, Y]).to(dev).requires_grad_(True)
            loss = m.get_loss_wassertein(out, data, 40, numItermax=200, stopThr=-1.0, sample=sample)
            assert (getattr(m, "_pending_solve", None) is not None) == ov
            loss.backward()
            m.join_pending_solve()
            assert getattr(m, "_pending_solve", None) is None
            torch.cuda.synchronize()
            res.append((float(loss), out.grad.clone()))
            loss2 = m.get_loss_wassertein(out, data, 40, numItermax=50, sample=sample)       # stop rule: serialised
            assert getattr(m, "_pending_solve", None) is None and abs(float(loss2) - float(loss)) < 1e-9
    finally:
        models_ea.OVERLAP_SINKHORN = saved
    assert res[0][0] == res[1][0] and torch.equal(res[0][1], res[1][1])
    # closed-form gradient of sum_i M[i, 0] against autograd through torch.cdist: same loss bits, gradient to rounding
    saved_cf = models_ea.CLOSED_FORM_COL0_GRAD
    try:
        models_ea.CLOSED_FORM_COL0_GRAD = False
        m = _U()
        out = torch.cat([X, Y]).to(dev).requires_grad_(True)
        loss = m.get_loss_wassertein(out, data, 40, numItermax=200, stopThr=-1.0, sample=sample)
        loss.backward()
        m.join_pending_solve()
    finally:
        models_ea.CLOSED_FORM_COL0_GRAD = saved_cf
    assert float(loss) == res[0][0]
    assert float((out.grad - res[0][1]).abs().max()) <= 1e-5 * float(out.grad.abs().max())


def test_margin_loss_golden_and_scale(golden_dir, dev):
    """Fused gather + L1 + hinge loss (models/models_ea.py:103-123): value and gradient."""
    from oracle import ea_oracle as orc
    from gnn_mtl_b200 import ops
    g = _load(golden_dir, "margin.npz")
    out = torch.from_numpy(g["out"]).to(dev).requires_grad_(True)
    k = int(g["k"])
    loss = ops.margin_loss(out, g["ILL"][:, 0], g["ILL"][:, 1], g["neg_left"], g["neg_right"], g["neg2_left"],
                           g["neg2_right"], k)
    (3.0 * loss).backward()
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 1e-5
    assert relerr(out.grad, 3.0 * g["grad"]) < REL
    # DBP15K-shaped: 600 anchors x 125 negatives x 300-d through the model entry point
    from gnn_mtl_b200.models.models_ea import _margin_loss
    rng = np.random.default_rng(0)
    n, d, t, k = 5000, 300, 600, 125
    x = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32) * 0.05)
    ILL = np.stack([rng.permutation(2500)[:t], rng.permutation(2500)[:t] + 2500], 1)
    nl, n2r = np.repeat(ILL[:, 0], k).astype(np.float64), np.repeat(ILL[:, 1], k).astype(np.float64)
    nr, n2l = rng.integers(0, n, t * k), rng.integers(0, n, t * k)
    xr = x.clone().requires_grad_(True)
    want = orc.margin_loss(xr, ILL, nl, nr, n2l, n2r, k)
    want.backward()
    xg = x.to(dev).requires_grad_(True)
    got = _margin_loss(xg, ILL, nl, nr, n2l, n2r, k)
    got.backward()
    assert abs(float(got) - float(want)) / float(want) < 1e-5
    assert relerr(xg.grad, xr.grad) < REL


def test_gromov_wasserstein_projection_golden(golden_dir, dev):
    """§8f rank 3: gw_iterative_1 (SinkhornOT/iterative_projection.py:119-120) on the log-domain kernels."""
    from gnn_mtl_b200.SinkhornOT import gw_iterative_1
    g = _load(golden_dir, "gw.npz")
    C1, C2 = torch.from_numpy(g["C1"]).to(dev), torch.from_numpy(g["C2"]).to(dev)
    mu = torch.full((18,), 1 / 18, dtype=torch.float64, device=dev)
    nu = torch.full((22,), 1 / 22, dtype=torch.float64, device=dev)
    T, gw = gw_iterative_1(C1, C2, mu, nu, epsilon=0.02, max_iter=6)
    assert T.shape == (1, 18, 22)
    assert relerr(T, g["T"]) < 1e-8 and abs(float(gw) - float(g["gw"])) / abs(float(g["gw"])) < 1e-9
    T2, log = gw_iterative_1(C1, C2, mu, nu, epsilon=0.02, max_iter=3, log=True)
    assert len(log["err"]) == 3 and log["gw_dist"] > 0


@pytest.mark.parametrize("algo", ["simt", "tcgen05"])
@pytest.mark.parametrize("cost", ["l2", "sqeuclid", "cos"])
def test_fused_lse_close_pairs(cost, algo, dev):
    """Aligned entities are CLOSE pairs: |a|²+|b|²-2a·b cancels there and would turn the 3xTF32 dot error into
    1e-3 of the cost; the kernels re-evaluate such pairs from the fp32 rows.  Checked against fp64."""
    from gnn_mtl_b200 import _lib, ops
    torch.manual_seed(5)
    n, d = 3000, 300
    X = torch.randn(n, d, device=dev) / d ** 0.5
    Y = X[torch.randperm(n, device=dev)] + 0.05 * torch.randn(n, d, device=dev) / d ** 0.5
    Y[:10] = X[:10]                                            # exact duplicates too
    cid = {"l2": _lib.COST_L2, "sqeuclid": _lib.COST_SQEUCLID, "cos": _lib.COST_COSINE}[cost]
    aid = {"simt": _lib.ALGO_SIMT, "tcgen05": _lib.ALGO_TCGEN05}[algo]
    A, B = ops.FusedOperand(X, cid, aid), ops.FusedOperand(Y, cid, aid)
    pot = torch.randn(n, device=dev)
    inv = 20.0 if cost == "l2" else 100.0
    _, got = ops.lse_fused(A, B, cid, inv, pot, None, aid, want_pot=False, want_lse=True)
    Xd, Yd = X.double(), Y.double()
    if cost == "cos":
        C = 1 - (Xd / Xd.norm(dim=1, keepdim=True)) @ (Yd / Yd.norm(dim=1, keepdim=True)).t()
    else:
        C = torch.cdist(Xd, Yd)
        C = C * C if cost == "sqeuclid" else C
    ref = torch.logsumexp(pot.double()[None, :] - C * inv, 1)
    # the exponent is cost*inv: a cost evaluated in fp32 carries ~1e-6 absolute error, amplified by inv
    assert float((got.double() - ref).abs().max()) < max(3e-5, 1.5e-6 * inv)


def test_one_graph_adjacency_with_id_remap(dev):
    """get_sparse_tensor_for_one_graph (utils/data_utils.py:339-350): ids remapped through index_R."""
    from oracle import ea_oracle as orc
    from gnn_mtl_b200.utils.data_utils import get_sparse_tensor_for_one_graph, sparse_mx_to_torch_sparse_tensor
    rng = np.random.default_rng(2)
    ids = rng.permutation(5000)[:300] + 100          # arbitrary global ids of one KG
    index_R = {int(v): i for i, v in enumerate(ids)}
    KG = [(int(ids[a]), int(r), int(ids[b])) for a, r, b in zip(rng.integers(0, 300, 900), rng.integers(0, 9, 900),
                                                               rng.integers(0, 300, 900))]
    adj = get_sparse_tensor_for_one_graph(300, KG, index_R, device=dev)
    h = np.array([index_R[t[0]] for t in KG]); t = np.array([index_R[t[2]] for t in KG])
    crow, col, val = orc.adjacency_csr(300, h, t)
    assert np.array_equal(adj.crow.cpu().numpy(), crow) and np.array_equal(adj.col64.cpu().numpy(), col)
    assert np.array_equal(adj.val.cpu().numpy().view(np.uint32), val.view(np.uint32))
    coo = sparse_mx_to_torch_sparse_tensor(adj)
    assert coo.is_sparse and coo.is_cuda and coo.shape == (300, 300) and coo._nnz() == len(col)


# ------------------------------------------------------------------ GAT (§8f rank 4) --

def test_gat_layers_golden(golden_dir, dev):
    """SpGraphAttentionLayer / GraphAttentionLayer against the live-reference fixture (outputs + all grads)."""
    from gnn_mtl_b200.adjacency import DeviceAdjacency
    from gnn_mtl_b200.layers.att_layers import GraphAttentionLayer, SpGraphAttentionLayer
    g = _load(golden_dir, "gat.npz")
    adj = DeviceAdjacency.from_triples(int(g["n_ent"]), g["triples"], device=dev).to_torch_coo()
    single = SpGraphAttentionLayer(24, 15, 0.0, 0.2, F.elu).to(dev)
    with torch.no_grad():
        single.W.copy_(torch.from_numpy(g["s_W"]))
        single.a.copy_(torch.from_numpy(g["s_a"]))
    x = torch.from_numpy(g["x"]).to(dev).requires_grad_(True)
    y = single(x, adj)
    (y * torch.from_numpy(g["s_seed"]).to(dev)).sum().backward()
    assert relerr(y, g["s_y"]) < REL
    assert relerr(x.grad, g["s_dx"]) < REL
    assert relerr(single.W.grad, g["s_dW"]) < REL
    assert relerr(single.a.grad, g["s_da"]) < REL
    multi = GraphAttentionLayer(24, 9, 0.0, F.relu, 0.2, 4, True).to(dev)
    with torch.no_grad():
        for i, head in enumerate(multi.attentions):
            head.W.copy_(torch.from_numpy(g["m_W%d" % i]))
            head.a.copy_(torch.from_numpy(g["m_a%d" % i]))
    x = torch.from_numpy(g["x"]).to(dev).requires_grad_(True)
    y, adj_out = multi((x, adj))
    assert adj_out is adj
    (y * torch.from_numpy(g["m_seed"]).to(dev)).sum().backward()
    assert relerr(y, g["m_y"]) < REL
    assert relerr(x.grad, g["m_dx"]) < REL
    for i, head in enumerate(multi.attentions):
        assert relerr(head.W.grad, g["m_dW%d" % i]) < REL
        assert relerr(head.a.grad, g["m_da%d" % i]) < REL


@pytest.mark.parametrize("n,d,alpha", [(700, 75, 0.2), (3000, 300, 0.2), (500, 7, 0.01), (400, 512, 0.3)])
def test_gat_aggregate_vs_oracle_fp64(n, d, alpha, dev):
    """Kernel-level check on a power-law graph with hub rows: forward, dh, ds1, ds2 against the fp64 oracle,
    with and without an edge-dropout scale."""
    from gnn_mtl_b200 import ops, synth
    from gnn_mtl_b200.adjacency import DeviceAdjacency
    from oracle import ea_oracle as orc
    heads, tails = synth.make_powerlaw_graph(n, 12, seed=n + d)
    chain = np.arange(n - 1)
    heads, tails = np.concatenate([heads, chain]), np.concatenate([tails, chain + 1])   # no isolated entity
    A = DeviceAdjacency.from_heads_tails(n, torch.from_numpy(heads).to(dev), torch.from_numpy(tails).to(dev))
    adj_cpu = orc.adjacency_torch_coo(n, heads, tails)
    gen = torch.Generator().manual_seed(3)
    h = torch.randn(n, d, generator=gen, dtype=torch.float64)
    a = torch.randn(1, 2 * d, generator=gen, dtype=torch.float64) / d ** 0.5
    seed = torch.randn(n, d, generator=gen, dtype=torch.float64)
    # oracle with W = I so the aggregation alone is exercised
    h64 = h.clone().requires_grad_(True)
    a64 = a.clone().requires_grad_(True)
    y64 = orc.gat_layer(h64, adj_cpu, torch.eye(d, dtype=torch.float64), a64, alpha, "identity")
    (y64 * seed).sum().backward()
    hg = h.float().to(dev).requires_grad_(True)
    ag = a.float().to(dev).requires_grad_(True)
    s1, s2 = hg @ ag[0, :d], hg @ ag[0, d:]
    y = ops.gat_aggregate(hg, s1, s2, A, alpha)
    (y * seed.float().to(dev)).sum().backward()
    assert relerr(y, y64.detach()) < REL
    assert relerr(hg.grad, h64.grad) < REL
    assert relerr(ag.grad, a64.grad) < REL
    # edge scale: compare with a dense torch evaluation on the GPU in fp64
    nnz = A.csr.nnz
    m = (torch.rand(nnz, generator=gen) > 0.3).float().to(dev) / 0.7
    crow = A.csr.rowptr.long()
    rows = torch.repeat_interleave(torch.arange(n, device=dev), crow[1:] - crow[:-1])
    cols = A.csr.col.long()
    hd = h.to(dev).requires_grad_(True)
    t = (hd @ a.to(dev)[0, :d])[rows] + (hd @ a.to(dev)[0, d:])[cols]
    w = torch.exp(-F.leaky_relu(t, alpha))
    W = torch.zeros(n, dtype=torch.float64, device=dev).index_add(0, rows, w)
    yd = torch.zeros(n, d, dtype=torch.float64, device=dev).index_add(0, rows, (w * m.double())[:, None] * hd[cols])
    yd = yd / W[:, None]
    (yd * seed.to(dev)).sum().backward()
    hg2 = h.float().to(dev).requires_grad_(True)
    y2 = ops.gat_aggregate(hg2, hg2 @ ag.detach()[0, :d], hg2 @ ag.detach()[0, d:], A, alpha, edge_scale=m)
    (y2 * seed.float().to(dev)).sum().backward()
    assert relerr(y2, yd.detach()) < REL
    assert relerr(hg2.grad, hd.grad) < REL


def test_gat_encoder_trains(dev):
    """model2encoder['GAT'] (4 heads x 75) wired as in models/encoders.py:69-86 runs fwd/bwd and matches the oracle."""
    from types import SimpleNamespace
    from gnn_mtl_b200 import synth
    from gnn_mtl_b200.adjacency import DeviceAdjacency
    from gnn_mtl_b200.models.encoders import model2encoder
    from oracle import ea_oracle as orc
    n = 900
    heads, tails = synth.make_powerlaw_graph(n, 8, seed=8)
    chain = np.arange(n - 1)
    heads, tails = np.concatenate([heads, chain]), np.concatenate([tails, chain + 1])
    adj = DeviceAdjacency.from_heads_tails(n, torch.from_numpy(heads).to(dev), torch.from_numpy(tails).to(dev)).to_torch_coo()
    args = SimpleNamespace(num_layers=3, dim=300, feat_dim=300, act="relu", n_heads=4, alpha=0.2, dropout=0.0,
                           bias=1, cuda=0, device=dev, task="ea")
    torch.manual_seed(0)
    enc = model2encoder["GAT"](args).to(dev)
    x = (torch.randn(n, 300) * 0.3).to(dev).requires_grad_(True)
    y = enc.encode(x, adj)
    assert y.shape == (n, 300)
    y.square().sum().backward()
    adj_cpu = orc.adjacency_torch_coo(n, heads, tails)
    xc = x.detach().cpu().double().requires_grad_(True)
    hcur = xc
    for layer in enc.layers:
        hs = [(att.W.detach().cpu().double(), att.a.detach().cpu().double()) for att in layer.attentions]
        hcur = orc.gat_multihead(hcur, adj_cpu, hs, 0.2, "relu", True)
    hcur.square().sum().backward()
    assert relerr(y, hcur.detach()) < REL
    assert relerr(x.grad, xc.grad) < 5e-4


@pytest.mark.parametrize("K,m,n", [(1, 4, 4), (37, 300, 300), (5000, 300, 300), (20011, 128, 160), (3333, 20, 516),
                                   (200000, 300, 300)])
def test_gemm_tn_3xtf32_matches_fp64(K, m, n, dev):
    """dW-shaped product sum_k A[k, m] B[k, n] on the MN-major split-K tcgen05 kernel against fp64; ragged K, tile
    edges, one and many K splits; bit-identical across repeats (fixed-order split reduction)."""
    from gnn_mtl_b200 import ops
    gen = torch.Generator().manual_seed(K + m)
    A = torch.randn(K, m, generator=gen)
    B = torch.randn(K, n, generator=gen)
    want = A.double().t() @ B.double()
    Ag, Bg = A.to(dev), B.to(dev)
    sa, sb = ops.split_tf32(Ag, ops._pad16(m)), ops.split_tf32(Bg, ops._pad16(n))
    got = ops.gemm_tn(sa, m, sb, n)
    scale = float(want.abs().max())
    assert float((got.double().cpu() - want).abs().max()) < 2e-6 * max(scale, 1.0) * max(1.0, (K / 1000.0) ** 0.5)
    assert torch.equal(got, ops.gemm_tn(sa, m, sb, n))
