"""CPU, build container only: oracle restatement vs the LIVE reference on larger
and randomised inputs (skipped where /root/reference is absent, e.g. the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import ea_oracle as orc
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")


def _ref_csr(n_ent, KG):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        du = ref_shim.ref.data_utils
        adj = du.sparse_mx_to_torch_sparse_tensor(du.get_sparse_tensor(n_ent, KG))
        csr = adj.coalesce().to_sparse_csr()
    return csr.crow_indices().numpy(), csr.col_indices().numpy(), csr.values().numpy()


@pytest.mark.parametrize("seed,n_ent,n_tri", [(0, 50, 0), (1, 1, 3), (2, 40, 300), (3, 500, 2500), (4, 2000, 9000)])
def test_adjacency_random(seed, n_ent, n_tri, capsys):
    rng = np.random.default_rng(seed)
    h = rng.integers(0, n_ent, n_tri)
    t = rng.integers(0, n_ent, n_tri)
    if n_tri > 20:
        t[:5] = h[:5]
        h[5:10], t[5:10] = t[10:15], h[10:15]
    KG = [(int(a), 0, int(b)) for a, b in zip(h, t)]
    if n_tri == 0:
        crow, col, val = orc.adjacency_csr(n_ent, h, t)
        assert crow.tolist() == [0] * (n_ent + 1) and len(col) == 0
        return
    rc, cc, vc = _ref_csr(n_ent, KG)
    crow, col, val = orc.adjacency_csr(n_ent, h, t)
    assert np.array_equal(crow, rc) and np.array_equal(col, cc)
    assert np.array_equal(val.view(np.uint32), vc.view(np.uint32))


def test_adjacency_dbp15k_shape():
    from gnn_mtl_b200.synth import make_kg_pair
    kg = make_kg_pair("dbp15k", features=False)
    tri = kg["triples"]
    KG = [tuple(r) for r in tri.tolist()]
    rc, cc, vc = _ref_csr(kg["n"], KG)
    crow, col, val = orc.adjacency_csr(kg["n"], tri[:, 0], tri[:, 2])
    assert np.array_equal(crow, rc) and np.array_equal(col, cc)
    assert np.array_equal(val.view(np.uint32), vc.view(np.uint32))


def test_sinkhorn_bsz300_reference_defaults():
    torch.manual_seed(0)
    X, Y = torch.randn(300, 32) * 0.1, torch.randn(300, 32) * 0.1
    M = torch.cdist(X, Y)
    a = b = torch.ones(300)
    P0, l0 = ref_shim.ref.ot_loss.sinkhorn(a, b, M, 0.01, numItermax=120)
    P1, l1 = orc.sinkhorn_scaling(a, b, M, 0.01, numItermax=120)
    assert torch.allclose(P0, P1, rtol=1e-12, atol=0) and torch.allclose(l0, l1, rtol=1e-12)


def test_hits_and_negatives_medium():
    rng = np.random.default_rng(5)
    vec = torch.from_numpy(rng.standard_normal((700, 24)).astype(np.float32))
    pairs = np.stack([rng.permutation(350)[:200], rng.permutation(350)[:200] + 350], 1)
    assert ref_shim.ref.eval_utils.get_hits(vec, pairs) == orc.get_hits(vec, pairs)

    class _A:
        n_nodes, device = 700, "cpu"
    neg = ref_shim.ref.models_ea.BaseModel(_A()).get_neg(pairs[:50, 0], vec, 25)
    assert np.array_equal(neg, orc.nearest_negatives(pairs[:50, 0], vec, 25))


def test_dbp15k_loader_host_part_matches_reference(tmp_path, monkeypatch):
    """§8f rank 2: the on-disk DBP15K layout parsed by read_dbp15k()/_split_links() vs the reference's
    load_data_ea (utils/data_utils.py:375-413) on a synthetic directory."""
    import types
    import warnings
    from gnn_mtl_b200.synth import make_kg_pair, write_dbp15k_dir
    from gnn_mtl_b200.utils import data_utils as mine
    kg = make_kg_pair("tiny", dim=12)
    write_dbp15k_dir(kg, str(tmp_path / "data" / "dbp15k"), "zh_en")
    monkeypatch.chdir(tmp_path)
    args = types.SimpleNamespace(dataset="zh_en", model="HGCN", task="ea")
    np.random.seed(7)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ref_shim.ref.data_utils.load_data_ea(args)
    np.random.seed(7)
    raw = mine.read_dbp15k("zh_en", "data/dbp15k")
    train, test = mine._split_links(raw["ill"])
    assert np.array_equal(train, ref["train"]) and np.array_equal(test, ref["test"])
    assert raw["kg1"] + raw["kg2"] == ref["triple"]
    assert torch.allclose(raw["x"], ref["x"].to_dense(), atol=1e-7)
    head, tail, head_r, tail_r = mine.rfunc(kg["n"], ref["triple"])
    assert head == ref["head"] and tail == ref["tail"]
    assert np.array_equal(head_r.toarray(), ref["head_r"]) and np.array_equal(tail_r.toarray(), ref["tail_r"])
    crow, col, val = orc.adjacency_csr(kg["n"], np.array(ref["triple"])[:, 0], np.array(ref["triple"])[:, 2])
    csr = ref["adj"].coalesce().to_sparse_csr()
    assert np.array_equal(csr.crow_indices().numpy(), crow) and np.array_equal(csr.values().numpy(), val)


@pytest.mark.parametrize("seed,n_ent,d_in,d_out,alpha", [(0, 120, 16, 12, 0.2), (1, 400, 30, 75, 0.2), (2, 60, 8, 5, 0.01)])
def test_gat_layer_random_graphs_vs_live_reference(seed, n_ent, d_in, d_out, alpha):
    """oracle.gat_layer against layers/att_layers.py::SpGraphAttentionLayer on random connected graphs: forward and
    all three gradients (the golden fixture pins one graph; this widens it where the reference can be imported)."""
    import warnings
    import torch.nn.functional as F
    rng = np.random.default_rng(seed)
    n_tri = 4 * n_ent
    h = rng.integers(0, n_ent, n_tri)
    t = rng.integers(0, n_ent, n_tri)
    chain = np.arange(n_ent - 1)                       # no isolated entity (the reference asserts on the NaN)
    h, t = np.concatenate([h, chain]), np.concatenate([t, chain + 1])
    KG = [(int(a), 0, int(b)) for a, b in zip(h, t)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        du = ref_shim.ref.data_utils
        adj = du.sparse_mx_to_torch_sparse_tensor(du.get_sparse_tensor(n_ent, KG))
        att = ref_shim.load("layers.att_layers")
        torch.manual_seed(seed)
        layer = att.SpGraphAttentionLayer(d_in, d_out, 0.0, alpha, F.elu)
        x = torch.randn(n_ent, d_in)
        xr = x.clone().requires_grad_(True)
        y_ref = layer(xr, adj)
        seed_t = torch.randn_like(y_ref)
        (y_ref * seed_t).sum().backward()
    xo = x.clone().requires_grad_(True)
    W = layer.W.detach().clone().requires_grad_(True)
    a = layer.a.detach().clone().requires_grad_(True)
    y = orc.gat_layer(xo, orc.adjacency_torch_coo(n_ent, h, t), W, a, alpha, "elu")
    (y * seed_t).sum().backward()
    np.testing.assert_allclose(y.detach().numpy(), y_ref.detach().numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(xo.grad.numpy(), xr.grad.numpy(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(W.grad.numpy(), layer.W.grad.numpy(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(a.grad.numpy(), layer.a.grad.numpy(), rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("seed,n,d,t,k", [(0, 90, 12, 20, 4), (1, 400, 300, 64, 25), (2, 50, 5, 7, 1)])
def test_margin_loss_random_vs_live_reference(seed, n, d, t, k):
    """oracle.margin_loss against EAModel.get_loss (models/models_ea.py:103-123): value and gradient."""
    rng = np.random.default_rng(seed)
    out = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32) * 0.3).requires_grad_(True)
    half = n // 2
    ILL = np.stack([rng.permutation(half)[:t], rng.permutation(half)[:t] + half], 1).astype(np.int64)

    class _Fake:
        pass
    me = _Fake()
    me.neg_num = k
    me.neg_left = np.repeat(ILL[:, 0], k)
    me.neg2_right = np.repeat(ILL[:, 1], k)
    me.neg_right = rng.integers(0, n, t * k)
    me.neg2_left = rng.integers(0, n, t * k)
    want = ref_shim.ref.models_ea.EAModel.get_loss(me, out, {"train": ILL}, "train")
    want.backward()
    out2 = out.detach().clone().requires_grad_(True)
    got = orc.margin_loss(out2, ILL, me.neg_left, me.neg_right, me.neg2_left, me.neg2_right, k)
    got.backward()
    np.testing.assert_allclose(float(got), float(want), rtol=1e-6)
    np.testing.assert_allclose(out2.grad.numpy(), out.grad.numpy(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("seed,I,J,eps,iters", [(0, 30, 40, 0.05, 60), (1, 64, 64, 0.01, 100), (2, 25, 18, 0.002, 40)])
def test_sinkhorn_iteration_random_vs_live_reference(seed, I, J, eps, iters):
    """oracle.sinkhorn_stabilised against SinkhornOT.sinkhorn_iteration (sinkhorn_loss.py:159-220), fp64, including
    a small-epsilon case where the reference's absorption / clamps act."""
    import contextlib
    import io
    rng = np.random.default_rng(seed)
    C = torch.from_numpy(rng.uniform(size=(1, I, J)))
    mu = torch.full((1, I, 1), 1.0 / I, dtype=torch.float64)
    nu = torch.full((1, 1, J), 1.0 / J, dtype=torch.float64)
    with contextlib.redirect_stdout(io.StringIO()):
        w_ref, k1_ref, k2_ref, K_ref = ref_shim.ref.sinkhorn_loss.sinkhorn_iteration(C, mu, nu, eps, numIterMax=iters,
                                                                                      tol=1e-12, debug=False)
    w, k1, k2, K = orc.sinkhorn_stabilised(C, mu, nu, eps, numIterMax=iters, tol=1e-12)
    np.testing.assert_allclose(K.numpy(), K_ref.numpy(), rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(float(w), float(w_ref), rtol=1e-10)


@pytest.mark.parametrize("seed,I,J,eps,iters", [(0, 14, 19, 0.05, 5), (1, 30, 30, 0.02, 8)])
def test_gw_projection_random_vs_live_reference(seed, I, J, eps, iters):
    """oracle.gw_iterative against SinkhornOT.iterative_projection.gw_iterative_1 (:8-60, :119-120), fp64."""
    import contextlib
    import io
    rng = np.random.default_rng(seed)
    Va = torch.from_numpy(rng.uniform(size=(I, 5)))
    Vb = torch.from_numpy(rng.uniform(size=(J, 5)))
    C1 = ref_shim.ref.cderivation.cos_dist_mat(Va, Va).double()
    C2 = ref_shim.ref.cderivation.cos_dist_mat(Vb, Vb).double()
    mu = torch.full((I,), 1.0 / I, dtype=torch.float64)
    nu = torch.full((J,), 1.0 / J, dtype=torch.float64)
    with contextlib.redirect_stdout(io.StringIO()):
        T_ref, gw_ref = ref_shim.ref.iterative_projection.gw_iterative_1(C1, C2, mu, nu, epsilon=eps, max_iter=iters)
        T, gw = orc.gw_iterative(C1, C2, mu, nu, eps, iters)
    np.testing.assert_allclose(T.numpy().reshape(I, J), T_ref.numpy().reshape(I, J), rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(float(gw), float(gw_ref), rtol=1e-9)
