"""CPU: the oracle restatement reproduces the fixtures the live reference wrote
(tests/golden/make_golden.py).  This is the pin the GPU parity tests rest on."""
import os

import numpy as np
import torch

from oracle import ea_oracle as orc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_adjacency_csr_bit_exact(golden_dir):
    g = _load(golden_dir, "adjacency.npz")
    tri = g["triples"]
    crow, col, val = orc.adjacency_csr(int(g["n_ent"]), tri[:, 0], tri[:, 2])
    assert np.array_equal(crow, g["crow"])
    assert np.array_equal(col, g["col"])
    assert np.array_equal(val.view(np.uint32), g["val"].view(np.uint32))


def test_adjacency_insertion_order_bit_exact(golden_dir):
    g = _load(golden_dir, "adjacency.npz")
    r, c, v = orc.adjacency_insertion_order([tuple(t) for t in g["triples"].tolist()])
    assert np.array_equal(r, g["coo_row"]) and np.array_equal(c, g["coo_col"])
    assert np.array_equal(v.view(np.uint64), g["coo_val"].view(np.uint64))


def _adj(g):
    tri = g["triples"]
    return orc.adjacency_torch_coo(int(g["n_ent"]), tri[:, 0], tri[:, 2])


def test_layers_forward_backward(golden_dir):
    adj = _adj(_load(golden_dir, "adjacency.npz"))
    g = _load(golden_dir, "layers.npz")
    for name, act in (("gc", "relu"), ("hw", "relu"), ("hwid", "identity")):
        x = torch.from_numpy(g["x"]).requires_grad_(True)
        W = torch.from_numpy(g[name + "_W"]).requires_grad_(True)
        b = torch.from_numpy(g[name + "_b"]).requires_grad_(True)
        if name == "gc":
            y = orc.gcn_layer(x, adj, W, b, act)
        else:
            y = orc.highway_layer(x, adj, W, b, torch.from_numpy(g[name + "_G"]),
                                  torch.from_numpy(g[name + "_c"]), act)
        (y * torch.from_numpy(g[name + "_seed"])).sum().backward()
        np.testing.assert_allclose(y.detach().numpy(), g[name + "_y"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(x.grad.numpy(), g[name + "_dx"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(W.grad.numpy(), g[name + "_dW"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(b.grad.numpy(), g[name + "_db"], rtol=1e-5, atol=1e-5)


def test_sinkhorn_scaling(golden_dir):
    g = _load(golden_dir, "sinkhorn.npz")
    a, b, M = (torch.from_numpy(g[k]) for k in ("a", "b", "M"))
    for tag, reg, iters in (("r05_i37", 0.05, 37), ("r01_i200", 0.01, 200), ("r1_conv", 0.5, 1000)):
        P, loss = orc.sinkhorn_scaling(a, b, M, reg, numItermax=iters)
        np.testing.assert_allclose(P.numpy(), g["P_" + tag], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(loss.numpy(), g["loss_" + tag], rtol=1e-12)
    X, Y = torch.from_numpy(g["X"]), torch.from_numpy(g["Y"])
    shipped = orc.wasserstein_loss_as_shipped(X, Y)
    np.testing.assert_allclose(float(shipped), float(g["loss_shipped"]), rtol=1e-6)
    np.testing.assert_allclose(orc.cost_l2(X, Y).numpy(), g["M"], rtol=1e-6, atol=1e-7)


def test_sinkhorn_stabilised(golden_dir):
    g = _load(golden_dir, "sinkhorn.npz")
    C = torch.from_numpy(g["C_cos"])
    mu = torch.full((1, 30, 1), 1 / 30, dtype=torch.float64)
    nu = torch.full((1, 1, 36), 1 / 36, dtype=torch.float64)
    for tag, eps, iters in (("e2", 1e-2, 100), ("e3_i25", 1e-3, 25)):
        w, k1, k2, K = orc.sinkhorn_stabilised(C, mu, nu, eps, numIterMax=iters)
        np.testing.assert_allclose(K.numpy(), g["S2_K_" + tag], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(float(w), float(g["S2_w_" + tag]), rtol=1e-12)
        np.testing.assert_allclose(float(k1), float(g["S2_kl1_" + tag]), rtol=1e-9, atol=1e-15)
        np.testing.assert_allclose(float(k2), float(g["S2_kl2_" + tag]), rtol=1e-9, atol=1e-15)
    Va, Vb = torch.from_numpy(g["Va"]), torch.from_numpy(g["Vb"])
    np.testing.assert_allclose(orc.cost_cosine(Va, Vb).numpy(), g["C_cos"][0], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(orc.cost_sqeuclid(Va, Vb).numpy(), g["C_sq"], rtol=1e-12)


def test_s1_s2_known_answer_relation(golden_dir):
    """with uniform marginals and eps=1e-2 the two reference
    solvers agree on the plan — a POT-free cross-check of both restatements."""
    g = _load(golden_dir, "sinkhorn.npz")
    C = torch.from_numpy(g["C_cos"])
    mu = torch.full((1, 30, 1), 1 / 30, dtype=torch.float64)
    nu = torch.full((1, 1, 36), 1 / 36, dtype=torch.float64)
    _, _, _, K = orc.sinkhorn_stabilised(C, mu, nu, 1e-2, numIterMax=1000, tol=1e-14)
    P, _ = orc.sinkhorn_scaling(mu.reshape(-1), nu.reshape(-1), C[0], 1e-2, numItermax=2000, stopThr=1e-15)
    assert float((K[0] - P).abs().max()) < 1e-9


def test_eval_entry_points(golden_dir):
    g = _load(golden_dir, "eval.npz")
    vec, pairs = torch.from_numpy(g["vec"]), g["pairs"]
    hits = orc.get_hits(vec, pairs, top_k=(1, 5, 10))
    assert list(hits.keys()) == [str(k) for k in g["hits_keys"]]
    assert list(hits.values()) == list(g["hits_vals"])
    assert float(orc.eval_at_1(vec, pairs)) == float(g["at1"])
    assert np.array_equal(orc.nearest_negatives(pairs[:, 0], vec, 7), g["neg"])
    mnn = orc.mutual_nearest_pairs(vec, np.arange(45), np.arange(45) + 45, 20)
    assert np.array_equal(mnn, g["mnn"])
    gw = orc.eval_matching_matrix(-g["T"], pairs, {i: i for i in range(45)},
                                  {i + 45: i for i in range(45)}, top_k=(1, 5))
    assert list(gw.values()) == list(g["gw_vals"])
    sim = orc.l1_matrix(g["vec"][pairs[:, 0]], g["vec"][pairs[:, 1]])
    a = orc.diagonal_ranks(sim)
    b = orc.diagonal_ranks_by_sort(sim)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_margin_loss(golden_dir):
    g = _load(golden_dir, "margin.npz")
    out = torch.from_numpy(g["out"]).requires_grad_(True)
    loss = orc.margin_loss(out, g["ILL"], g["neg_left"], g["neg_right"], g["neg2_left"], g["neg2_right"], int(g["k"]))
    loss.backward()
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-6)
    np.testing.assert_allclose(out.grad.numpy(), g["grad"], rtol=1e-5, atol=1e-8)


def test_gromov_wasserstein_projection(golden_dir):
    g = _load(golden_dir, "gw.npz")
    C1, C2 = torch.from_numpy(g["C1"]), torch.from_numpy(g["C2"])
    mu = torch.full((18,), 1 / 18, dtype=torch.float64)
    nu = torch.full((22,), 1 / 22, dtype=torch.float64)
    T, gw = orc.gw_iterative(C1, C2, mu, nu, 0.02, 6)
    np.testing.assert_allclose(T.numpy(), g["T"], rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(float(gw), float(g["gw"]), rtol=1e-10)


def test_gat_layer_forward_backward(golden_dir):
    """oracle.gat_layer / gat_multihead reproduce layers/att_layers.py outputs and gradients."""
    g = _load(golden_dir, "gat.npz")
    adj = _adj(g)
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    W = torch.from_numpy(g["s_W"]).requires_grad_(True)
    a = torch.from_numpy(g["s_a"]).requires_grad_(True)
    y = orc.gat_layer(x, adj, W, a, 0.2, "elu")
    (y * torch.from_numpy(g["s_seed"])).sum().backward()
    np.testing.assert_allclose(y.detach().numpy(), g["s_y"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), g["s_dx"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(W.grad.numpy(), g["s_dW"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(a.grad.numpy(), g["s_da"], rtol=1e-4, atol=1e-5)
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    heads = [(torch.from_numpy(g["m_W%d" % i]).requires_grad_(True), torch.from_numpy(g["m_a%d" % i]).requires_grad_(True))
             for i in range(4)]
    y = orc.gat_multihead(x, adj, heads, 0.2, "relu", concat=True)
    (y * torch.from_numpy(g["m_seed"])).sum().backward()
    np.testing.assert_allclose(y.detach().numpy(), g["m_y"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), g["m_dx"], rtol=1e-4, atol=1e-5)
    for i, (Wh, ah) in enumerate(heads):
        np.testing.assert_allclose(Wh.grad.numpy(), g["m_dW%d" % i], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(ah.grad.numpy(), g["m_da%d" % i], rtol=1e-4, atol=1e-5)
