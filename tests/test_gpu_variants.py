"""GPU: every kernel variant / dispatch branch behind the same entry points gives the same answer
(on-chip vs persistent vs streaming Sinkhorn, vector vs scalar SpMM, fused vs un-fused activation, batched S2)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def relerr(got, want):
    got = torch.as_tensor(got).detach().double().cpu()
    want = torch.as_tensor(want).detach().double().cpu()
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _problem(I, J, seed=0):
    torch.manual_seed(seed)
    X, Y = torch.randn(I, 16) * 0.3, torch.randn(J, 16) * 0.3
    a = torch.rand(I) + 0.5
    b = torch.rand(J) + 0.5
    return torch.cdist(X, Y), a, b * a.sum() / b.sum()


# (I, J): on-chip fp32 kernel with register rows / without; J % 4 != 0 (streaming); J > 4096 (generic persistent)
@pytest.mark.parametrize("I,J", [(3000, 3000), (3100, 2996), (2000, 2400), (700, 1001), (300, 4200), (5000, 800)])
def test_sinkhorn_dispatch_branches_agree_with_oracle(I, J, dev):
    from oracle import ea_oracle as orc
    from gnn_mtl_b200 import _lib
    from gnn_mtl_b200.utils.ot_loss import sinkhorn
    M, a, b = _problem(I, J, seed=I)
    _, loss_ref, ref = orc.sinkhorn_scaling(a, b, M, 0.1, numItermax=25, return_info=True)
    results = []
    for knobs in ((1, 1, 1), (1, 1, 0), (1, 0, 0), (0, 0, 0)):      # persistent, resident rows, on-chip
        for key, val in zip((3, 4, 5), knobs):
            _lib.lib.eg_debug_set(key, val)
        info = {}
        _, loss = sinkhorn(a.to(dev), b.to(dev), M.to(dev), 0.1, numItermax=25, return_plan=False, info=info)
        assert info["sweeps"] == 25
        assert relerr(info["log_u"], ref["log_u"]) < 2e-5 and relerr(info["log_v"], ref["log_v"]) < 2e-5, knobs
        assert abs(float(loss) - float(loss_ref)) / float(loss_ref) < 1e-4
        results.append(info["log_u"])
    for key in (3, 4, 5):
        _lib.lib.eg_debug_set(key, 1)
    for r in results[1:]:
        assert relerr(r, results[0]) < 1e-5


@pytest.mark.parametrize("I,J,reg,iters", [(3000, 3000, 0.01, 120), (2500, 3000, 0.02, 64), (1200, 1600, 0.05, 40)])
def test_sinkhorn_scaling_domain_continuation(I, J, reg, iters, dev):
    """The scaling-domain on-chip kernel (default for the reference's 3000 x 3000 batch) against the fp64 oracle
    and against the all-log-domain kernel; with the fold-into-kernel step forced to fire (tiny threshold) and
    with the redo-in-log-domain path forced."""
    from oracle import ea_oracle as orc
    from gnn_mtl_b200 import _lib
    from gnn_mtl_b200.utils.ot_loss import sinkhorn
    M, a, b = _problem(I, J, seed=J + iters)
    M = M * 1.5
    _, loss_ref, ref = orc.sinkhorn_scaling(a, b, M, reg, numItermax=iters, stopThr=-1.0, return_info=True)
    dbg = _lib.lib.eg_debug_set
    out, raw = {}, {}
    try:
        for name, knobs in (("scaling", {}), ("rowblock", {12: 0}), ("log", {7: 0}), ("absorb_often", {10: 10}),
                            ("absorb_often_rowblock", {10: 10, 12: 0}), ("forced_redo", {11: 1})):
            for key, val in knobs.items():
                dbg(key, val)
            fallbacks0, folds0 = dbg(8, 0), dbg(9, 0)
            info = {}
            _, loss = sinkhorn(a.to(dev), b.to(dev), M.to(dev), reg, numItermax=iters, stopThr=-1.0, return_plan=False,
                               info=info)
            assert info["sweeps"] == iters, name
            # potentials are fixed up to the (c, -c) shift the iteration itself leaves free: compare u_i + v_j
            got = info["log_u"].double().cpu()[:, None] + info["log_v"].double().cpu()[None, :64]
            want = ref["log_u"].double()[:, None] + ref["log_v"].double()[None, :64]
            # absolute, log units.  The scaling-domain kernels carry O(1) scalings relative to fixed potentials; the
            # log-domain safety path carries the O(100) potentials themselves in fp32 (ulp 1.5e-5 at 128..256, a few
            # of those per sweep survive): documented 2e-3
            log_path = name in ("log", "forced_redo")
            assert float((got - want).abs().max()) < (2e-3 if log_path else 1e-4), name
            assert abs(float(loss) - float(loss_ref)) / abs(float(loss_ref)) < 1e-4, name
            if name.startswith("absorb_often"):
                assert dbg(9, 0) > folds0                 # the fold step really ran
            if name == "forced_redo":
                assert dbg(8, 0) == fallbacks0 + 1
            if name == "scaling":
                assert dbg(8, 0) == fallbacks0           # and did not need the redo
            out[name] = got
            raw[name] = info["log_u"].double().cpu()
            for key in knobs:
                dbg(key, {7: 1, 10: 32000, 11: 0, 12: 1}[key])
    finally:
        dbg(7, 1); dbg(10, 32000); dbg(11, 0); dbg(12, 1)
    assert float((raw["forced_redo"] - raw["log"]).abs().max()) == 0.0     # the redo IS the log-domain kernel
    # the scaling-domain variants agree with each other to 1e-4 absolute in log units (each is within 1e-4 of the
    # fp64 oracle above); the fold step does not change that
    assert float((out["scaling"] - out["rowblock"]).abs().max()) < 1e-4
    assert float((out["absorb_often"] - out["scaling"]).abs().max()) < 1e-4
    assert float((out["absorb_often_rowblock"] - out["rowblock"]).abs().max()) < 1e-4


def test_sinkhorn_fp64_persistent_vs_streaming(dev):
    from oracle import ea_oracle as orc
    from gnn_mtl_b200 import _lib
    from gnn_mtl_b200.utils.ot_loss import sinkhorn
    M, a, b = _problem(900, 1100, seed=5)
    P_ref, _ = orc.sinkhorn_scaling(a, b, M, 0.05, numItermax=60)
    for persistent in (1, 0):
        _lib.lib.eg_debug_set(3, persistent)
        P, _ = sinkhorn(a.double().to(dev), b.double().to(dev), M.double().to(dev), 0.05, numItermax=60)
        assert relerr(P, P_ref) < 1e-9
    _lib.lib.eg_debug_set(3, 1)


def test_sinkhorn_empty_weights_mean_uniform(dev):
    from gnn_mtl_b200.utils.ot_loss import sinkhorn
    M, _, _ = _problem(64, 80)
    empty = torch.zeros(0, device=dev)
    P, _ = sinkhorn(empty, empty, M.to(dev), 0.2, numItermax=200, stopThr=1e-7)
    assert abs(float(P.sum()) - 1.0) < 1e-5
    assert float((P.sum(1) - 1 / 64).abs().max()) < 1e-6


def test_sinkhorn_iteration_batched(dev):
    from oracle import ea_oracle as orc
    from gnn_mtl_b200.SinkhornOT import sinkhorn_iteration
    torch.manual_seed(2)
    C = torch.rand(3, 20, 24, dtype=torch.float64)
    mu = torch.full((3, 20, 1), 1 / 20, dtype=torch.float64)
    nu = torch.full((3, 1, 24), 1 / 24, dtype=torch.float64)
    # the reference's stopping test is scalar-only (a batch > 1 makes its `if` ambiguous, sinkhorn_loss.py:211),
    # so the batched call here solves each element independently; compare element by element
    w, k1, k2, K = sinkhorn_iteration(C.to(dev), mu.to(dev), nu.to(dev), 0.05, numIterMax=40)
    assert K.shape == (3, 20, 24) and w.shape == (3,)
    for i in range(3):
        w_ref, k1_ref, k2_ref, K_ref = orc.sinkhorn_stabilised(C[i:i + 1], mu[i:i + 1], nu[i:i + 1], 0.05, numIterMax=40)
        assert relerr(K[i], K_ref[0]) < 1e-9 and abs(float(w[i]) - float(w_ref)) / float(w_ref) < 1e-10
        assert abs(float(k1[i]) - float(k1_ref)) < 1e-10 and abs(float(k2[i]) - float(k2_ref)) < 1e-10


def test_layer_forward_only_dropout_and_unfused_activation(dev):
    from oracle import ea_oracle as orc
    from gnn_mtl_b200.adjacency import DeviceAdjacency
    from gnn_mtl_b200.layers.layers import GraphConvolution, HighWayGraphConvolution
    from gnn_mtl_b200.synth import make_kg_pair
    kg = make_kg_pair("tiny", dim=64)
    adj = DeviceAdjacency.from_triples(kg["n"], kg["triples"], device=dev).to_torch_coo()
    adj_cpu = orc.adjacency_torch_coo(kg["n"], kg["triples"][:, 0], kg["triples"][:, 2])
    x = torch.from_numpy(kg["x"])
    torch.manual_seed(0)
    # activation the epilogue does not know (elu): aggregate fused, activate outside
    layer = HighWayGraphConvolution(64, 64, 0.0, F.elu, True, -1, "cpu")
    want = orc.highway_layer(x, adj_cpu, layer.linear.weight.detach(), layer.linear.bias.detach(),
                             layer.kernel_gate, layer.bias_gate, F.elu)
    layer.to(dev)
    xg = x.to(dev).requires_grad_(True)
    got, _ = layer((xg, adj))
    assert relerr(got, want) < 1e-4
    got.sum().backward()
    assert xg.grad is not None
    # eval mode under no_grad: nothing is saved for backward
    gc = GraphConvolution(64, 32, 0.0, F.relu, False).to(dev).eval()
    with torch.no_grad():
        y, _ = gc((x.to(dev), adj))
    want = orc.gcn_layer(x, adj_cpu, gc.linear.weight.detach().cpu(), None, "relu")
    assert relerr(y, want) < 1e-4
    # dropout > 0 runs in training mode and is the identity in eval mode
    do = GraphConvolution(64, 32, 0.5, F.relu, True).to(dev)
    do.train()
    y1, _ = do((x.to(dev), adj))
    do.eval()
    y2, _ = do((x.to(dev), adj))
    y3, _ = do((x.to(dev), adj))
    assert y1.shape == y2.shape and torch.equal(y2, y3) and not torch.equal(y1, y2)


def test_sparse_feature_input_and_dense_adjacency_branch(dev):
    """Layer 0 of the reference gets the feature matrix as a sparse COO tensor (data_utils.py:358,397);
    a dense adjacency goes down the torch.mm branch (layers.py:36-37)."""
    from gnn_mtl_b200.adjacency import DeviceAdjacency
    from gnn_mtl_b200.layers.layers import HighWayGraphConvolution
    from gnn_mtl_b200.synth import make_kg_pair
    kg = make_kg_pair("tiny", dim=32)
    adj = DeviceAdjacency.from_triples(kg["n"], kg["triples"], device=dev).to_torch_coo()
    x = torch.from_numpy(kg["x"]).to(dev)
    torch.manual_seed(1)
    layer = HighWayGraphConvolution(32, 32, 0.0, F.relu, True, 0, dev).to(dev)
    y_dense_x, _ = layer((x, adj))
    y_sparse_x, _ = layer((x.to_sparse(), adj))
    assert torch.equal(y_dense_x, y_sparse_x)
    y_dense_adj, _ = layer((x, adj.to_dense()))
    assert relerr(y_dense_adj, y_dense_x) < 1e-4


def test_get_hits_single_pair_and_cpu_input(dev):
    from gnn_mtl_b200.utils.eval_utils import get_hits, eval_at_1
    vec = torch.randn(10, 8)
    hits = get_hits(vec, np.array([[2, 7]]), top_k=(1, 10))
    assert hits == {"Hits@1_l": 100.0, "Hits@10_l": 100.0, "Hits@1_r": 100.0, "Hits@10_r": 100.0}
    assert float(eval_at_1(vec.to(dev), {"test": np.array([[2, 7], [3, 3]])})) in (50.0, 100.0)


def test_drivers_end_to_end_on_dbp15k_layout(tmp_path, dev):
    """§8f rank 2 + drivers: synthetic pair written in the DBP15K on-disk layout, loaded by the mirrored
    loaders, trained for a few epochs by both schedules."""
    from gnn_mtl_b200.config import make_args
    from gnn_mtl_b200.run.train_ea import train_ea
    from gnn_mtl_b200.run.train_unsup_ea import train_unsup_ea
    from gnn_mtl_b200.synth import make_kg_pair, write_dbp15k_dir
    kg = make_kg_pair("tiny", dim=32)
    root = str(tmp_path / "dbp15k")
    write_dbp15k_dir(kg, root, "zh_en")
    quiet = lambda *a, **k: None
    args = make_args(model="HGCN", epochs=3, refine_epochs=2, batch_size=100, neg_num=5, dim=32, data_root=root)
    model, hist = train_unsup_ea(args, log=quiet)
    assert len(hist["wasserstein"]) == 3 and len(hist["refine"]) == 2
    assert all(np.isfinite(l) for l, _ in hist["wasserstein"] + hist["refine"])
    assert set(hist["test"]) == {"Hits@1_l", "Hits@1_r"}
    args = make_args(model="GCN", epochs=4, neg_num=5, dim=32, min_epochs=2, data_root=root)
    model, metrics = train_ea(args, log=quiet)
    assert metrics["Hits@1_l"] >= 0


def test_mrr_from_ranks(dev):
    from oracle import ea_oracle as orc
    from gnn_mtl_b200.utils.eval_utils import get_hits
    rng = np.random.default_rng(4)
    vec = rng.standard_normal((400, 16)).astype(np.float32)
    pairs = np.stack([np.arange(150), np.arange(150) + 200], 1)
    vec[pairs[:, 1]] = vec[pairs[:, 0]] + 0.9 * rng.standard_normal((150, 16)).astype(np.float32)
    m = get_hits(torch.from_numpy(vec).to(dev), pairs, top_k=(1, 10), mrr=True)
    rr, cr = orc.diagonal_ranks(orc.l1_matrix(vec[pairs[:, 0]], vec[pairs[:, 1]]))
    assert abs(m["MRR_l"] - float(np.mean(1.0 / (rr + 1)))) < 1e-12
    assert abs(m["MRR_r"] - float(np.mean(1.0 / (cr + 1)))) < 1e-12
    assert list(m.keys()) == ["Hits@1_l", "Hits@10_l", "Hits@1_r", "Hits@10_r", "MRR_l", "MRR_r"]


def test_sharded_adjacency_halo_single_process(dev):
    """parallel.ShardedAdjacency(halo=True) in a one-rank world: the remapped CSR over [own rows | fetched rows] (nothing
    to fetch) must reproduce the plain SpMM and the layer gradients; halo="auto" keeps the plan (no remote rows)."""
    from gnn_mtl_b200 import ops, parallel as par
    from gnn_mtl_b200.adjacency import DeviceAdjacency
    from gnn_mtl_b200.synth import make_kg_pair
    kg = make_kg_pair("tiny", dim=64)
    full = DeviceAdjacency.from_triples(kg["n"], kg["triples"], device=dev)
    sh = par.ShardedAdjacency(full, halo=True)
    assert sh.halo and sh.plan.n_need == 0 and sh.r0 == 0 and sh.r1 == kg["n"]
    H = torch.randn(kg["n"], 64, device=dev)
    want = ops.spmm(full.csr, H)[0]
    got = ops.spmm(sh.csr, sh.gather(H))[0]
    assert torch.equal(got, want)
    got_t = ops.spmm(sh.csr_t, sh.gather(H, transposed=True))[0]
    assert torch.equal(got_t, ops.spmm(full.csr_t, H)[0])
    assert par.ShardedAdjacency(full, halo="auto").halo
