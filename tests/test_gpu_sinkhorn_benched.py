"""GPU: the Sinkhorn solve exactly as bench.py runs it — 3000 x 3000, reg 0.01, a = b = ones (reference
models/models_ea.py:217-220), 1000 sweeps, default kernel path — against the fp64 oracle
(utils/ot_loss.py:26-76) at the level of the PLAN: max|dP| / max|P| <= 1e-4, loss <= 1e-4 relative, potentials
(up to the (c, -c) shift the iteration leaves free) <= 1e-4 absolute in log units."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

PLAN_TOL = 1e-4       # BASELINE.json north_star: "OT loss/plan within 1e-4 relative (fp32)"
POT_TOL = 1e-4        # absolute, log units


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _bench_embeddings(dev):
    """Outputs of the (untrained, seeded) 3-layer H-GCN on the benchmark graph: what get_loss_wassertein samples from
    in bench.py."""
    from gnn_mtl_b200.adjacency import DeviceAdjacency
    from gnn_mtl_b200.layers.layers import HighWayGraphConvolution
    from gnn_mtl_b200.synth import make_kg_pair
    torch.manual_seed(10086)
    kg = make_kg_pair("dbp100k")
    adj = DeviceAdjacency.from_triples(kg["n"], kg["triples"], device=dev).to_torch_coo()
    h = torch.from_numpy(kg["x"]).to(dev)
    with torch.no_grad():
        for act in (F.relu, F.relu, (lambda z: z)):
            layer = HighWayGraphConvolution(300, 300, 0.0, act, True, -1, "cpu").to(dev)
            h, _ = layer((h, adj))
    return kg, h


def _costs(dev):
    kg, out = _bench_embeddings(dev)
    rng = np.random.default_rng(5)
    L = torch.from_numpy(rng.permutation(kg["e1"])[:3000]).to(dev)
    R = torch.from_numpy(rng.permutation(kg["e2"])[:3000] + kg["e1"]).to(dev)
    cases = {"bench": torch.cdist(out[L], out[R], p=2)}
    # trained-looking: unit-norm embeddings whose aligned partner is the nearest neighbour but not by a wide margin
    g = torch.Generator().manual_seed(3)
    X = F.normalize(torch.randn(3000, 300, generator=g), dim=1)
    Y = F.normalize(X[torch.randperm(3000, generator=g)] + 1.2 * torch.randn(3000, 300, generator=g) / 300 ** 0.5, dim=1)
    cases["aligned"] = torch.cdist(X.to(dev), Y.to(dev), p=2)
    # test split of the graph's own features (linked pairs are close: x[e2] = x[e1] + noise)
    t = torch.from_numpy(kg["test"][:3000]).to(dev)
    x = torch.from_numpy(kg["x"]).to(dev)
    cases["linked_features"] = torch.cdist(x[t[:, 0]], x[t[:, 1]], p=2)
    return cases


@pytest.fixture(scope="module")
def costs(dev):
    return _costs(dev)


@pytest.mark.parametrize("name", ["bench", "aligned", "linked_features"])
@pytest.mark.parametrize("tile2d", [1, 0])
def test_plan_parity_at_the_benched_configuration(name, tile2d, costs, dev):
    from oracle import ea_oracle as orc
    from gnn_mtl_b200 import _lib
    from gnn_mtl_b200.utils.ot_loss import sinkhorn
    M = costs[name]
    a = torch.ones(3000)
    P_ref, loss_ref, ref = orc.sinkhorn_scaling(a, a, M.cpu(), 0.01, numItermax=1000, stopThr=-1.0, return_info=True)
    redo0 = _lib.lib.eg_debug_set(8, 0)
    _lib.lib.eg_debug_set(12, tile2d)
    try:
        info = {}
        P, loss = sinkhorn(a.to(dev), a.to(dev), M, 0.01, numItermax=1000, stopThr=-1.0, info=info)
    finally:
        _lib.lib.eg_debug_set(12, 1)
    assert info["sweeps"] == 1000
    assert _lib.lib.eg_debug_set(8, 0) == redo0, "the scaling-domain solve fell back to the log-domain kernel"
    P = P.cpu()
    plan_err = float((P - P_ref).abs().max() / P_ref.abs().max())
    loss_err = abs(float(loss) - float(loss_ref)) / abs(float(loss_ref))
    got = info["log_u"].double().cpu()[:, None] + info["log_v"].double().cpu()[None, :]
    want = ref["log_u"].double()[:, None] + ref["log_v"].double()[None, :]
    # only where the plan has mass do the potentials matter (and are they determined)
    mask = P_ref > 1e-6 * P_ref.max()
    pot_err = float((got - want)[mask].abs().max())
    row_err = float((P.sum(1) - P_ref.sum(1)).abs().max())
    print("%s tile2d=%d: plan %.2e loss %.2e potentials %.2e row-marginal %.2e" % (name, tile2d, plan_err, loss_err,
                                                                                  pot_err, row_err))
    assert plan_err <= PLAN_TOL and loss_err <= PLAN_TOL, (plan_err, loss_err)
    assert pot_err <= POT_TOL, pot_err


def test_stop_rule_with_the_default_kernel(costs, dev):
    """reference default stopThr = 1e-9 on the default path: same sweep count (or within one check) and same plan as
    the oracle when the rule fires; sweep count and error are read back."""
    from oracle import ea_oracle as orc
    from gnn_mtl_b200.utils.ot_loss import sinkhorn
    torch.manual_seed(4)
    M = torch.rand(3000, 3000) * 0.5
    a = torch.ones(3000) / 3000
    for thr in (1e-5, 1e-7):
        P_ref, loss_ref, ref = orc.sinkhorn_scaling(a, a, M, 0.05, numItermax=1000, stopThr=thr, return_info=True)
        info = {}
        P, loss = sinkhorn(a.to(dev), a.to(dev), M.to(dev), 0.05, numItermax=1000, stopThr=thr, info=info)
        assert ref["sweeps"] < 1000
        assert info["sweeps"] == ref["sweeps"], (thr, info["sweeps"], ref["sweeps"])
        assert info["err"] <= thr
        assert float((P.cpu() - P_ref).abs().max() / P_ref.abs().max()) <= PLAN_TOL
        assert abs(float(loss) - float(loss_ref)) / float(loss_ref) <= PLAN_TOL
