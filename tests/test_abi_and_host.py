"""CPU: the C-ABI library loads and exports every symbol include/eagraft.h
declares (no compute without a GPU), plus host-side logic."""
import os
import re

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "eagraft.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gnn_mtl_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(_lib.lib, name), name
        assert name in _lib.SIGNATURES, "binding missing for " + name
    assert set(_lib.SIGNATURES) == set(declared)
    assert _lib.lib.eg_version() >= 100
    assert _lib.lib.eg_strerror(-3) == b"workspace too small"


def test_no_cpu_path():
    import pytest
    from gnn_mtl_b200 import _lib, ops
    with pytest.raises(_lib.EagraftError):
        ops.l1_matrix(torch.zeros(3, 4), torch.zeros(3, 4))
    if not torch.cuda.is_available():
        assert _lib.lib.eg_device_check() != 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gnn_mtl_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                assert "oracle" not in src, os.path.join(base, f)


def test_activation_classifier():
    from gnn_mtl_b200 import _lib
    from gnn_mtl_b200.layers.layers import classify_activation
    assert classify_activation(F.relu) == _lib.ACT_RELU
    assert classify_activation(lambda x: x) == _lib.ACT_IDENTITY
    assert classify_activation(torch.relu) == _lib.ACT_RELU
    assert classify_activation(F.elu) is None
    assert classify_activation(torch.tanh) is None


def test_long_row_segmentation_host_logic():
    from gnn_mtl_b200.adjacency import _Csr
    deg = np.array([3, 0, 1300, 512, 513, 7])
    rowptr = torch.from_numpy(np.concatenate([[0], np.cumsum(deg)]).astype(np.int32))
    nnz = int(rowptr[-1])
    csr = _Csr(6, 6, rowptr, torch.zeros(nnz, dtype=torch.int32), torch.zeros(nnz), threshold=512)
    assert csr.n_long == 2 and csr.long_rows.tolist() == [2, 4]
    assert csr.long_first.tolist() == [0, 3, 5]
    assert csr.seg_begin.tolist() == [3, 515, 1027, 1815, 2327]
    assert csr.seg_end.tolist() == [515, 1027, 1303, 2327, 2328]
    assert csr.seg_row.tolist() == [2, 2, 2, 4, 4]


def test_state_dict_keys_match_reference_layout():
    from gnn_mtl_b200.layers.layers import HighWayGraphConvolution
    layer = HighWayGraphConvolution(6, 6, 0.0, F.relu, True, -1, "cpu")
    assert sorted(layer.state_dict().keys()) == ["linear.bias", "linear.weight"]   # gates are not saved
    assert layer.kernel_gate.shape == (6, 6) and float(layer.bias_gate.abs().sum()) == 0.0
    torch.manual_seed(3)
    a = HighWayGraphConvolution(5, 5, 0.0, F.relu, True, -1, "cpu")
    torch.manual_seed(3)
    lin = torch.nn.Linear(5, 5, True)
    gate = torch.FloatTensor(5, 5).uniform_(-np.sqrt(6.0 / 10), np.sqrt(6.0 / 10))
    assert torch.equal(a.linear.weight, lin.weight) and torch.equal(a.kernel_gate, gate)   # RNG order parity


def test_synth_shapes():
    from gnn_mtl_b200.synth import make_kg_pair
    kg = make_kg_pair("tiny", dim=16)
    assert kg["x"].shape == (610, 16) and kg["triples"].shape == (2700, 3)
    assert len(kg["train"]) == 60 and len(kg["test"]) == 140
    assert kg["triples"][:, [0, 2]].max() < 610
    kg2 = make_kg_pair("tiny", dim=16)
    assert np.array_equal(kg["triples"], kg2["triples"]) and np.array_equal(kg["x"], kg2["x"])


def test_column0_loss_closed_form_gradient_matches_autograd_through_cdist():
    """models/models_ea.py:218-224 as shipped is sum_i cdist(X, Y)[i, 0]; the closed-form backward used by
    get_loss_wassertein gives the gradient autograd derives through torch.cdist (fp64, CPU)."""
    import torch
    from gnn_mtl_b200.models.models_ea import _Column0Loss
    torch.manual_seed(0)
    X = torch.randn(50, 30, dtype=torch.float64, requires_grad=True)
    Y = torch.randn(40, 30, dtype=torch.float64, requires_grad=True)
    want_loss = torch.cdist(X, Y, p=2)[:, 0].sum()
    want = torch.autograd.grad(want_loss, [X, Y])
    with torch.no_grad():
        m0 = torch.cdist(X, Y, p=2)[:, 0].contiguous()
    got_loss = _Column0Loss.apply(X, Y, m0)
    got = torch.autograd.grad(got_loss * 3.0, [X, Y])
    assert float(want_loss.detach() - got_loss.detach()) == 0.0
    assert float((3.0 * want[0] - got[0]).abs().max()) < 1e-12 and float((3.0 * want[1] - got[1]).abs().max()) < 1e-12
    assert float(got[1][1:].abs().max()) == 0.0
