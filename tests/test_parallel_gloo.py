"""CPU, world_size 2, gloo: the collective logic of gnn_mtl_b200/parallel.py driven by
torch-CPU stand-ins for the per-rank kernels (the kernels themselves are covered by -m gpu)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(fn, world=2):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    results = [q.get() for _ in range(world)]
    for r in results:
        assert r == "ok", r


def _entry(fn, rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        torch.set_num_threads(1)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        fn(rank, world)
        dist.barrier()
        dist.destroy_process_group()
        q.put("ok")
    except Exception as e:  # pragma: no cover
        import traceback
        q.put("rank %d: %s\n%s" % (rank, e, traceback.format_exc()))


def _w_sinkhorn(rank, world):
    from oracle import ea_oracle as orc
    from gnn_mtl_b200 import parallel as par
    torch.manual_seed(0)
    I, J = 37, 29                       # uneven split: 19 + 18 rows
    X, Y = torch.randn(I, 6).double() * 0.3, torch.randn(J, 6).double() * 0.3
    a = torch.rand(I).double() + 0.5
    b = torch.rand(J).double() + 0.5
    b = b * a.sum() / b.sum()
    M = torch.cdist(X, Y)
    reg = 0.2
    r0, r1 = par.shard_range(I, rank, world)
    Ml = M[r0:r1]

    def col_lse_local(log_u):
        return torch.logsumexp(log_u[:, None] - Ml / reg, 0)

    def row_update_local(log_v):
        return torch.log(a[r0:r1]) - torch.logsumexp(log_v[None, :] - Ml / reg, 1)

    for thr, iters in ((1e-9, 23), (1e-4, 1000)):
        lu, lv, sweeps, err = par.sharded_sinkhorn(col_lse_local, row_update_local, I, J, torch.log(b), b, r1 - r0,
                                                   "cpu", torch.float64, numItermax=iters, stopThr=thr)
        P_ref, _, info = orc.sinkhorn_scaling(a, b, M, reg, numItermax=iters, stopThr=thr, return_info=True)
        assert sweeps == info["sweeps"], (sweeps, info["sweeps"])
        P_loc = torch.exp(lu[:, None] + lv[None, :] - Ml / reg)
        assert torch.allclose(P_loc, P_ref[r0:r1], rtol=1e-9, atol=1e-14)


def _w_gather_and_grads(rank, world):
    from gnn_mtl_b200 import parallel as par
    n = 7
    r0, r1 = par.shard_range(n, rank, world)
    full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
    got = par.all_gather_rows(full[r0:r1], n)
    assert torch.equal(got, full)
    lse = par.combine_partial_lse(torch.tensor([0.0 + rank, 1.0, -float("inf")]))
    want = torch.logsumexp(torch.tensor([[0.0, 1.0, -float("inf")], [1.0, 1.0, -float("inf")]]), 0)
    assert torch.allclose(lse[:2], want[:2]) and lse[2] == -float("inf")
    lin = torch.nn.Linear(4, 3)
    lin.weight.grad = torch.full_like(lin.weight, float(rank + 1))
    lin.bias.grad = torch.full_like(lin.bias, float(10 * (rank + 1)))
    par.allreduce_grads(lin.parameters())
    assert torch.allclose(lin.weight.grad, torch.full_like(lin.weight, 1.5))
    assert torch.allclose(lin.bias.grad, torch.full_like(lin.bias, 15.0))


def _w_rank_merge(rank, world):
    from oracle import ea_oracle as orc
    from gnn_mtl_b200 import parallel as par
    rng = np.random.default_rng(3)
    n = 41
    L = rng.standard_normal((n, 5)).astype(np.float32)
    R = L + 0.7 * rng.standard_normal((n, 5)).astype(np.float32)
    sim = orc.l1_matrix(L, R)
    rr, cr = orc.diagonal_ranks(sim)
    r0, r1 = par.shard_range(n, rank, world)
    diag = np.diag(sim)
    blk = sim[r0:r1]
    gi = np.arange(r0, r1)
    row_local = ((blk < diag[gi, None]).sum(1) + ((blk == diag[gi, None]) & (np.arange(n)[None, :] < gi[:, None])).sum(1))
    col_part = ((blk < diag[None, :]).sum(0) + ((blk == diag[None, :]) & (gi[:, None] < np.arange(n)[None, :])).sum(0))
    rows, cols = par.merge_rank_counts(torch.from_numpy(row_local.astype(np.int32)),
                                       torch.from_numpy(col_part.astype(np.int32)), n)
    assert np.array_equal(rows.numpy(), rr) and np.array_equal(cols.numpy(), cr)


def _w_sharded_adjacency(rank, world):
    import scipy.sparse as sp
    from oracle import ea_oracle as orc
    from gnn_mtl_b200 import parallel as par
    from gnn_mtl_b200.adjacency import _Csr
    rng = np.random.default_rng(1)
    n = 53
    h, t = rng.integers(0, n, 300), rng.integers(0, n, 300)
    crow, col, val = orc.adjacency_csr(n, h, t)
    A = sp.csr_matrix((val, col, crow), shape=(n, n))
    At = A.T.tocsr(); At.sort_indices()

    class _Full:
        pass
    full = _Full()
    full.n, full.device = n, torch.device("cpu")
    full.csr = _Csr(n, n, torch.from_numpy(crow.astype(np.int32)), torch.from_numpy(col.astype(np.int32)),
                    torch.from_numpy(val), threshold=4)
    full.csr_t = _Csr(n, n, torch.from_numpy(At.indptr.astype(np.int32)), torch.from_numpy(At.indices.astype(np.int32)),
                      torch.from_numpy(At.data.astype(np.float32)), threshold=4)
    sh = par.ShardedAdjacency(full)
    H = torch.from_numpy(rng.standard_normal((n, 4)).astype(np.float32))
    Hg = sh.gather(sh.local(H))
    assert torch.equal(Hg, H)
    c = sh.csr
    loc = sp.csr_matrix((c.val.numpy(), c.col.numpy(), c.rowptr.numpy()), shape=(c.n_rows, n))
    assert np.allclose(loc @ H.numpy(), (A @ H.numpy())[sh.r0:sh.r1])
    # hub segmentation indices are local to the slice
    if c.n_seg:
        assert int(c.seg_begin.min()) >= 0 and int(c.seg_end.max()) <= c.nnz


def _w_halo_exchange(rank, world):
    """Needed-rows exchange: the remapped block times the received rows == the block of the full product, bit for
    bit (the remap is monotone, so the summation order inside a row does not change); only referenced rows travel."""
    import scipy.sparse as sp
    from oracle import ea_oracle as orc
    from gnn_mtl_b200 import parallel as par
    from gnn_mtl_b200.adjacency import _Csr
    rng = np.random.default_rng(2)
    for n, nt in ((53, 120), (7, 3), (200, 150)):
        h, t = rng.integers(0, n, nt), rng.integers(0, max(n // 2, 1), nt)     # unsymmetric reference pattern
        A = sp.csr_matrix((rng.standard_normal(nt).astype(np.float32), (h, t)), shape=(n, n))
        A.sum_duplicates(); A.sort_indices()
        At = A.T.tocsr(); At.sort_indices()

        class _Full:
            pass
        full = _Full()
        full.n, full.device = n, torch.device("cpu")
        mk = lambda M: _Csr(n, n, torch.from_numpy(M.indptr.astype(np.int32)), torch.from_numpy(M.indices.astype(np.int32)),
                            torch.from_numpy(M.data.astype(np.float32)), threshold=4)
        full.csr, full.csr_t = mk(A), mk(At)
        sh = par.ShardedAdjacency(full, halo=True)
        plain = par.ShardedAdjacency(full)
        H = rng.standard_normal((n, 5)).astype(np.float32)
        Hl = torch.from_numpy(H[sh.r0:sh.r1])
        for transposed, M in ((False, A), (True, At)):
            plan = sh.plan_t if transposed else sh.plan
            buf = sh.gather(Hl, transposed=transposed)
            c = sh.csr_t if transposed else sh.csr
            nl = sh.r1 - sh.r0
            assert buf.shape[0] == max(nl + plan.n_need, 1) and c.n_cols == buf.shape[0]
            pc = plain.csr_t if transposed else plain.csr
            cols = pc.col.numpy()
            need = np.unique(cols[(cols < sh.r0) | (cols >= sh.r1)])
            assert plan.n_need == need.size                                        # only referenced REMOTE rows travel
            assert np.array_equal(buf.numpy()[:nl], H[sh.r0:sh.r1])                 # own rows in place
            assert np.array_equal(buf.numpy()[nl:nl + plan.n_need], H[need])
            loc = sp.csr_matrix((c.val.numpy(), c.col.numpy(), c.rowptr.numpy()), shape=(c.n_rows, c.n_cols))
            assert np.allclose(loc @ buf.numpy(), (M @ H)[sh.r0:sh.r1], atol=1e-5)
        assert 0.0 <= sh.remote_fraction <= 1.0
        auto = par.ShardedAdjacency(full, halo="auto")
        assert auto.halo == (auto.remote_fraction < 0.6)
        assert auto.csr.n_cols == (sh.csr.n_cols if auto.halo else n)


def _w_overlapped_gather(rank, world):
    """Column-chunked gather + column-wise operator == unchunked call, bit for bit; chunk bounds aligned."""
    from gnn_mtl_b200 import parallel as par
    assert par.column_chunks(300, 4) == [(0, 76), (76, 152), (152, 228), (228, 300)]
    assert par.column_chunks(7, 4, align=4) == [(0, 4), (4, 7)]
    assert par.column_chunks(8, 1) == [(0, 8)]
    g = torch.Generator().manual_seed(0)
    n, d = 37, 22
    H = torch.randn(n, d, generator=g)
    M = torch.randn(11, n, generator=g)
    r0, r1 = par.shard_range(n, rank, world)
    calls = []

    def op(full_rows):
        calls.append(full_rows.shape[1])
        return M @ full_rows
    got = par.gather_apply_overlapped(H[r0:r1], n, op, n_chunks=3)
    assert calls == [8, 8, 6]
    want = torch.cat([M @ H[:, c0:c1] for c0, c1 in par.column_chunks(d, 3)], 1)
    assert torch.equal(got, want)
    assert torch.allclose(got, M @ H, atol=1e-5)


def _w_sharded_neg_and_pairs(rank, world):
    """get_neg / generate_pairs with rows split over ranks == the oracle's single-process result, incl. exact ties
    (duplicated embeddings) where the lowest index must win on every rank."""
    from oracle import ea_oracle as orc
    from gnn_mtl_b200 import parallel as par
    rng = np.random.default_rng(11)
    e1, e2, d = 23, 19, 4
    x = (rng.integers(-2, 3, (e1 + e2, d)) * 0.5).astype(np.float32)       # coarse grid: many tied distances
    x[e1 + 3] = x[e1 + 7]
    x[5] = x[9]
    out = torch.from_numpy(x)

    def topk_fn(A, B, skip, k):
        D = torch.from_numpy(orc.l1_matrix(A.numpy(), B.numpy()))
        return torch.argsort(D, dim=1, stable=True)[:, skip:skip + k]

    def argmins_fn(Lr, Rr):
        D = orc.l1_matrix(Lr.numpy(), Rr.numpy())
        return (torch.from_numpy(D.min(1)), torch.from_numpy(D.argmin(1)), torch.from_numpy(D.min(0)),
                torch.from_numpy(D.argmin(0)))
    anchors = np.array([0, 5, 9, 22, 30, 41, 7], dtype=np.int64)
    got = par.get_neg_sharded(anchors, out, 6, topk_fn=topk_fn)
    assert np.array_equal(got, orc.nearest_negatives(anchors, out, 6))
    data = {"e1": e1, "e2": e2, "index1": np.arange(e1), "index2": np.arange(e2) + e1}
    pairs = par.generate_pairs_sharded(out, data, 10, argmins_fn=argmins_fn)
    want = orc.mutual_nearest_pairs(out, data["index1"], data["index2"], 10)
    assert np.array_equal(pairs, want), (pairs, want)
    # merge rule on its own: equal minima on both ranks -> the lower row index wins
    v = torch.tensor([1.0, 2.0 + rank, 0.5], dtype=torch.float64)
    a = torch.tensor([10 + rank, 3 - rank, 7 * (1 - rank) + 2], dtype=torch.int64)
    m, arg = par.merge_col_argmin(v, a)
    assert m.tolist() == [1.0, 2.0, 0.5] and arg.tolist() == [10, 3, 2]


def _w_overlapped_grad_sync(rank, world):
    from gnn_mtl_b200 import parallel as par
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.ReLU(), torch.nn.Linear(4, 3))
    ref = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.ReLU(), torch.nn.Linear(4, 3))
    ref.load_state_dict(net.state_dict())
    xs = [torch.randn(6, 5, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    sync = par.OverlappedGradSync(net.parameters())
    for step in range(2):
        net.zero_grad()
        net(xs[rank]).square().sum().backward()
        sync.finish()
        ref.zero_grad()
        for r in range(world):
            (ref(xs[r]).square().sum() / world).backward()
        for p, q in zip(net.parameters(), ref.parameters()):
            assert torch.allclose(p.grad, q.grad, atol=1e-6), step
    sync.remove()


@pytest.mark.parametrize("worker", [_w_sinkhorn, _w_gather_and_grads, _w_rank_merge, _w_sharded_adjacency,
                                    _w_overlapped_gather, _w_sharded_neg_and_pairs, _w_overlapped_grad_sync,
                                    _w_halo_exchange])
def test_world2_gloo(worker):
    _run(worker, 2)
