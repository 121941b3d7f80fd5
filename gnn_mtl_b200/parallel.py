"""One-node multi-GPU layer: one process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch) for the exchanges the path really has.

  * Sinkhorn / eval: source rows are block-partitioned across ranks, the target
    set is replicated.  Row updates are local.  The column update needs
    LSE over ALL rows: each rank reduces its own rows with the fused kernel and
    the [G, J] partial log-sum-exps are all-gathered and combined (one 4·J-byte
    exchange per sweep).  Eval: diagonal all-gather + one int32 sum of column
    rank counts.
  * SpMM: 1-D row partition of A (and of Aᵀ for the backward); before each aggregation a rank fetches the
    feature rows its block references (needed-rows exchange, one all-to-all with uneven splits — HaloPlan), or,
    as the plain variant, all rows (all-gather).  Only worth it for graphs beyond one GPU.
  * Weights are replicated; `allreduce_grads` averages their gradients.

The reference has no distributed code at all (single process), so
these entry points are additions, not mirrors.  The collective logic is written
against callables so the world_size-2 gloo tests can drive it on CPU tensors.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n, rank, size):
    """Contiguous block partition with equal ceil(n/size) blocks (last may be short/empty)."""
    per = -(-n // size)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def all_gather_rows(local, n_total, group=None):
    """Concatenate equal-size row blocks from every rank ([per, ...] each, padded) -> [n_total, ...]."""
    rank, size = world(group)
    if size == 1:
        return local
    per = -(-n_total // size)
    if local.shape[0] != per:
        pad = torch.zeros((per - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad])
    out = torch.empty((per * size,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:n_total]


def combine_partial_lse(local_lse, group=None):
    """log sum_r exp(lse_r) over ranks: all-gather the [J] partials, reduce the [G, J] stack."""
    rank, size = world(group)
    if size == 1:
        return local_lse
    flat = local_lse.contiguous().reshape(-1)
    stack = torch.empty(size * flat.numel(), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(stack, flat, group=group)
    return torch.logsumexp(stack.view(size, -1), dim=0).view(local_lse.shape)


def allreduce_grads(params, group=None):
    """Average parameter gradients over ranks with ONE flat all-reduce (≈1 MB for the EA model)."""
    rank, size = world(group)
    params = [p for p in params if p.grad is not None]
    if size == 1 or not params:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, group=group)
    flat /= size
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p))
        off += n


class OverlappedGradSync:
    """Gradient averaging that overlaps with the backward pass: every parameter gets a post-accumulate hook that
    starts an asynchronous all-reduce of its gradient the moment autograd has produced it (the last layer's
    gradients travel while the earlier layers are still being differentiated); ``finish()`` waits for the
    outstanding reductions and divides by the world size — call it between ``loss.backward()`` and
    ``optimizer.step()``.  Replaces the flat, synchronous ``allreduce_grads`` after backward (the 1→8 GPU curve of
    round 1 lost its last 2.6 % there).

    Not together with a Sinkhorn solve running on a side stream (models_ea.OVERLAP_SINKHORN): an NCCL kernel that has
    to squeeze in next to a solve holding 120 SMs on one rank stalls its peer on the other rank (measured at N = 2:
    23.7 ms per step against 12.6 ms with ``join_pending_solve()`` + the flat ``allreduce_grads`` after backward)."""

    def __init__(self, params, group=None):
        self.group = group
        self.rank, self.size = world(group)
        self.params = [p for p in params if p.requires_grad]
        self.pending = []
        self.handles = []
        if self.size > 1:
            for p in self.params:
                self.handles.append(p.register_post_accumulate_grad_hook(self._hook))

    def _hook(self, p):
        work = dist.all_reduce(p.grad, group=self.group, async_op=True)
        self.pending.append((work, p))

    def finish(self):
        if self.size == 1:
            return
        for work, p in self.pending:
            work.wait()
            p.grad.div_(self.size)
        self.pending = []

    def remove(self):
        for h in self.handles:
            h.remove()
        self.handles = []


def column_chunks(d, n_chunks, align=4):
    """Split [0, d) into at most n_chunks contiguous ranges whose starts are multiples of `align`."""
    per = -(-d // max(n_chunks, 1))
    per = -(-per // align) * align
    return [(c0, min(d, c0 + per)) for c0 in range(0, d, per)]


def gather_apply_overlapped(local_rows, n_total, apply_fn, n_chunks=4, group=None, align=4):
    """out = apply_fn(all-gathered rows), pipelined over column chunks: while apply_fn works on the gathered
    columns of chunk c (on the current stream) the all-gather of chunk c+1 is already in flight on a side stream.
    apply_fn must act column-wise (out[:, j] depends on input[:, j] only — an SpMM does), so the result is
    bit-identical to the unchunked call.  Returns cat of the per-chunk outputs along dim 1.

    Measured on 2 B200s (profiles/r01_multi_gpu.md) this LOSES for the SpMM: narrow column chunks cost the gather
    kernel more than the overlap saves, so nothing calls it by default; kept as the tested building block for
    operators whose cost does not depend on the row width."""
    rank, size = world(group)
    d = local_rows.shape[1]
    chunks = column_chunks(d, n_chunks, align)
    if size == 1 or len(chunks) <= 1:
        return apply_fn(all_gather_rows(local_rows, n_total, group))
    if not local_rows.is_cuda:                       # gloo / CPU: same schedule, no streams
        return torch.cat([apply_fn(all_gather_rows(local_rows[:, c0:c1].contiguous(), n_total, group))
                          for c0, c1 in chunks], dim=1)
    cur = torch.cuda.current_stream(local_rows.device)
    side = torch.cuda.Stream(device=local_rows.device)
    side.wait_stream(cur)
    gathered, ready = [], []
    with torch.cuda.stream(side):
        for c0, c1 in chunks:
            g = all_gather_rows(local_rows[:, c0:c1].contiguous(), n_total, group)
            ev = torch.cuda.Event()
            ev.record(side)
            gathered.append(g)
            ready.append(ev)
    local_rows.record_stream(side)
    outs = []
    for g, ev in zip(gathered, ready):
        cur.wait_event(ev)
        g.record_stream(cur)
        outs.append(apply_fn(g))
    return torch.cat(outs, dim=1)


# --------------------------------------------------------------------------- Sinkhorn

def sharded_sinkhorn(col_lse_local, row_update_local, n_rows_total, n_cols, log_b, b, n_local, device, dtype,
                     numItermax=1000, stopThr=1e-9, group=None):
    """Row-sharded log-domain Sinkhorn (the iteration of utils/ot_loss.py:50-70).

    col_lse_local(log_u_local) -> [J] log-sum-exp over THIS rank's rows
    row_update_local(log_v)    -> [n_local] new log_u for this rank's rows
    Returns (log_u_local, log_v, sweeps, err).  Same schedule as the single-GPU
    solver: column then row per sweep, marginal error from the next column pass
    on sweeps 0, 10, 20, ….
    """
    log_u = torch.full((n_local,), -math.log(n_rows_total), dtype=dtype, device=device)
    log_v = torch.full((n_cols,), -math.log(n_cols), dtype=dtype, device=device)
    err, sweeps = 1.0, 0
    for cpt in range(numItermax):
        col_lse = combine_partial_lse(col_lse_local(log_u), group)
        if cpt >= 1 and (cpt - 1) % 10 == 0:
            err = float(torch.linalg.vector_norm(torch.exp(log_v + col_lse).double() - b.double()))
            if not err > stopThr:
                break
        log_v = log_b - col_lse
        log_u = row_update_local(log_v)
        sweeps = cpt + 1
    return log_u, log_v, sweeps, err


def sinkhorn_fused_sharded(X_local, Y, a_local, b, reg, n_rows_total, numItermax=50, stopThr=0.0, cost="l2",
                           algo="tcgen05", group=None):
    """Fused (cost never materialised) Sinkhorn with the rows of X sharded over
    ranks and Y replicated.  Returns (log_u_local, log_v, loss, info)."""
    from . import _lib, ops
    cost_id = {"l2": _lib.COST_L2, "sqeuclid": _lib.COST_SQEUCLID, "cos": _lib.COST_COSINE}[cost]
    algo_id = {"simt": _lib.ALGO_SIMT, "tcgen05": _lib.ALGO_TCGEN05}[algo]
    A = ops.FusedOperand(X_local.detach(), cost_id, algo_id)
    B = ops.FusedOperand(Y.detach(), cost_id, algo_id)
    inv_reg = 1.0 / reg
    log_a = torch.log(a_local.detach().float()).contiguous()
    log_b = torch.log(b.detach().float()).contiguous()

    def col_lse_local(log_u):
        return ops.lse_fused(B, A, cost_id, inv_reg, log_u, None, algo_id, want_pot=False, want_lse=True)[1]

    def row_update_local(log_v):
        return ops.lse_fused(A, B, cost_id, inv_reg, log_v, log_a, algo_id)[0]

    log_u, log_v, sweeps, err = sharded_sinkhorn(col_lse_local, row_update_local, n_rows_total, B.n, log_b,
                                                 b.detach().float(), A.n, A.X.device, torch.float32,
                                                 numItermax, stopThr, group)
    _, loss, _ = ops.plan_fused(A, B, cost_id, inv_reg, log_u, log_v, want_plan=False, want_rows=False, algo=algo_id)
    loss = loss.reshape(1).clone()
    if world(group)[1] > 1:
        dist.all_reduce(loss, group=group)
    return log_u, log_v, loss[0], {"sweeps": sweeps, "err": err}


# --------------------------------------------------------------------------- eval

def merge_rank_counts(rank_row_local, rank_col_partial, n_total, group=None):
    """Row ranks are complete per rank (all-gather); column ranks are partial counts (sum)."""
    rank, size = world(group)
    if size == 1:
        return rank_row_local, rank_col_partial
    rows = all_gather_rows(rank_row_local, n_total, group)
    cols = rank_col_partial.clone()
    dist.all_reduce(cols, group=group)
    return rows, cols


def get_hits_sharded(vec, test_pair, top_k=(1, 10, 50, 100), group=None):
    """utils/eval_utils.py:71-98 with the rows of the L1 matrix split over ranks.
    Every rank passes the same `vec` / `test_pair` and gets the same dict."""
    from . import ops
    from .utils.eval_utils import _hits_dict, _pair_index, _to_cuda
    rank, size = world(group)
    vec = _to_cuda(vec.detach())
    left, right = _pair_index(test_pair, vec.device)
    n = int(left.numel())
    r0, r1 = shard_range(n, rank, size)
    R = vec.index_select(0, right).float()
    L_loc = vec.index_select(0, left[r0:r1]).float()
    diag_loc = ops.l1_paired(L_loc, R[r0:r1])
    diag = all_gather_rows(diag_loc, n, group)
    rank_row = torch.zeros(n, dtype=torch.int32, device=vec.device)
    rank_col = torch.zeros(n, dtype=torch.int32, device=vec.device)
    if r1 > r0:
        ops.l1_rank_fused(L_loc, r0, R, diag, rank_row, rank_col)     # streamed: no rows x n block is stored
    rank_row, rank_col = merge_rank_counts(rank_row[r0:r1], rank_col, n, group)
    return _hits_dict(rank_row, rank_col, top_k, n)


def merge_col_argmin(col_min_part, col_arg_part, group=None):
    """Column arg-min over row blocks held by different ranks, exact (value, lowest row index) semantics
    (models/models_ea.py:151-153 takes np.argmin of the full matrix): all-reduce MIN of the fp64 values, then
    all-reduce MIN of the candidate row indices (a rank whose partial minimum is not the global one proposes
    +inf).  Two 8·E2-byte reductions; an fp64 value and an index do not fit one 64-bit word, so the packed
    single-reduce form of SURVEY §8e is split in two."""
    rank, size = world(group)
    if size == 1:
        return col_min_part, col_arg_part
    best = col_min_part.clone()
    dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
    big = torch.iinfo(torch.int64).max
    cand = torch.where((col_min_part == best) & (col_arg_part >= 0), col_arg_part, torch.full_like(col_arg_part, big))
    dist.all_reduce(cand, op=dist.ReduceOp.MIN, group=group)
    return best, cand


def get_neg_sharded(ILL, output, k, group=None, topk_fn=None):
    """BaseModel.get_neg (models/models_ea.py:19-30) with the anchors split over the ranks: every rank ranks its
    block of anchors against ALL entities (no exchange), the [t, k] index blocks are all-gathered.  Every rank passes
    the same arguments and gets the same flat int64 array.  ``topk_fn(anchor_rows, all_rows, skip, k)`` defaults to
    the L1 top-k kernels."""
    import numpy as np
    rank, size = world(group)
    if topk_fn is None:
        from . import ops
        topk_fn = ops.l1_topk
    out = output.detach().to(torch.float32)
    anchors = torch.as_tensor(np.asarray(ILL, dtype=np.int64), device=out.device)
    t = int(anchors.numel())
    r0, r1 = shard_range(t, rank, size)
    local = topk_fn(out.index_select(0, anchors[r0:r1]), out, 1, k)
    full = all_gather_rows(local, t, group)
    return full.reshape(-1).cpu().numpy()


def generate_pairs_sharded(outputs, data, bsz, group=None, argmins_fn=None):
    """UEAModel.generate_pairs (models/models_ea.py:143-167) with the rows of the E1 x E2 L1 matrix split over the
    ranks: row arg-mins are local, column arg-mins are merged with ``merge_col_argmin``.  Returns the [<= bsz, 2]
    array of mutual nearest neighbours, closest first (local positions, like the reference), identical on every
    rank.  ``argmins_fn(L_rows, R_rows) -> (row_min, row_arg, col_min, col_arg)`` defaults to the L1 kernels."""
    rank, size = world(group)
    if argmins_fn is None:
        from . import ops
        argmins_fn = ops.l1_argmins
    e1, e2 = data['e1'], data['e2']
    index1, index2 = data['index1'], data['index2']
    out = outputs.detach().to(torch.float32)
    L = torch.as_tensor([index1[i] for i in range(e1)], device=out.device)
    R = torch.as_tensor([index2[i] for i in range(e2)], device=out.device)
    r0, r1 = shard_range(e1, rank, size)
    Rrows = out.index_select(0, R)
    if r1 > r0:
        row_min_l, row_arg_l, col_min_p, col_arg_p = argmins_fn(out.index_select(0, L[r0:r1]), Rrows)
        col_arg_p = torch.where(col_arg_p >= 0, col_arg_p + r0, col_arg_p)      # block-local row -> global row
    else:
        row_min_l = torch.empty(0, dtype=torch.float64, device=out.device)
        row_arg_l = torch.empty(0, dtype=torch.int64, device=out.device)
        col_min_p = torch.full((e2,), float("inf"), dtype=torch.float64, device=out.device)
        col_arg_p = torch.full((e2,), -1, dtype=torch.int64, device=out.device)
    row_min = all_gather_rows(row_min_l, e1, group)
    row_arg = all_gather_rows(row_arg_l, e1, group)
    _, col_arg = merge_col_argmin(col_min_p, col_arg_p, group)
    mutual = col_arg[row_arg] == torch.arange(e1, device=out.device)
    keep = torch.nonzero(mutual).reshape(-1)
    pairs = torch.stack([keep, row_arg[keep]], 1)
    order = torch.argsort(row_min[keep], stable=True)[:bsz]
    return pairs[order].cpu().numpy()


# --------------------------------------------------------------------------- SpMM

class HaloPlan:
    """Needed-rows ("halo") exchange plan of one row block of a CSR matrix.

    The all-gather moves every feature row to every rank; a rank only reads the rows its own block of A references
    (on the 10M-node power-law graph of BASELINE.json config 4 at 8 ranks: under half of the remote ones).  The plan is
    built once per adjacency: the sorted distinct REMOTE column ids of the block, grouped by owner rank (block
    partition), are sent to their owners (all-to-all of counts, then of indices); every aggregation then is
        pack = H_local[send_idx]  ->  all_to_all_single with uneven splits into the tail of the operand buffer
        [ H_local | rows received from rank 0, 1, ... ]  ->  SpMM on the block with its columns REMAPPED into that buffer
    (own columns: col - r0; remote columns: n_local + position in the sorted need list).  Own rows never travel and are
    never packed.  The remap is monotone within a row only per segment (own / remote), so the summation order inside a
    row can differ from the all-gather route's: results agree to fp32 rounding, not bit for bit, unless the block has
    no own columns."""

    def __init__(self, csr_block, n_total, group=None):
        from .adjacency import _Csr
        self.group = group
        rank, size = world(group)
        dev = csr_block.col.device
        per = -(-n_total // size)
        r0, r1 = shard_range(n_total, rank, size)
        self.n_local = r1 - r0
        cols = csr_block.col.to(torch.int64)
        remote = (cols < r0) | (cols >= r1)
        need = torch.unique(cols[remote])                           # sorted -> grouped by owner rank, in rank order
        owner = torch.div(need, per, rounding_mode="floor")
        counts_in = torch.bincount(owner, minlength=size).to(torch.int64)       # rows I receive from every rank
        counts_out = torch.empty_like(counts_in)
        if size > 1:
            dist.all_to_all_single(counts_out, counts_in, group=group)
        else:
            counts_out.copy_(counts_in)
        self.in_splits = [int(v) for v in counts_in.tolist()]       # receive sizes (rows)
        self.out_splits = [int(v) for v in counts_out.tolist()]     # send sizes (rows)
        wanted = torch.empty(sum(self.out_splits), dtype=torch.int64, device=dev)
        if size > 1:
            dist.all_to_all_single(wanted, need, output_split_sizes=self.out_splits, input_split_sizes=self.in_splits,
                                   group=group)
        self.send_idx = (wanted - r0).contiguous()                  # my local rows, in the order the peers expect them
        self.n_need = int(need.numel())                             # remote rows this rank reads
        self.n_remote_total = n_total - self.n_local
        new_col = torch.where(remote, self.n_local + torch.searchsorted(need, cols), cols - r0)
        # the kernels read the (col, val) pairs of a row in storage order: keep every row sorted by the new column id
        # is not required for correctness, only the order of summation changes
        self.csr = _Csr(csr_block.n_rows, max(self.n_local + self.n_need, 1), csr_block.rowptr,
                        new_col.to(torch.int32).contiguous(), csr_block.val, csr_block.threshold)

    def exchange(self, local_rows):
        """[rows of this rank, d] -> [n_local + n_need, d]: own rows, then the remote rows this block references."""
        rank, size = world(self.group)
        out = torch.empty((max(self.n_local + self.n_need, 1),) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype,
                          device=local_rows.device)
        if self.n_local + self.n_need == 0:
            out.zero_()
        out[:self.n_local].copy_(local_rows)
        if size > 1:
            pack = local_rows.index_select(0, self.send_idx)
            dist.all_to_all_single(out[self.n_local:self.n_local + self.n_need], pack,
                                   output_split_sizes=self.in_splits, input_split_sizes=self.out_splits, group=self.group)
        return out

    def remote_fraction(self):
        """Largest fraction, over ranks, of the remote rows a rank has to fetch (1.0 = as many as an all-gather)."""
        rank, size = world(self.group)
        f = torch.tensor([self.n_need / max(self.n_remote_total, 1)], dtype=torch.float64, device=self.csr.col.device)
        if size > 1:
            dist.all_reduce(f, op=dist.ReduceOp.MAX, group=self.group)
        return float(f[0])


class ShardedAdjacency:
    """Row block [r0, r1) of A and of Aᵀ (kernel-format CSR).
    ``halo=False``: full column range; `gather` all-gathers every feature row before an aggregation.
    ``halo=True``: columns remapped into [own rows | fetched rows]; `gather` exchanges only the remote rows the block
    references (HaloPlan) — a fraction of the bytes over NVLink where the graph is sparse enough.
    ``halo="auto"``: the plan is built, and used iff every rank fetches under 60 % of its remote rows (at 2 ranks the
    power-law benchmark graph needs 70 % of them and the all-gather wins; at 8 ranks under half)."""

    sharded = True

    def __init__(self, full, group=None, halo=False):
        from .adjacency import _Csr
        self.group = group
        self.rank, self.size = world(group)
        self.n = full.n
        self.r0, self.r1 = shard_range(full.n, self.rank, self.size)
        self.csr = self._slice(full.csr, _Csr)
        self.csr_t = self._slice(full.csr_t, _Csr)
        self.device = full.device
        self.halo = bool(halo)
        self.remote_fraction = None
        if self.halo:
            self.plan = HaloPlan(self.csr, self.n, group)
            self.plan_t = HaloPlan(self.csr_t, self.n, group)
            self.remote_fraction = max(self.plan.remote_fraction(), self.plan_t.remote_fraction())
            if halo == "auto" and self.remote_fraction >= 0.6:
                self.halo = False
                del self.plan, self.plan_t
            else:
                self.csr, self.csr_t = self.plan.csr, self.plan_t.csr

    def _slice(self, c, _Csr):
        e0, e1 = int(c.rowptr[self.r0]), int(c.rowptr[self.r1])
        rowptr = (c.rowptr[self.r0:self.r1 + 1] - e0).contiguous()
        return _Csr(self.r1 - self.r0, c.n_cols, rowptr, c.col[e0:e1].clone(), c.val[e0:e1].clone(), c.threshold)

    def gather(self, local_rows, transposed=False):
        """The operand the SpMM on `csr` (or, transposed, `csr_t`) reads: all rows, or this rank's halo."""
        if self.halo:
            return (self.plan_t if transposed else self.plan).exchange(local_rows)
        return all_gather_rows(local_rows, self.n, self.group)

    def local(self, full_rows):
        return full_rows[self.r0:self.r1]

    def aggregate_overlapped(self, local_rows, transposed=False, n_chunks=4):
        """(A or Aᵀ)[my rows] · H with the feature all-gather pipelined against the SpMM over column chunks."""
        from . import ops
        if self.halo:
            raise ValueError("aggregate_overlapped is the all-gather route; build the adjacency with halo=False")
        csr = self.csr_t if transposed else self.csr
        return gather_apply_overlapped(local_rows, self.n, lambda g: ops.spmm(csr, g)[0], n_chunks, self.group)
