"""Thin tensor-level wrappers over the C ABI (allocation + argument marshalling).

Everything here runs on CUDA tensors; nothing falls back to PyTorch math.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import lib, ptr, stream, check


def _f32c(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.to(torch.float32).contiguous()


# ---- (a) SpMM ----------------------------------------------------------------

# bench.py sets this to a list to collect (start_event, end_event, csr, d, fused, saved) per launch
SPMM_TIMER = None

def spmm(csr, H, act=_lib.ACT_IDENTITY, gate_pre=None, x_res=None, save_act=False):
    """out = epilogue(csr · H); returns (out, act_out or None)."""
    _lib.require_cuda(H, gate_pre, x_res)
    H = _f32c(H)
    n_rows, d = csr.n_rows, H.shape[1]
    if H.shape[0] != csr.n_cols:
        raise ValueError("spmm: H has %d rows, adjacency has %d columns" % (H.shape[0], csr.n_cols))
    if gate_pre is not None:
        gate_pre, x_res = _f32c(gate_pre), _f32c(x_res)
    out = torch.empty(n_rows, d, dtype=torch.float32, device=H.device)
    act_out = torch.empty_like(out) if save_act else None
    with torch.cuda.device(H.device):
        scratch = csr.scratch(d)
        if SPMM_TIMER is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        check(lib.eg_spmm(ptr(csr.rowptr), ptr(csr.col), ptr(csr.val), n_rows, ptr(H), d, act,
                          ptr(gate_pre), ptr(x_res), ptr(out), ptr(act_out), csr.threshold,
                          ptr(csr.seg_row), ptr(csr.seg_begin), ptr(csr.seg_end), csr.n_seg,
                          ptr(csr.long_rows), ptr(csr.long_first), csr.n_long, ptr(scratch), stream()),
              "eg_spmm")
        if SPMM_TIMER is not None:
            ev1.record()
            SPMM_TIMER.append((ev0, ev1, csr, d, gate_pre is not None, save_act))
    return out, act_out


def epilogue_bwd(dout, a, gate_pre, x_res, act, need_gate, need_xres):
    dout = _f32c(dout)
    dS = torch.empty_like(dout)
    d_gate = torch.empty_like(dout) if (gate_pre is not None and need_gate) else None
    d_xres = torch.empty_like(dout) if (gate_pre is not None and need_xres) else None
    with torch.cuda.device(dout.device):
        check(lib.eg_epilogue_bwd(ptr(dout), ptr(a), ptr(gate_pre), ptr(x_res), dout.numel(), act, ptr(dS),
                                  ptr(d_gate), ptr(d_xres), stream()), "eg_epilogue_bwd")
    return dS, d_gate, d_xres


# ---- (c) L1 evaluation -----------------------------------------------------------

L1_BLOCK_BYTES = 2 << 30   # fp64 distance rows materialised per launch


def l1_block_rows(n_cols, budget=None):
    rows = max(64, (budget or L1_BLOCK_BYTES) // (8 * max(int(n_cols), 1)))
    return int(min(rows, 65535 * 32))


def l1_matrix(L, R, out=None):
    _lib.require_cuda(L, R)
    L, R = _f32c(L), _f32c(R)
    nL, d = L.shape
    nR = R.shape[0]
    if out is None:
        out = torch.empty(nL, nR, dtype=torch.float64, device=L.device)
    with torch.cuda.device(L.device):
        check(lib.eg_l1_matrix(ptr(L), nL, ptr(R), nR, d, ptr(out), out.stride(0) if nL else nR, stream()),
              "eg_l1_matrix")
    return out


def l1_paired(L, R):
    L, R = _f32c(L), _f32c(R)
    out = torch.empty(L.shape[0], dtype=torch.float64, device=L.device)
    with torch.cuda.device(L.device):
        check(lib.eg_l1_paired(ptr(L), ptr(R), L.shape[0], L.shape[1], ptr(out), stream()), "eg_l1_paired")
    return out


def rank_accumulate(D, row0, diag, rank_row, rank_col):
    with torch.cuda.device(D.device):
        check(lib.eg_rank_accumulate(ptr(D), D.stride(0), row0, D.shape[0], D.shape[1], ptr(diag), ptr(rank_row),
                                     ptr(rank_col), stream()), "eg_rank_accumulate")


def argmin_accumulate(D, row0, row_min, row_arg, col_min, col_arg):
    with torch.cuda.device(D.device):
        check(lib.eg_argmin_accumulate(ptr(D), D.stride(0), row0, D.shape[0], D.shape[1], ptr(row_min), ptr(row_arg),
                                       ptr(col_min), ptr(col_arg), stream()), "eg_argmin_accumulate")


def topk_rows(D, skip, k):
    out = torch.empty(D.shape[0], k, dtype=torch.int64, device=D.device)
    with torch.cuda.device(D.device):
        check(lib.eg_topk_rows(ptr(D), D.stride(0), D.shape[0], D.shape[1], skip, k, ptr(out), stream()),
              "eg_topk_rows")
    return out


FUSED_ROW_CHUNK = 65535 * 64      # rows per streamed launch (grid.y limit x 64-row strips)


L1_RANK_FILTER = True              # fp32 candidate filter + exact fp64 decisions (identical ranks); False: all fp64
FILTER_PAIRS_PER_CALL = 1 << 35   # bounds the candidate queue (pairs / 256 entries of 8 bytes) to ~1 GiB


def l1_rank_fused(L_block, row0, R, diag, rank_row, rank_col, filtered=None):
    """Streamed rank counts for rows [row0, row0 + len(L_block)) of the L1 matrix against all of R: nothing of
    size rows x cols is stored.  rank_row[row0:...] is overwritten, rank_col accumulated."""
    _lib.require_cuda(L_block, R, diag, rank_row, rank_col)
    L_block, R = _f32c(L_block), _f32c(R)
    filtered = L1_RANK_FILTER if filtered is None else filtered
    nR = R.shape[0]
    with torch.cuda.device(R.device):
        if filtered and nR < (1 << 31):
            chunk = max(128, min(65535 * 128, FILTER_PAIRS_PER_CALL // max(nR, 1)) // 128 * 128)
            ws = None
            for c0 in range(0, L_block.shape[0], chunk):
                blk = L_block[c0:c0 + chunk]
                nb = int(lib.eg_l1_rank_filtered_workspace_bytes(blk.shape[0], nR))
                if ws is None or ws.numel() < nb:
                    ws = torch.empty(nb, dtype=torch.uint8, device=R.device)
                check(lib.eg_l1_rank_filtered(ptr(blk), blk.shape[0], row0 + c0, ptr(R), nR, R.shape[1], ptr(diag),
                                              ptr(rank_row), ptr(rank_col), ptr(ws), ws.numel(), stream()),
                      "eg_l1_rank_filtered")
            return
        for c0 in range(0, L_block.shape[0], FUSED_ROW_CHUNK):
            blk = L_block[c0:c0 + FUSED_ROW_CHUNK]
            check(lib.eg_l1_rank_fused(ptr(blk), blk.shape[0], row0 + c0, ptr(R), R.shape[0], R.shape[1], ptr(diag),
                                       ptr(rank_row), ptr(rank_col), stream()), "eg_l1_rank_fused")


def l1_ranks(L, R, block_bytes=None, streamed=True, filtered=None):
    """Ranks of the diagonal (true match) per row and per column of the fp64 L1
    matrix between L and R (same length).  ``streamed=False`` takes the two-kernel route through a stored
    distance block (kept for cross-checking; bit-identical); ``filtered`` picks the fp32-filter kernel (default)
    or the all-fp64 streamed kernel — identical ranks."""
    n = L.shape[0]
    dev = L.device
    diag = l1_paired(L, R)
    rank_row = torch.zeros(n, dtype=torch.int32, device=dev)
    rank_col = torch.zeros(n, dtype=torch.int32, device=dev)
    if streamed:
        l1_rank_fused(L, 0, R, diag, rank_row, rank_col, filtered=filtered)
        return rank_row, rank_col
    rows = l1_block_rows(n, block_bytes)
    buf = torch.empty(min(rows, n), n, dtype=torch.float64, device=dev)
    for r0 in range(0, n, rows):
        r1 = min(n, r0 + rows)
        D = l1_matrix(L[r0:r1], R, out=buf[: r1 - r0])
        rank_accumulate(D, r0, diag, rank_row, rank_col)
    return rank_row, rank_col


def matrix_ranks(S):
    """Same ranking on a given fp64 score matrix (square)."""
    S = S.to(torch.float64).contiguous()
    n = S.shape[0]
    diag = torch.diagonal(S).contiguous()
    rank_row = torch.zeros(n, dtype=torch.int32, device=S.device)
    rank_col = torch.zeros(n, dtype=torch.int32, device=S.device)
    step = 65535 * 32
    for r0 in range(0, n, step):
        rank_accumulate(S[r0:r0 + step], r0, diag, rank_row, rank_col)
    return rank_row, rank_col


def l1_argmins(L, R, block_bytes=None):
    """(row_min, row_arg, col_min, col_arg) of the fp64 L1 matrix, lowest index on ties."""
    nL, nR = L.shape[0], R.shape[0]
    dev = L.device
    row_min = torch.empty(nL, dtype=torch.float64, device=dev)
    row_arg = torch.empty(nL, dtype=torch.int64, device=dev)
    col_min = torch.full((nR,), float("inf"), dtype=torch.float64, device=dev)
    col_arg = torch.full((nR,), -1, dtype=torch.int64, device=dev)
    rows = l1_block_rows(nR, block_bytes)
    buf = torch.empty(min(rows, max(nL, 1)), nR, dtype=torch.float64, device=dev)
    for r0 in range(0, nL, rows):
        r1 = min(nL, r0 + rows)
        D = l1_matrix(L[r0:r1], R, out=buf[: r1 - r0])
        argmin_accumulate(D, r0, row_min, row_arg, col_min, col_arg)
    return row_min, row_arg, col_min, col_arg


# The streamed kernel serves skip + k <= 128, but its per-row lists cost shared memory (one CTA per SM beyond ~48
# entries); measured on B200 it wins for short lists (top-10 over 30,000 x 200,000: 301 vs 383 ms) and loses for the
# reference's neg_num = 125 (424 vs 385 ms), so longer lists keep the stored-block radix select.
TOPK_FUSED_MAX = 32


def l1_topk_fused(L, R, skip, k):
    """Streamed per-row top-k by (distance, index): entries [skip, skip + k) of every row's stable argsort
    (skip + k <= 128)."""
    _lib.require_cuda(L, R)
    L, R = _f32c(L), _f32c(R)
    nL, nR = L.shape[0], R.shape[0]
    out = torch.empty(nL, k, dtype=torch.int64, device=L.device)
    with torch.cuda.device(L.device):
        for c0 in range(0, nL, FUSED_ROW_CHUNK):
            blk = L[c0:c0 + FUSED_ROW_CHUNK]
            nb = int(lib.eg_l1_topk_fused_workspace_bytes(blk.shape[0], nR, skip, k))
            ws = torch.empty(max(nb, 1), dtype=torch.uint8, device=L.device)
            check(lib.eg_l1_topk_fused(ptr(blk), blk.shape[0], ptr(R), nR, L.shape[1], skip, k, ptr(ws), nb,
                                       ptr(out[c0:c0 + FUSED_ROW_CHUNK]), stream()), "eg_l1_topk_fused")
    return out


def l1_topk(L, R, skip, k, block_bytes=None, streamed=True):
    nL, nR = L.shape[0], R.shape[0]
    if streamed is True and skip + k <= TOPK_FUSED_MAX and nL > 0 and nR > 0:
        return l1_topk_fused(L, R, skip, k)
    if streamed == "force" and nL > 0 and nR > 0:
        return l1_topk_fused(L, R, skip, k)
    rows = l1_block_rows(nR, block_bytes)
    buf = torch.empty(min(rows, max(nL, 1)), nR, dtype=torch.float64, device=L.device)
    outs = []
    for r0 in range(0, nL, rows):
        r1 = min(nL, r0 + rows)
        D = l1_matrix(L[r0:r1], R, out=buf[: r1 - r0])
        outs.append(topk_rows(D, skip, k))
    return torch.cat(outs) if outs else torch.empty(0, k, dtype=torch.int64, device=L.device)


# ---- (b) Sinkhorn ------------------------------------------------------------------

def _dt(t):
    if t.dtype == torch.float32:
        return _lib.DT_F32
    if t.dtype == torch.float64:
        return _lib.DT_F64
    raise TypeError("Sinkhorn kernels take fp32 or fp64, got %s" % t.dtype)


def lse_dense(M, inv_reg, pot_in, logw=None, want_pot=True, want_lse=False):
    """Row pass over a materialised cost: lse_i = LSE_j(pot_in_j - M_ij*inv_reg)."""
    dt = _dt(M)
    n_rows, n_cols = M.shape
    pot_out = torch.empty(n_rows, dtype=M.dtype, device=M.device) if want_pot else None
    lse_out = torch.empty(n_rows, dtype=M.dtype, device=M.device) if want_lse else None
    with torch.cuda.device(M.device):
        check(lib.eg_lse_dense(dt, ptr(M), n_rows, n_cols, M.stride(0), float(inv_reg), ptr(pot_in), ptr(logw),
                               ptr(pot_out), ptr(lse_out), stream()), "eg_lse_dense")
    return pot_out, lse_out


def transpose(M):
    out = torch.empty(M.shape[1], M.shape[0], dtype=M.dtype, device=M.device)
    with torch.cuda.device(M.device):
        check(lib.eg_transpose(_dt(M), ptr(M), M.shape[0], M.shape[1], M.stride(0), ptr(out), out.stride(0),
                               stream()), "eg_transpose")
    return out


def plan_dense(M, inv_reg, f, g, want_plan=True, want_rows=False, want_cols=False):
    dt = _dt(M)
    I, J = M.shape
    dev = M.device
    P = torch.empty(I, J, dtype=M.dtype, device=dev) if want_plan else None
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    rs = torch.empty(I, dtype=M.dtype, device=dev) if want_rows else None
    cs = torch.empty(J, dtype=M.dtype, device=dev) if want_cols else None
    with torch.cuda.device(dev):
        check(lib.eg_plan_dense(dt, ptr(M), I, J, M.stride(0), float(inv_reg), ptr(f), ptr(g), ptr(P),
                                J, ptr(loss), ptr(rs), ptr(cs), stream()), "eg_plan_dense")
    return P, loss[0], rs, cs


# bench.py sets this to a list to collect (start_event, end_event, I, J, sweeps, itemsize) per solve
SINKHORN_TIMER = None


def sinkhorn_dense(M, a, b, reg, max_iter, stop_thr):
    """Device solver of utils/ot_loss.py:26-76; returns (log_u, log_v, sweeps, err)."""
    dt = _dt(M)
    I, J = M.shape
    dev = M.device
    log_u = torch.empty(I, dtype=M.dtype, device=dev)
    log_v = torch.empty(J, dtype=M.dtype, device=dev)
    with torch.cuda.device(dev):
        ws_bytes = int(lib.eg_sinkhorn_dense_workspace_bytes(dt, I, J))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        sweeps, err = C.c_int(0), C.c_double(0.0)
        if SINKHORN_TIMER is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        rc = lib.eg_sinkhorn_dense(dt, ptr(M), I, J, float(reg), ptr(a), ptr(b), int(max_iter), float(stop_thr),
                                   None, ptr(log_u), ptr(log_v), ptr(ws), ws_bytes, C.byref(sweeps),
                                   C.byref(err), stream())
        if rc == -3:    # EG_ERR_WORKSPACE: this shape takes the streaming path, which wants the transposed copy
            Mt = torch.empty(J, I, dtype=M.dtype, device=dev)
            rc = lib.eg_sinkhorn_dense(dt, ptr(M), I, J, float(reg), ptr(a), ptr(b), int(max_iter), float(stop_thr),
                                       ptr(Mt), ptr(log_u), ptr(log_v), ptr(ws), ws_bytes, C.byref(sweeps),
                                       C.byref(err), stream())
        check(rc, "eg_sinkhorn_dense")
        if SINKHORN_TIMER is not None:
            ev1.record()
            SINKHORN_TIMER.append((ev0, ev1, I, J, int(sweeps.value), M.element_size()))
    return log_u, log_v, int(sweeps.value), float(err.value)


def row_norms(A, squared=True):
    A = _f32c(A)
    out = torch.empty(A.shape[0], dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        check(lib.eg_row_norms(ptr(A), A.shape[0], A.shape[1], 1 if squared else 0, ptr(out), stream()),
              "eg_row_norms")
    return out


def split_tf32(X, d_pad=None):
    X = _f32c(X)
    n, d = X.shape
    d_pad = d_pad or (d + 7) // 8 * 8
    hi = torch.empty(n, d_pad, dtype=torch.float32, device=X.device)
    lo = torch.empty(n, d_pad, dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        check(lib.eg_split_tf32(ptr(X), n, d, d_pad, ptr(hi), ptr(lo), stream()), "eg_split_tf32")
    return hi, lo


class FusedOperand:
    """One embedding set prepared for the fused passes: fp32 rows, norms for the
    chosen cost, and (tcgen05 path) the 3xTF32 hi/lo split."""

    def __init__(self, X, cost, algo):
        self.X = _f32c(X)
        self.n, self.d = self.X.shape
        self.norm = row_norms(self.X, squared=(cost != _lib.COST_COSINE))
        self.hi = self.lo = None
        if algo == _lib.ALGO_TCGEN05:
            self.hi, self.lo = split_tf32(self.X)


def lse_fused(A, B, cost, inv_reg, pot_in, logw=None, algo=_lib.ALGO_SIMT, want_pot=True, want_lse=False):
    """lse_i = LSE_j(pot_in_j - cost(A_i,B_j)*inv_reg) with A, B FusedOperand."""
    dev = A.X.device
    pot_out = torch.empty(A.n, dtype=torch.float32, device=dev) if want_pot else None
    lse_out = torch.empty(A.n, dtype=torch.float32, device=dev) if want_lse else None
    with torch.cuda.device(dev):
        ws_bytes = int(lib.eg_lse_fused_workspace_bytes(algo, A.n, B.n, A.d))
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        check(lib.eg_lse_fused(algo, cost, ptr(A.X), A.n, ptr(B.X), B.n, A.d, ptr(A.norm), ptr(B.norm),
                               float(inv_reg), ptr(pot_in), ptr(logw), ptr(pot_out), ptr(lse_out),
                               ptr(A.hi), ptr(A.lo), ptr(B.hi), ptr(B.lo), ptr(ws), ws.numel(), stream()),
              "eg_lse_fused")
    return pot_out, lse_out


def plan_fused(A, B, cost, inv_reg, f, g, want_plan=False, want_rows=True, algo=_lib.ALGO_SIMT):
    dev = A.X.device
    if want_plan:
        algo = _lib.ALGO_SIMT      # only the SIMT tiles write P (small problems)
    P = torch.empty(A.n, B.n, dtype=torch.float32, device=dev) if want_plan else None
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    rs = torch.empty(A.n, dtype=torch.float32, device=dev) if want_rows else None
    with torch.cuda.device(dev):
        check(lib.eg_plan_fused(algo, cost, ptr(A.X), A.n, ptr(B.X), B.n, A.d, ptr(A.norm), ptr(B.norm),
                                float(inv_reg), ptr(f), ptr(g), ptr(P), B.n, ptr(loss), ptr(rs),
                                ptr(A.hi), ptr(A.lo), ptr(B.hi), ptr(B.lo), stream()), "eg_plan_fused")
    return P, loss[0], rs


def plan_grad_fused(A, B, cost, inv_reg, f, g, scale=1.0):
    """dA = scale * sum_j P_ij d cost(A_i, B_j)/dA_i with P = exp(f_i + g_j - cost/reg); A, B FusedOperand."""
    dev = A.X.device
    dA = torch.empty_like(A.X)
    with torch.cuda.device(dev):
        nb = int(lib.eg_plan_grad_fused_workspace_bytes(A.n, B.n, A.d))
        ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=dev)
        check(lib.eg_plan_grad_fused(cost, ptr(A.X), A.n, ptr(B.X), B.n, A.d, ptr(A.norm), ptr(B.norm), float(inv_reg),
                                     ptr(f), ptr(g), float(scale), ptr(ws), ws.numel(), ptr(dA), stream()),
              "eg_plan_grad_fused")
    return dA


# ---- dense layer products on the tcgen05 3xTF32 tiles -----------------------------------------

def _pad16(k):
    return (k + 15) // 16 * 16


def gemm_tn(a_split, m, b_split, n):
    """C[m, n] = sum_k A[k, m] B[k, n] with fp32 accuracy (3xTF32 on tcgen05, split-K, deterministic).
    a_split / b_split: (hi, lo) pairs from split_tf32 of the row-major [K, m] / [K, n] operands."""
    (a_hi, a_lo), (b_hi, b_lo) = a_split, b_split
    K = a_hi.shape[0]
    if b_hi.shape[0] != K:
        raise ValueError("gemm_tn: operands disagree on the reduction length")
    dev = a_hi.device
    out = torch.empty(m, n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nb = int(lib.eg_gemm_tn_3xtf32_workspace_bytes(K, m, n))
        ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
        check(lib.eg_gemm_tn_3xtf32(ptr(a_hi), ptr(a_lo), m, a_hi.shape[1], ptr(b_hi), ptr(b_lo), n, b_hi.shape[1], K,
                                    ptr(ws), nb, ptr(out), n, stream()), "eg_gemm_tn_3xtf32")
    return out


def gemm_nt_raw(A_parts, B, bias=None, n1=None, addend=None):
    """C = [A1 | A2] · Bᵀ + bias (+ addend) with RAW fp32 A operands: their hi/lo split happens inside the kernel
    (eg_gemm_nt_3xtf32_raw) — no eg_split_tf32 pass over A, a deeper TMA ring.  Bit-identical to ``gemm_nt``.
    A_parts: one or two [m, k_i] fp32 CUDA tensors with k_i % 4 == 0; B: [n, sum k_i]; addend: [m, n] (n1 == n)."""
    A_parts = [_f32c(a) for a in A_parts]
    if not 1 <= len(A_parts) <= 2:
        raise ValueError("gemm_nt_raw takes one or two A operands")
    m = A_parts[0].shape[0]
    ks = [a.shape[1] for a in A_parts]
    B = _f32c(B)
    n = B.shape[0]
    if B.shape[1] != sum(ks):
        raise ValueError("gemm_nt_raw: B has %d columns, A parts have %s" % (B.shape[1], ks))
    n1 = n if n1 is None else n1
    dev = B.device
    if len(ks) == 1 and ks[0] % 16 == 0:
        Bp = B
    else:
        Bp = torch.zeros(n, sum(_pad16(k) for k in ks), dtype=torch.float32, device=dev)
        src = dst = 0
        for k in ks:
            Bp[:, dst:dst + k] = B[:, src:src + k]
            src += k
            dst += _pad16(k)
    b_hi, b_lo = split_tf32(Bp, Bp.shape[1])
    out1 = torch.empty(m, n1, dtype=torch.float32, device=dev)
    out2 = torch.empty(m, n - n1, dtype=torch.float32, device=dev) if n1 < n else None
    bias = _f32c(bias) if bias is not None else None
    addend = _f32c(addend) if addend is not None else None
    a2 = A_parts[1] if len(A_parts) == 2 else None
    with torch.cuda.device(dev):
        check(lib.eg_gemm_nt_3xtf32_raw(ptr(A_parts[0]), ks[0], ks[0], ptr(a2), ks[1] if a2 is not None else 0,
                                        ks[1] if a2 is not None else 0, m, ptr(b_hi), ptr(b_lo), Bp.shape[1], n, ptr(bias),
                                        ptr(addend), n if addend is not None else 0,
                                        ptr(out1), n1, n1, ptr(out2), (n - n1) if out2 is not None else 0, stream()),
              "eg_gemm_nt_3xtf32_raw")
    return (out1, out2) if out2 is not None else out1


def gemm_nt(A_parts, B, bias=None, n1=None, a_splits=None, return_splits=False, chained=False):
    """C = [A1 | A2 ...] · Bᵀ + bias with fp32 accuracy (3xTF32 on tcgen05).
    ``chained=True``: short accumulation chains folded in fp32 registers (eg_gemm_nt_3xtf32_chained): ~3e-7 relative
    instead of ~2e-6 — for products whose result decides a ReLU branch.
    A_parts: one or two [m, k_i] fp32 CUDA tensors; B: [n, sum k_i] fp32 (row j = output column j).
    Returns out1 [m, n1] (and out2 [m, n - n1] when n1 < n).  ``a_splits``: reuse hi/lo pairs computed earlier
    (one per A part, padded to 16 columns); ``return_splits=True`` appends the list of pairs used."""
    A_parts = [_f32c(a) for a in A_parts]
    if not 1 <= len(A_parts) <= 2:
        raise ValueError("gemm_nt takes one or two A operands")
    m = A_parts[0].shape[0]
    ks = [a.shape[1] for a in A_parts]
    B = _f32c(B)
    n = B.shape[0]
    if B.shape[1] != sum(ks):
        raise ValueError("gemm_nt: B has %d columns, A parts have %s" % (B.shape[1], ks))
    n1 = n if n1 is None else n1
    dev = B.device
    splits = a_splits if a_splits is not None else [split_tf32(a, _pad16(a.shape[1])) for a in A_parts]
    # B's K axis is laid out part by part, each padded to 16, to match the k-blocks of the A parts
    if len(ks) == 1 and ks[0] % 16 == 0:
        Bp = B
    else:
        Bp = torch.zeros(n, sum(_pad16(k) for k in ks), dtype=torch.float32, device=dev)
        src = dst = 0
        for k in ks:
            Bp[:, dst:dst + k] = B[:, src:src + k]
            src += k
            dst += _pad16(k)
    b_hi, b_lo = split_tf32(Bp, Bp.shape[1])
    out1 = torch.empty(m, n1, dtype=torch.float32, device=dev)
    out2 = torch.empty(m, n - n1, dtype=torch.float32, device=dev) if n1 < n else None
    a2_hi, a2_lo = (splits[1] if len(splits) == 2 else (None, None))
    bias = _f32c(bias) if bias is not None else None
    with torch.cuda.device(dev):
        fn = lib.eg_gemm_nt_3xtf32_chained if chained else lib.eg_gemm_nt_3xtf32
        check(fn(ptr(splits[0][0]), ptr(splits[0][1]), _pad16(ks[0]), ptr(a2_hi), ptr(a2_lo),
                 _pad16(ks[1]) if len(ks) == 2 else 0, m, ptr(b_hi), ptr(b_lo), n, ptr(bias),
                 ptr(out1), n1, n1, ptr(out2), (n - n1) if out2 is not None else 0, stream()),
              "eg_gemm_nt_3xtf32_chained" if chained else "eg_gemm_nt_3xtf32")
    res = (out1, out2) if out2 is not None else out1
    return (res, splits) if return_splits else res


# ---- margin ranking loss (fused gather + L1 + hinge) --------------------------------------------

class _MarginLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, left, right, nl, nr, n2l, n2r, k, gamma):
        out = _f32c(outputs)
        t = int(left.numel())
        loss = torch.zeros(1, dtype=torch.float64, device=out.device)
        with torch.cuda.device(out.device):
            check(lib.eg_margin_loss_fwd(ptr(out), out.shape[0], out.shape[1], ptr(left), ptr(right), ptr(nl), ptr(nr),
                                         ptr(n2l), ptr(n2r), t, int(k), float(gamma), ptr(loss), stream()),
                  "eg_margin_loss_fwd")
        ctx.save_for_backward(out, left, right, nl, nr, n2l, n2r)
        ctx.k, ctx.gamma, ctx.t = int(k), float(gamma), t
        return (loss[0] / (2.0 * t * k)).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        out, left, right, nl, nr, n2l, n2r = ctx.saved_tensors
        grad = torch.zeros_like(out)
        scale = float(g) / (2.0 * ctx.t * ctx.k)
        with torch.cuda.device(out.device):
            check(lib.eg_margin_loss_bwd(ptr(out), out.shape[0], out.shape[1], ptr(left), ptr(right), ptr(nl), ptr(nr),
                                         ptr(n2l), ptr(n2r), ctx.t, ctx.k, ctx.gamma, scale, ptr(grad), stream()),
                  "eg_margin_loss_bwd")
        return grad, None, None, None, None, None, None, None, None


_MARGIN_INDEX_OK = {}     # what -> (data_ptr, numel, version, n) of device index tensors already range-checked


def _row_indices(idx, n, what, dev):
    """int64 CUDA copy of an index array, range-checked against [0, n): the kernels index ``outputs`` without
    bounds checks where the reference's fancy indexing raises IndexError (models/models_ea.py:108-118).  Host
    arrays (what the reference's models hold) are checked on the host before the upload; device tensors with one
    device-side min/max per *new* tensor, so the steady-state step adds no host synchronisation."""
    import numpy as np
    if not torch.is_tensor(idx) or not idx.is_cuda:
        host = idx.numpy() if torch.is_tensor(idx) else np.asarray(idx)
        if host.size and (host.min() < 0 or host.max() >= n):
            raise IndexError("margin_loss: %s has an index outside [0, %d) (min %d, max %d)"
                             % (what, n, int(host.min()), int(host.max())))
        return torch.as_tensor(host).to(device=dev, dtype=torch.int64).contiguous()
    idx = idx.to(device=dev, dtype=torch.int64).contiguous()
    key = (idx.data_ptr(), idx.numel(), idx._version, int(n))
    if idx.numel() and _MARGIN_INDEX_OK.get(what) != key:
        lo, hi = (int(v) for v in torch.aminmax(idx))
        if lo < 0 or hi >= n:
            raise IndexError("margin_loss: %s has an index outside [0, %d) (min %d, max %d)" % (what, n, lo, hi))
        _MARGIN_INDEX_OK[what] = key
    return idx


def margin_loss(outputs, left, right, nl, nr, n2l, n2r, k, gamma=1.0):
    """(sum relu(A+gamma-B1) + sum relu(A+gamma-B2)) / (2 t k) — models/models_ea.py:103-123."""
    _lib.require_cuda(outputs)
    dev = outputs.device

    def ix(a, what):
        return _row_indices(a, outputs.shape[0], what, dev)
    left, right = ix(left, "left"), ix(right, "right")
    nl, nr, n2l, n2r = ix(nl, "neg_left"), ix(nr, "neg_right"), ix(n2l, "neg2_left"), ix(n2r, "neg2_right")
    if not (left.numel() == right.numel() and nl.numel() == nr.numel() == n2l.numel() == n2r.numel()
            == left.numel() * int(k)):
        raise ValueError("margin_loss: index arrays disagree (t = %d, k = %d)" % (left.numel(), int(k)))
    return _MarginLoss.apply(outputs, left, right, nl, nr, n2l, n2r, k, gamma)


# ---- GAT edge-softmax aggregation (layers/att_layers.py:29-61) ------------------------------------

class _GatAggregate(torch.autograd.Function):
    """y_i = (sum_j m_ij w_ij h_j) / sum_j w_ij, w_ij = exp(-leakyrelu(s1_i + s2_j)) over the edges of A."""

    @staticmethod
    def forward(ctx, h, s1, s2, adjacency, alpha, edge_scale):
        c = adjacency.csr
        h, s1, s2 = _f32c(h), _f32c(s1), _f32c(s2)
        if h.shape[0] != c.n_cols or s1.numel() != c.n_rows or s2.numel() != c.n_cols:
            raise ValueError("gat_aggregate: h/s1/s2 do not match the adjacency shape")
        y = torch.empty(c.n_rows, h.shape[1], dtype=torch.float32, device=h.device)
        wsum = torch.empty(c.n_rows, dtype=torch.float32, device=h.device)
        with torch.cuda.device(h.device):
            check(lib.eg_gat_fwd(ptr(c.rowptr), ptr(c.col), c.n_rows, ptr(h), h.shape[1], ptr(s1), ptr(s2),
                                 float(alpha), ptr(edge_scale), ptr(y), ptr(wsum), c.threshold, ptr(c.seg_row),
                                 ptr(c.seg_begin), ptr(c.seg_end), c.n_seg, ptr(c.long_rows), ptr(c.long_first),
                                 c.n_long, ptr(c.scratch(h.shape[1] + 1)), stream()), "eg_gat_fwd")
        ctx.save_for_backward(h, s1, s2, y, wsum)
        ctx.adjacency, ctx.alpha, ctx.edge_scale = adjacency, float(alpha), edge_scale
        return y

    @staticmethod
    def backward(ctx, dy):
        h, s1, s2, y, wsum = ctx.saved_tensors
        adjacency = ctx.adjacency
        c, ct = adjacency.csr, adjacency.csr_t
        dy = _f32c(dy)
        dev = h.device
        p_edge = torch.empty(max(c.nnz, 1), dtype=torch.float32, device=dev)
        p_t = torch.empty_like(p_edge)
        ds1 = torch.empty_like(s1)
        ds2 = torch.empty_like(s2)
        with torch.cuda.device(dev):
            check(lib.eg_gat_bwd_edges(ptr(c.rowptr), ptr(c.col), c.n_rows, c.n_cols, ptr(h), h.shape[1], ptr(s1),
                                       ptr(s2), ctx.alpha, ptr(ctx.edge_scale), ptr(y), ptr(wsum), ptr(dy),
                                       ptr(p_edge), ptr(ds1), ptr(ds2), c.threshold, ptr(c.seg_row),
                                       ptr(c.seg_begin), ptr(c.seg_end), c.n_seg, stream()), "eg_gat_bwd_edges")
            check(lib.eg_permute_edges(ptr(p_edge), ptr(ct.perm), c.nnz, ptr(p_t), stream()), "eg_permute_edges")
            dh = torch.empty_like(h)
            scratch = ct.scratch(h.shape[1])
            check(lib.eg_spmm(ptr(ct.rowptr), ptr(ct.col), ptr(p_t), ct.n_rows, ptr(dy), h.shape[1],
                              _lib.ACT_IDENTITY, None, None, ptr(dh), None, ct.threshold, ptr(ct.seg_row),
                              ptr(ct.seg_begin), ptr(ct.seg_end), ct.n_seg, ptr(ct.long_rows), ptr(ct.long_first),
                              ct.n_long, ptr(scratch), stream()), "eg_spmm")
        return dh, ds1, ds2, None, None, None


def gat_aggregate(h, s1, s2, adjacency, alpha, edge_scale=None):
    """Differentiable (h, s1, s2) -> y; ``adjacency`` is a DeviceAdjacency (only its structure is used)."""
    _lib.require_cuda(h, s1, s2, edge_scale)
    if getattr(adjacency, "sharded", False):
        raise NotImplementedError("gat_aggregate: row-sharded adjacencies are not supported")
    if edge_scale is not None:
        edge_scale = _f32c(edge_scale)
        if edge_scale.numel() != adjacency.csr.nnz:
            raise ValueError("gat_aggregate: edge_scale must have one entry per stored edge")
    return _GatAggregate.apply(h, s1.reshape(-1), s2.reshape(-1), adjacency, alpha, edge_scale)
