"""Device-resident adjacency in the layout the SpMM kernels read.

Host-side mirror of the reference's adjacency objects (a scipy COO matrix from
utils/data_utils.py:325-336 and the torch sparse tensor made from it at :51-57).
Holds, all in HBM:
  * the contract arrays  crow int64 [n+1], col int64 [nnz], val fp32 [nnz]
    (== reference tensor .coalesce().to_sparse_csr(), bit for bit), and
  * the kernel arrays    rowptr int32, col int32, val fp32  for A and for Aᵀ
    (the transposed-backward pass reads Aᵀ as CSR), plus the segment lists that
    cut hub rows into fixed-length pieces.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, ptr, stream, check

LONG_ROW_THRESHOLD = 512   # nnz; rows above this are split into segments of this length


class _Csr:
    """Kernel-format CSR (+ hub-row segmentation) of one matrix."""

    def __init__(self, n_rows, n_cols, rowptr, col, val, threshold=LONG_ROW_THRESHOLD):
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.rowptr, self.col, self.val = rowptr, col, val
        self.nnz = int(col.numel())
        self.threshold = int(threshold)
        dev = rowptr.device
        deg = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
        long_rows = torch.nonzero(deg > threshold).reshape(-1)
        self.n_long = int(long_rows.numel())
        if self.n_long:
            n_seg_per = (deg[long_rows] + threshold - 1) // threshold
            first = torch.zeros(self.n_long + 1, dtype=torch.int64, device=dev)
            first[1:] = torch.cumsum(n_seg_per, 0)
            n_seg = int(first[-1])
            owner = torch.repeat_interleave(torch.arange(self.n_long, device=dev), n_seg_per)
            within = torch.arange(n_seg, device=dev) - first[owner]
            begin = rowptr[long_rows].to(torch.int64)[owner] + within * threshold
            end = torch.minimum(begin + threshold, rowptr[long_rows + 1].to(torch.int64)[owner])
            self.seg_row = long_rows[owner].to(torch.int32).contiguous()
            self.seg_begin = begin.to(torch.int32).contiguous()
            self.seg_end = end.to(torch.int32).contiguous()
            self.long_rows = long_rows.to(torch.int32).contiguous()
            self.long_first = first.to(torch.int32).contiguous()
            self.n_seg = n_seg
        else:
            self.seg_row = self.seg_begin = self.seg_end = self.long_rows = self.long_first = None
            self.n_seg = 0
        self._scratch = {}

    def scratch(self, d):
        if not self.n_seg:
            return None
        buf = self._scratch.get(d)
        if buf is None:
            buf = torch.empty(self.n_seg * d, dtype=torch.float32, device=self.rowptr.device)
            self._scratch[d] = buf
        return buf


class DeviceAdjacency:
    """Degree-normalised adjacency living on the GPU (see module docstring)."""

    def __init__(self, n, crow, col, val, rowptr32, col32, symmetric_hint=False):
        self.n = int(n)
        self.shape = (self.n, self.n)
        self.crow, self.col64, self.val = crow, col, val
        self.csr = _Csr(n, n, rowptr32, col32, val)
        self._csr_t = None
        self.symmetric_hint = symmetric_hint

    @property
    def nnz(self):
        return self.csr.nnz

    @property
    def device(self):
        return self.val.device

    # ---- construction --------------------------------------------------------
    @classmethod
    def from_triples(cls, n_ent, triples, device=None):
        """triples: sequence of (h, r, t) or int array [T,3] (or [T,2] = (h,t)).
        Builds on the device with eg_adj_build (utils/data_utils.py:296-336)."""
        device = torch.device(device or "cuda")
        if device.type != "cuda":
            raise _lib.EagraftError("adjacency is built on a CUDA device; there is no CPU path")
        arr = np.asarray(triples, dtype=np.int64)
        arr = arr.reshape(arr.shape[0], -1) if arr.size else np.zeros((0, 3), dtype=np.int64)
        heads = torch.from_numpy(np.ascontiguousarray(arr[:, 0])).to(device)
        tails = torch.from_numpy(np.ascontiguousarray(arr[:, -1])).to(device)
        return cls.from_heads_tails(n_ent, heads, tails)

    @classmethod
    def from_heads_tails(cls, n_ent, heads, tails):
        _lib.require_cuda(heads, tails)
        n_ent = int(n_ent)
        if heads.numel() and (int(torch.max(heads)) >= n_ent or int(torch.max(tails)) >= n_ent
                              or int(torch.min(heads)) < 0 or int(torch.min(tails)) < 0):
            raise IndexError("triple endpoint outside [0, n_ent)")
        heads = heads.to(torch.int64).contiguous()
        tails = tails.to(torch.int64).contiguous()
        dev = heads.device
        T = int(heads.numel())
        cap = 2 * T + n_ent
        with torch.cuda.device(dev):
            ws_bytes = int(lib.eg_adj_workspace_bytes(T, n_ent))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            crow = torch.empty(n_ent + 1, dtype=torch.int64, device=dev)
            col = torch.empty(cap, dtype=torch.int64, device=dev)
            val = torch.empty(cap, dtype=torch.float32, device=dev)
            rowptr32 = torch.empty(n_ent + 1, dtype=torch.int32, device=dev)
            col32 = torch.empty(cap, dtype=torch.int32, device=dev)
            nnz = C.c_int64(0)
            check(lib.eg_adj_build(ptr(heads), ptr(tails), T, n_ent, ptr(ws), ws_bytes, ptr(crow), ptr(col),
                                   ptr(val), ptr(rowptr32), ptr(col32), C.byref(nnz), stream()), "eg_adj_build")
        k = int(nnz.value)
        # slices keep 16-byte alignment of the base allocations
        return cls(n_ent, crow, col[:k].clone(), val[:k].clone(), rowptr32, col32[:k].clone(),
                   symmetric_hint=True)

    @classmethod
    def from_torch_sparse(cls, adj):
        """Any square torch sparse tensor (COO or CSR) already on the GPU; the
        layers accept whatever the caller passes (layers/layers.py:34)."""
        _lib.require_cuda(adj)
        if adj.layout == torch.sparse_coo:
            adj = adj.coalesce().to_sparse_csr()
        elif adj.layout != torch.sparse_csr:
            raise TypeError("unsupported sparse layout %s" % adj.layout)
        n, m = adj.shape
        if n != m:
            raise ValueError("adjacency must be square")
        crow = adj.crow_indices().to(torch.int64).contiguous()
        col = adj.col_indices().to(torch.int64).contiguous()
        val = adj.values().to(torch.float32).contiguous()
        return cls(n, crow, col, val, crow.to(torch.int32), col.to(torch.int32))

    # ---- views -----------------------------------------------------------------
    @property
    def csr_t(self):
        """CSR of Aᵀ, built on first use (eg_csr_transpose)."""
        if self._csr_t is None:
            c = self.csr
            dev = self.device
            with torch.cuda.device(dev):
                ws_bytes = int(lib.eg_csr_transpose_workspace_bytes(c.nnz, c.n_rows, c.n_cols))
                ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
                rp = torch.empty(c.n_cols + 1, dtype=torch.int32, device=dev)
                ct = torch.empty(max(c.nnz, 1), dtype=torch.int32, device=dev)[:c.nnz]
                vt = torch.empty(max(c.nnz, 1), dtype=torch.float32, device=dev)[:c.nnz]
                pt = torch.empty(max(c.nnz, 1), dtype=torch.int32, device=dev)[:c.nnz]
                check(lib.eg_csr_transpose(c.n_rows, c.n_cols, c.nnz, ptr(c.rowptr), ptr(c.col), ptr(c.val), ptr(ws),
                                           ws.numel(), ptr(rp), ptr(ct), ptr(vt), ptr(pt), stream()),
                      "eg_csr_transpose")
            self._csr_t = _Csr(c.n_cols, c.n_rows, rp, ct, vt)
            self._csr_t.perm = pt        # position in CSR(A) of every entry of CSR(Aᵀ)
        return self._csr_t

    def to_torch_coo(self):
        """Coalesced torch sparse COO on the same device (what the reference's
        layers receive), with this object attached so our layers skip re-analysis."""
        rows = torch.repeat_interleave(torch.arange(self.n, device=self.device), self.crow[1:] - self.crow[:-1])
        t = torch.sparse_coo_tensor(torch.stack([rows, self.col64]), self.val, self.shape, is_coalesced=True)
        t._eg_adj = self
        return t

    def tocoo(self):
        """scipy COO on the host (row-major sorted) — shape-compatible with what
        the reference's get_sparse_tensor returns."""
        import scipy.sparse as sp
        crow = self.crow.cpu().numpy()
        rows = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(crow))
        m = sp.coo_matrix((self.val.cpu().numpy().astype(np.float64), (rows, self.col64.cpu().numpy())),
                          shape=self.shape)
        m._eg_adj = self
        return m


def resolve(adj):
    """Whatever a layer was handed -> DeviceAdjacency (cached on the object)."""
    if isinstance(adj, DeviceAdjacency) or getattr(adj, "sharded", False):
        return adj
    cached = getattr(adj, "_eg_adj", None)
    if cached is not None and cached.device == adj.device:
        return cached
    if torch.is_tensor(adj) and adj.layout in (torch.sparse_coo, torch.sparse_csr):
        built = DeviceAdjacency.from_torch_sparse(adj)
        try:
            adj._eg_adj = built
        except Exception:
            pass
        return built
    raise TypeError("expected a sparse adjacency, got %r" % type(adj))
