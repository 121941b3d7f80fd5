"""ctypes binding of libeagraft.so (the C ABI declared in include/eagraft.h).

There is no other implementation behind these calls: if the shared library is
missing, or a call returns an error, this module raises.  PyTorch is used by the
callers for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libeagraft.so")

EG_OK = 0
ACT_IDENTITY, ACT_RELU = 0, 1
COST_L2, COST_SQEUCLID, COST_COSINE = 0, 1, 2
ALGO_SIMT, ALGO_TCGEN05 = 0, 1
DT_F32, DT_F64 = 0, 1

_vp, _i64, _i32, _f64, _f32, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/eagraft.h declares.
SIGNATURES = {
    "eg_version": (C.c_int, []),
    "eg_strerror": (C.c_char_p, [C.c_int]),
    "eg_last_cuda_error": (C.c_int, []),
    "eg_device_check": (C.c_int, []),
    "eg_launch_count": (_i64, []),
    "eg_launch_count_reset": (None, []),
    "eg_debug_set": (C.c_int, [C.c_int, C.c_int]),
    "eg_adj_workspace_bytes": (_sz, [_i64, _i64]),
    "eg_adj_build": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _sz, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i64), _vp]),
    "eg_csr_transpose_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "eg_csr_transpose": (C.c_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "eg_spmm": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _i32,
                          _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp]),
    "eg_epilogue_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "eg_l1_matrix": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp]),
    "eg_l1_paired": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "eg_rank_accumulate": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "eg_argmin_accumulate": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "eg_topk_rows": (C.c_int, [_vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp]),
    "eg_lse_dense": (C.c_int, [_i32, _vp, _i64, _i64, _i64, _f64, _vp, _vp, _vp, _vp, _vp]),
    "eg_transpose": (C.c_int, [_i32, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "eg_plan_dense": (C.c_int, [_i32, _vp, _i64, _i64, _i64, _f64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "eg_sinkhorn_dense_workspace_bytes": (_sz, [_i32, _i64, _i64]),
    "eg_sinkhorn_sync_floor": (C.c_int, [_i64, _i64, _i32, _vp, _sz, _vp]),
    "eg_issue_peak": (C.c_int, [_i32, _i32, C.POINTER(C.c_double), _vp, _vp]),
    "eg_sinkhorn_dense": (C.c_int, [_i32, _vp, _i64, _i64, _f64, _vp, _vp, _i32, _f64, _vp, _vp, _vp, _vp, _sz,
                                    C.POINTER(C.c_int), C.POINTER(_f64), _vp]),
    "eg_row_norms": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "eg_lse_fused_workspace_bytes": (_sz, [_i32, _i64, _i64, _i32]),
    "eg_lse_fused": (C.c_int, [_i32, _i32, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _f32, _vp, _vp, _vp, _vp,
                               _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "eg_split_tf32": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "eg_plan_fused": (C.c_int, [_i32, _i32, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _f32, _vp, _vp, _vp, _i64, _vp, _vp,
                                _vp, _vp, _vp, _vp, _vp]),
    "eg_plan_grad_fused_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "eg_plan_grad_fused": (C.c_int, [_i32, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _f32, _vp, _vp, _f32, _vp, _sz, _vp, _vp]),
    "eg_gemm_nt_3xtf32": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _i32, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _vp,
                                    _i64, _vp]),
    "eg_gemm_nt_3xtf32_chained": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _i32, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _vp,
                                    _i64, _vp]),
    "eg_gemm_nt_3xtf32_raw": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _i32, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _i64, _vp,
                                        _i64, _i64, _vp, _i64, _vp]),
    "eg_margin_loss_fwd": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp, _vp]),
    "eg_margin_loss_bwd": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _vp, _vp]),
    "eg_gemm_tn_3xtf32_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "eg_gemm_tn_3xtf32": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _i32, _i32, _i64, _vp, _sz, _vp, _i64, _vp]),
    "eg_l1_rank_fused": (C.c_int, [_vp, _i64, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "eg_l1_rank_filtered_workspace_bytes": (_sz, [_i64, _i64]),
    "eg_l1_rank_filtered": (C.c_int, [_vp, _i64, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "eg_l1_topk_fused_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "eg_l1_topk_fused": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _sz, _vp, _vp]),
    "eg_gat_fwd": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _vp, _vp, _f32, _vp, _vp, _vp,
                           _i32, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp]),
    "eg_gat_bwd_edges": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i32, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _i32, _vp, _vp, _vp, _i64, _vp]),
    "eg_permute_edges": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
}


class EagraftError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libeagraft.so not found at %s — build it with `python __graft_entry__.py build` "
            "(or `make -C gnn_mtl_b200/csrc`). There is no fallback path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int, what: str = ""):
    if rc != EG_OK:
        msg = lib.eg_strerror(rc).decode()
        extra = ""
        if rc == -2:
            code = lib.eg_last_cuda_error()
            extra = " [cudaError %d]" % code
        raise EagraftError("%s failed: %s%s" % (what or "eagraft call", msg, extra))


def ptr(t):
    """Device (or host) pointer of a tensor, None -> NULL."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise EagraftError("eagraft kernels run on CUDA tensors only (got a %s tensor); "
                               "there is no CPU path" % t.device)


def launch_count() -> int:
    return int(lib.eg_launch_count())


def reset_launch_count():
    lib.eg_launch_count_reset()
