"""Cost matrices of SinkhornOT/cderivation.py:14-61, for callers that want the
matrix itself.  (Inside the fused Sinkhorn these are computed per tile in the
kernel and never stored.)  Plain tensor algebra; GEMM-shaped via ``x @ yᵀ``."""
import torch


def p_norm_dist_mat(x, y, p=2):
    """sum_k (x_k - y_k)^p, no root (:14-26)."""
    assert x.shape[1] == y.shape[1]
    if p == 2:
        sq = (x * x).sum(1, keepdim=True) + (y * y).sum(1)[None, :] - 2.0 * (x @ y.t())
        return sq.clamp_min(0)
    return ((x[:, None, :] - y[None, :, :]) ** p).sum(-1)


def norm_dist_mat(x, y, p=2):
    """(:29-37)"""
    return p_norm_dist_mat(x, y, p) ** (1.0 / p)


def cos_dist_mat(x, y):
    """1 - cosine similarity with torch's eps of 1e-8 (:44-61)."""
    assert x.shape[1] == y.shape[1]
    xn = x / x.norm(dim=1, keepdim=True).clamp_min(1e-8)
    yn = y / y.norm(dim=1, keepdim=True).clamp_min(1e-8)
    return 1 - xn @ yn.t()
