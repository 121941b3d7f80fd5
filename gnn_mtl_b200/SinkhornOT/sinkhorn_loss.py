"""``sinkhorn_iteration`` with the reference's signature
(SinkhornOT/sinkhorn_loss.py:159-220) on the eagraft log-domain kernels.

The reference alternates a = mu / (K b), b = nu / (Kᵀ a) on K = exp((u+v-C)/eps)
and folds eps·log a, eps·log b into (u, v) every 10 sweeps ("absorption").  In
log form the absorbed and un-absorbed states are the same numbers, so here the
sweeps run on alpha = u/eps + log a and beta = v/eps + log b directly; the
absorption schedule only decides WHEN the primal cost is evaluated and the
stopping test made (sweeps 0, 10, 20, … and the last), which is reproduced.
The reference's clamp a, b <= 1e30 (:14-15,197-201) and its early absorption when
a or b exceeds 1e20 (:203) are reproduced in log form (they change the iterates
for small eps); the clamp on K itself (:184-188) cannot trigger once a, b are bounded.
"""
from __future__ import annotations

import math

import torch

from .. import _lib, ops

small = 1e-7


def kl_div(x, y):
    """SinkhornOT/sinkhorn_loss.py:20-30."""
    div = torch.div(x, y + small)
    return torch.mul(y, div * torch.log(div + small) - div + 1)


LOG_HUGE = math.log(1e30)   # myclamp upper bound of the reference (:12,14-15)
LOG_BIG = math.log(1e20)    # early-absorption trigger (:11,203)


def _solve_one(C, mu, nu, epsilon, numIterMax, tol):
    """alpha = u/eps + log a, beta = v/eps + log b of the reference's state; (u_abs, v_abs)
    are the potentials as of the last absorption, needed only to reproduce the
    reference's clamp a, b <= 1e30 (it bites for small eps in the first sweeps) and
    its "absorb when a or b exceeds 1e20" rule, which adds stopping tests."""
    I, J = C.shape
    inv = 1.0 / epsilon
    Ct = ops.transpose(C)
    log_mu, log_nu = torch.log(mu), torch.log(nu)
    alpha = torch.zeros(I, dtype=C.dtype, device=C.device)
    beta = torch.zeros(J, dtype=C.dtype, device=C.device)
    u_abs, v_abs = alpha, beta
    _, transport, _, _ = ops.plan_dense(C, inv, alpha, beta, want_plan=False)
    transport = transport.to(C.dtype)
    transport_new = transport
    for ii in range(numIterMax):
        alpha, _ = ops.lse_dense(C, inv, beta, log_mu)      # a = mu / (K b)
        alpha = torch.minimum(alpha, u_abs + LOG_HUGE)
        beta, _ = ops.lse_dense(Ct, inv, alpha, log_nu)     # b = nu / (K^T a)
        beta = torch.minimum(beta, v_abs + LOG_HUGE)
        absorb = ii % 10 == 0 or ii == numIterMax - 1
        if not absorb:
            absorb = bool(((alpha - u_abs).max() > LOG_BIG) | ((beta - v_abs).max() > LOG_BIG))
        if absorb:
            u_abs, v_abs = alpha, beta
            _, transport_new, _, _ = ops.plan_dense(C, inv, alpha, beta, want_plan=False)
            transport_new = transport_new.to(C.dtype)
            if abs(transport_new - transport) / abs(transport) < tol:
                break
            transport = transport_new
    K, _, rows, cols = ops.plan_dense(C, inv, alpha, beta, want_plan=True, want_rows=True, want_cols=True)
    return transport_new, rows, cols, K


def sinkhorn_iteration(C, mu, nu, epsilon, numIterMax=100, tol=1e-9, debug=True):
    """C [*, I, J], mu [*, I, 1], nu [*, 1, J] -> (transport, margin1, margin2, K)."""
    *_, I, J = C.shape
    if debug:
        assert mu.shape[-2] == I and nu.shape[-1] == J
        assert len(C.shape) == len(mu.shape) == len(nu.shape)
    _lib.require_cuda(C)
    batched = C.dim() == 3
    Cb = C.detach().contiguous().reshape(-1, I, J)
    B = Cb.shape[0]
    mub = mu.detach().to(C.dtype).expand(*C.shape[:-2], I, 1).reshape(B, I).contiguous()
    nub = nu.detach().to(C.dtype).expand(*C.shape[:-2], 1, J).reshape(B, J).contiguous()
    ts, ks, rs, cs = [], [], [], []
    for bidx in range(B):
        t, rows, cols, K = _solve_one(Cb[bidx], mub[bidx], nub[bidx], epsilon, numIterMax, tol)
        ts.append(t); ks.append(K); rs.append(rows); cs.append(cols)
    K = torch.stack(ks).reshape(C.shape)
    row_marg = torch.stack(rs).reshape(*C.shape[:-2], I, 1)
    col_marg = torch.stack(cs).reshape(*C.shape[:-2], 1, J)
    transport = torch.stack(ts).squeeze() if batched else ts[0]
    margin1 = torch.sum(kl_div(row_marg, mu.to(C.dtype)), -2).squeeze()
    margin2 = torch.sum(kl_div(col_marg, nu.to(C.dtype)), -1).squeeze()
    return transport, margin1, margin2, K
