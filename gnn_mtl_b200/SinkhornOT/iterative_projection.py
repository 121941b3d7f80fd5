"""Gromov-Wasserstein by iterative projection: the outer loop of the reference's
SinkhornOT/iterative_projection.py:8-60 around the eagraft `sinkhorn_iteration`.

Each projection linearises the GW objective at the current plan, L(T) = constC - C1·T·C2ᵀ
(cderivation.py:147-163, two dense GEMMs on cuBLAS), and solves the entropic OT problem with cost 2·L by the
log-domain kernels.  Only the balanced variants (`gw_iterative_1`, `fgw_iterative_1`) are provided; the relaxed
ones need `forward_relax_sinkhorn_iteration`, which is outside the hot path.
"""
import torch

from .cderivation import FGW_cost_matrix, GW_cost_matrix, get_init_matrices
from .sinkhorn_loss import sinkhorn_iteration


def iterative_1(C1, C2, mu, nu, epsilon, max_iter, log, tol=1e-9, g=False, cost_mat_func=GW_cost_matrix, lambdda=0):
    if g:
        raise NotImplementedError("relaxed (unbalanced) GW needs forward_relax_sinkhorn_iteration — out of scope")
    I, J = C1.shape[0], C2.shape[0]
    assert C1.device == C2.device
    mu = mu.view(1, I, 1)
    nu = nu.view(1, 1, J)
    T_old = torch.full((I, J), 1.0 / (I * J), dtype=C1.dtype, device=C1.device)
    constC, hC1, hC2 = get_init_matrices(C1, C2, mu, nu)
    lt, _ = cost_mat_func(constC, hC1, hC2, T_old, epsilon)
    gw_dist = torch.sum(T_old * lt)
    trace = {'err': [], 'gwd': []} if log else None
    T = T_old
    for _ in range(max_iter):
        gw_dist, *_, T = sinkhorn_iteration(2 * lt.view(1, I, J), mu, nu, epsilon)
        err = torch.norm(T_old - T)
        if log:
            trace['err'].append(float(err))
            trace['gwd'].append(float(gw_dist))
        if err < tol:
            break
        T_old = T
        lt, _ = GW_cost_matrix(constC, hC1, hC2, T_old.reshape(I, J), epsilon)
    if log:
        trace['gw_dist'] = float(gw_dist) / 2
        return T, trace
    return T, gw_dist


def gw_iterative_1(C1, C2, mu, nu, epsilon, max_iter, log=False, tol=1e-9):
    return iterative_1(C1, C2, mu, nu, epsilon, max_iter, log, tol, False, GW_cost_matrix)


def fgw_iterative_1(D, C1, C2, mu, nu, alpha, p, max_iter, epsilon, log=False, tol=1e-6):
    def cost_matrix_func(constC, hC1, hC2, T, eps):
        return FGW_cost_matrix(D, constC, hC1, hC2, T, alpha, eps, p)
    return iterative_1(C1, C2, mu, nu, epsilon, max_iter, log, tol, False, cost_matrix_func)
