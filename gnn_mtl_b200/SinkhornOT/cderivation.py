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


small = 1e-7


# ---- Gromov-Wasserstein helpers (SinkhornOT/cderivation.py:138-189) ----------------------------------------

def get_intra_sim(x, sim_func):
    x = x.detach()
    return sim_func(x, x)


def get_inter_sim(x, y, sim_func):
    return sim_func(x.detach(), y.detach())


def get_init_matrices(C1, C2, mu, nu, div_type='l2'):
    """constC[i,j] = ½ Σ_k C1[i,k]² mu_k + ½ Σ_l nu_l C2[j,l]²  (:147-158), hC1 = C1, hC2 = C2."""
    I, J = C1.shape[0], C2.shape[0]
    a = 0.5 * (C1 ** 2) @ mu.reshape(I, 1)
    b = 0.5 * nu.reshape(1, J) @ (C2.t() ** 2)
    return a + b, C1, C2


def get_LT(constC, hC1, hC2, T):
    """constC - hC1·T·hC2ᵀ (:161-163)."""
    return constC - hC1 @ (T @ hC2.t())


def GW_cost_matrix(constC, hC1, hC2, T_old, epsilon):
    lt = get_LT(constC, hC1, hC2, T_old)
    return lt, lt - epsilon * torch.log(T_old + small)


def FGW_cost_matrix(D, constC, hC1, hC2, T, alpha, epsilon, p):
    A = (1 - alpha) * D ** p + alpha * get_LT(constC, hC1, hC2, T) ** p
    return A, A - epsilon * torch.log(T)
