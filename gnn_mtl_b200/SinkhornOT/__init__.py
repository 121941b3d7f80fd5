from .sinkhorn_loss import sinkhorn_iteration  # noqa: F401
from .cderivation import (p_norm_dist_mat, norm_dist_mat, cos_dist_mat, get_inter_sim, get_intra_sim,  # noqa: F401
                          get_init_matrices, get_LT, GW_cost_matrix, FGW_cost_matrix)
from .iterative_projection import gw_iterative_1, fgw_iterative_1  # noqa: F401
