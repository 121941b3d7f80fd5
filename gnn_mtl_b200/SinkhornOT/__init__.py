from .sinkhorn_loss import sinkhorn_iteration  # noqa: F401
from .cderivation import p_norm_dist_mat, norm_dist_mat, cos_dist_mat  # noqa: F401
