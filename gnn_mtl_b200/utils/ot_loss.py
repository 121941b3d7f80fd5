"""``sinkhorn`` with the reference's signature (utils/ot_loss.py:5), solved in the
log domain on the GPU by eg_sinkhorn_dense.

Differences a caller can observe, all documented in DESIGN.md:
  * the iteration is algebraically the same (u0 = 1/I, v0 = 1/J, v then u per
    sweep, marginal-error check on sweeps 0,10,…, same stopping rule) but runs on
    log u / log v, so the reference's "numerical errors" bail-out (:57-62, Gibbs
    kernel underflow) never triggers;
  * arithmetic is fp32 when M is fp32 and fp64 when M is fp64 (the reference
    always widens to fp64); returned P / loss are float64 like the reference's.
"""
from __future__ import annotations

import torch

from .. import _lib, ops


def sinkhorn(a, b, M, reg, numItermax=1000, stopThr=1e-9, verbose=False, *, return_plan=True, info=None):
    """Entropic OT.  a [I], b [J], M [I,J] on one CUDA device -> (P [I,J] fp64, <P,M> fp64).

    Extra keyword-only arguments (not in the reference): ``return_plan=False``
    skips materialising P (returns None); ``info`` (a dict) receives sweeps, err,
    log_u, log_v.
    """
    assert a.device == b.device and b.device == M.device, "a, b, M must be on the same device"
    _lib.require_cuda(M)
    I, J = M.shape
    if len(a) == 0:
        a = torch.ones(I, dtype=torch.float64, device=M.device) / I
    if len(b) == 0:
        b = torch.ones(J, dtype=torch.float64, device=M.device) / J
    assert len(a) == I and len(b) == J, "the dimension of weights and distance matrix don't match"
    work = torch.float64 if M.dtype == torch.float64 else torch.float32
    Mw = M.detach().to(work).contiguous()
    aw = a.detach().to(work).contiguous().reshape(-1)
    bw = b.detach().to(work).contiguous().reshape(-1)
    log_u, log_v, sweeps, err = ops.sinkhorn_dense(Mw, aw, bw, reg, numItermax, stopThr)
    if verbose:
        print("sinkhorn: {} sweeps, marginal err {:.3e}".format(sweeps, err))
    P, loss, _, _ = ops.plan_dense(Mw, 1.0 / reg, log_u, log_v, want_plan=return_plan)
    if info is not None:
        info.update(sweeps=sweeps, err=err, log_u=log_u, log_v=log_v)
    return (P.to(torch.float64) if P is not None else None), loss


def sinkhorn_fused(X, Y, a, b, reg, numItermax=1000, stopThr=1e-9, cost="l2", algo="tcgen05",
                   return_plan=False, info=None):
    """Same solver with the cost recomputed tile by tile from the embeddings
    (models/models_ea.py:218 fused into utils/ot_loss.py:53-55); the I×J cost is
    never stored.  X [I,d], Y [J,d] fp32.  Returns (P or None, <P,M> fp64)."""
    _lib.require_cuda(X, Y)
    cost_id = {"l2": _lib.COST_L2, "sqeuclid": _lib.COST_SQEUCLID, "cos": _lib.COST_COSINE}[cost]
    algo_id = {"simt": _lib.ALGO_SIMT, "tcgen05": _lib.ALGO_TCGEN05}[algo]
    A = ops.FusedOperand(X.detach(), cost_id, algo_id)
    B = ops.FusedOperand(Y.detach(), cost_id, algo_id)
    I, J = A.n, B.n
    dev = A.X.device
    log_a = torch.log(a.detach().to(torch.float32).reshape(-1)).contiguous()
    log_b = torch.log(b.detach().to(torch.float32).reshape(-1)).contiguous()
    b32 = b.detach().to(torch.float32).reshape(-1)
    import math
    log_u = torch.full((I,), -math.log(I), dtype=torch.float32, device=dev)
    log_v = torch.full((J,), -math.log(J), dtype=torch.float32, device=dev)
    inv_reg = 1.0 / reg
    err, sweeps = 1.0, 0
    for cpt in range(numItermax):
        new_v, col_lse = ops.lse_fused(B, A, cost_id, inv_reg, log_u, log_b, algo_id, want_lse=True)
        if cpt >= 1 and (cpt - 1) % 10 == 0:
            err = float(torch.linalg.vector_norm(torch.exp(log_v + col_lse).double() - b32.double()))
            if not err > stopThr:
                break
        log_v = new_v
        log_u, _ = ops.lse_fused(A, B, cost_id, inv_reg, log_v, log_a, algo_id)
        sweeps = cpt + 1
    P, loss, _ = ops.plan_fused(A, B, cost_id, inv_reg, log_u, log_v, want_plan=return_plan, want_rows=False,
                                algo=algo_id)
    if info is not None:
        info.update(sweeps=sweeps, err=err, log_u=log_u, log_v=log_v)
    return P, loss


class _FusedOTLoss(torch.autograd.Function):
    """<P, C(X, Y)> with P the Sinkhorn plan of the detached cost — the differentiable form of
    models/models_ea.py:218-224 (``sinkhorn(a, b, M.detach(), reg)`` then ``sum(T * M)``) without ever storing
    M or T: forward = the fused tcgen05 solve, backward = one streamed pass per operand (eg_plan_grad_fused)."""

    @staticmethod
    def forward(ctx, X, Y, a, b, reg, numItermax, stopThr, cost, algo):
        info = {}
        _, loss = sinkhorn_fused(X, Y, a, b, reg, numItermax=numItermax, stopThr=stopThr, cost=cost, algo=algo,
                                 return_plan=False, info=info)
        ctx.save_for_backward(X.detach(), Y.detach(), info["log_u"], info["log_v"])
        ctx.reg, ctx.cost = float(reg), cost
        ctx.sweeps = info["sweeps"]
        return loss.to(torch.float32) if X.dtype == torch.float32 else loss

    @staticmethod
    def backward(ctx, gout):
        X, Y, log_u, log_v = ctx.saved_tensors
        cost_id = {"l2": _lib.COST_L2, "sqeuclid": _lib.COST_SQEUCLID, "cos": _lib.COST_COSINE}[ctx.cost]
        A = ops.FusedOperand(X, cost_id, _lib.ALGO_SIMT)
        B = ops.FusedOperand(Y, cost_id, _lib.ALGO_SIMT)
        gout = gout.to(torch.float32)               # stays on the device: no host synchronisation
        dX = ops.plan_grad_fused(A, B, cost_id, 1.0 / ctx.reg, log_u, log_v).mul_(gout) if ctx.needs_input_grad[0] else None
        dY = ops.plan_grad_fused(B, A, cost_id, 1.0 / ctx.reg, log_v, log_u).mul_(gout) if ctx.needs_input_grad[1] else None
        return dX, dY, None, None, None, None, None, None, None


def sinkhorn_fused_loss(X, Y, a, b, reg, numItermax=1000, stopThr=1e-9, cost="l2", algo="tcgen05"):
    """Differentiable OT loss sum_ij T_ij cost(X_i, Y_j), T = sinkhorn(a, b, cost.detach(), reg): what
    ``get_loss_wassertein`` computes before its argmax-of-zeros quirk (models/models_ea.py:218-224), with the cost
    recomputed tile by tile in both passes.  Gradients flow to X and Y through the cost only, as in the reference."""
    return _FusedOTLoss.apply(X, Y, a, b, reg, numItermax, stopThr, cost, algo)
