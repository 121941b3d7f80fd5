"""Entity-alignment models: the callers of the hot path in models/models_ea.py
(BaseModel.get_neg :19-30, compute_metrics :59-66, EAModel.get_loss :103-123,
UEAModel.generate_pairs :143-167, generate_neg :169-183, get_loss_wassertein
:206-224), keeping names, signatures and the reference's observable quirks.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..utils.eval_utils import get_hits
from ..utils.ot_loss import sinkhorn
from .decoders import model2decoder
from .encoders import model2encoder


# The Sinkhorn solve of get_loss_wassertein feeds nothing downstream (the plan is computed and then not used,
# models_ea.py:221-222), so nothing in the backward pass or the optimizer step waits for it: it runs on a side stream
# next to them and is joined at the end of the step (join_pending_solve).  The solve is latency-bound and leaves most
# issue slots and all of the HBM bandwidth idle; the streaming kernels of the backward pass fill them.
# EG_SINKHORN_OVERLAP=0 keeps everything on the caller's stream.
OVERLAP_SINKHORN = os.environ.get("EG_SINKHORN_OVERLAP", "1") != "0"
_SIDE_STREAMS = {}


def _side_stream(dev):
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=key)
    return _SIDE_STREAMS[key]


# The as-shipped objective is sum_i M[i, 0] with M = cdist(X, Y) (models_ea.py:218-224: the one-hot comes from the argmax
# of a zero tensor).  Its gradient is known in closed form — d M[i,0] / d X_i = (X_i - Y_0) / M[i,0], and Y_0 gets minus
# the sum — so the backward pass does not have to push a one-hot [bsz, bsz] gradient through cdist's matmul
# formulation (two 3000 x 3000 x 300 products and a dozen element-wise passes for 3000 non-zero entries).  The value is
# still read from the same cdist matrix the Sinkhorn solve gets.  EG_COL0_GRAD=0 restores autograd through cdist.
CLOSED_FORM_COL0_GRAD = os.environ.get("EG_COL0_GRAD", "1") != "0"


class _Column0Loss(torch.autograd.Function):
    """sum_i M[i, 0] as a function of X, Y with m0 = M[:, 0] = ||X_i - Y_0||_2 given (cdist's own values)."""

    @staticmethod
    def forward(ctx, X, Y, m0):
        ctx.save_for_backward(X, Y[0], m0)
        ctx.y_shape = Y.shape
        return torch.sum(m0.to(torch.float64))

    @staticmethod
    def backward(ctx, g):
        X, y0, m0 = ctx.saved_tensors
        # cdist's backward divides by the distance too (and gives 0 at an exact zero distance)
        w = torch.where(m0 > 0, g.to(m0.dtype) / m0, torch.zeros_like(m0)).unsqueeze(1)
        dX = (X - y0) * w
        dY = torch.zeros(ctx.y_shape, dtype=X.dtype, device=X.device)
        dY[0] = -dX.sum(0)
        return dX, dY, None


class BaseModel(nn.Module):
    def __init__(self, args):
        super(BaseModel, self).__init__()
        self.n_nodes = args.n_nodes
        self.device = args.device

    def get_neg(self, ILL, output, k):
        """k nearest (L1, fp64) entities per anchor, rank 0 dropped; flat int64 [t*k]."""
        out = output.detach().to(torch.float32)
        anchors = torch.as_tensor(np.asarray(ILL, dtype=np.int64), device=out.device)
        idx = ops.l1_topk(out.index_select(0, anchors), out, 1, k)
        return idx.reshape(-1).cpu().numpy()

    def compute_metrics(self, outputs, data, split):
        pair = data['train'] if split == 'train' else data['test']
        return get_hits(outputs, pair, top_k=[1])

    def has_improved(self, m1, m2):
        return (m1['Hits@10_l'] < m2['Hits@10_l']) or (m1['Hits@10_r'] < m2['Hits@10_r'])

    def init_metric_dict(self):
        return {'Hits@1_l': -1, 'Hits@10_l': -1, 'Hits@50_l': -1, 'Hits@100_l': -1,
                'Hits@1_r': -1, 'Hits@10_r': -1, 'Hits@50_r': -1, 'Hits@100_r': -1}


def _margin_loss(outputs, ILL, neg_left, neg_right, neg2_left, neg2_right, k):
    """Margin-based L1 ranking loss with hard negatives (:103-123 / :185-204) on the fused
    gather + L1 + hinge kernel (eg_margin_loss_fwd/bwd); the reference's index arrays are float
    NumPy arrays of integral values — they are cast to int64 here."""
    ILL = np.asarray(ILL)
    return ops.margin_loss(outputs, np.asarray(ILL[:, 0], dtype=np.int64), np.asarray(ILL[:, 1], dtype=np.int64),
                           np.asarray(neg_left, dtype=np.int64), np.asarray(neg_right, dtype=np.int64),
                           np.asarray(neg2_left, dtype=np.int64), np.asarray(neg2_right, dtype=np.int64), k, 1.0)


class EAModel(BaseModel):
    def __init__(self, args):
        super(EAModel, self).__init__(args)
        self.encoder = model2encoder[args.model](args)
        self.decoder = model2decoder[args.model](args)
        ILL = args.data['train']
        t, k = len(ILL), args.neg_num
        self.neg_num = k
        self.neg_left = (np.ones((t, k)) * (ILL[:, 0].reshape((t, 1)))).reshape((t * k,))
        self.neg2_right = (np.ones((t, k)) * (ILL[:, 1].reshape((t, 1)))).reshape((t * k,))
        self.neg_right = None
        self.neg2_left = None

    def encode(self, x, adj):
        return self.encoder.encode(x, adj)

    def decode(self, h, adj):
        return self.decoder.decode(h, adj)

    def get_loss(self, outputs, data, split):
        return _margin_loss(outputs, data[split], self.neg_left, self.neg_right, self.neg2_left,
                            self.neg2_right, self.neg_num)


class UEAModel(BaseModel):
    def __init__(self, args):
        super(UEAModel, self).__init__(args)
        self.ILL = None
        self.encoder = model2encoder[args.model](args)
        self.decoder = model2decoder[args.model](args)

    def encode(self, x, adj):
        return self.encoder.encode(x, adj)

    def decode(self, h, adj):
        return self.decoder.decode(h, adj)

    def generate_pairs(self, outputs, data, bsz):
        """Mutual nearest neighbours under fp64 L1, closest first, at most bsz
        (:143-167; positions are local, as in the reference)."""
        e1, e2 = data['e1'], data['e2']
        index1, index2 = data['index1'], data['index2']
        out = outputs.detach().to(torch.float32)
        L = torch.as_tensor([index1[i] for i in range(e1)], device=out.device)
        R = torch.as_tensor([index2[i] for i in range(e2)], device=out.device)
        row_min, row_arg, _, col_arg = ops.l1_argmins(out.index_select(0, L), out.index_select(0, R))
        mutual = col_arg[row_arg] == torch.arange(e1, device=out.device)
        keep = torch.nonzero(mutual).reshape(-1)
        pairs = torch.stack([keep, row_arg[keep]], 1)
        print("generate {} pairs by the L1 distance".format(min(len(pairs), bsz)))
        order = torch.argsort(row_min[keep], stable=True)[:bsz]
        self.ILL = pairs[order].cpu().numpy()
        return

    def generate_neg(self, outputs, k):
        t = len(self.ILL)
        self.neg_num = k
        self.neg_left = (np.ones((t, k)) * (self.ILL[:, 0].reshape((t, 1)))).reshape((t * k,))
        self.neg2_right = (np.ones((t, k)) * (self.ILL[:, 1].reshape((t, 1)))).reshape((t * k,))
        self.neg_right = self.get_neg(self.ILL[:, 0], outputs, k)
        self.neg2_left = self.get_neg(self.ILL[:, 1], outputs, k)
        return

    def get_loss(self, outputs):
        return _margin_loss(outputs, self.ILL, self.neg_left, self.neg_right, self.neg2_left,
                            self.neg2_right, self.neg_num)

    def get_loss_wassertein(self, outputs, data, bsz, *, numItermax=1000, stopThr=1e-9, sample=None):
        """:206-224, quirk included: the Sinkhorn plan is computed and then not
        used — the one-hot is taken from argmax of a zero tensor, i.e. column 0 —
        so the value (and its gradient) is sum_i ||X_i - Y_0||_2.

        ``sample`` (keyword-only extension): a pair of int64 CUDA index tensors to
        use instead of drawing np.random.permutation on the host (:211-212)."""
        dev = outputs.device
        if sample is None:
            e1, e2 = data['e1'], data['e2']
            index1, index2 = data['index1'], data['index2']
            L = _take(index1, np.random.permutation(e1)[:bsz])
            R = _take(index2, np.random.permutation(e2)[:bsz])
            sample = (_host_to_device(L, dev), _host_to_device(R, dev))
        X = outputs[sample[0]]
        Y = outputs[sample[1]]
        a, b = torch.ones(bsz, device=dev), torch.ones(bsz, device=dev)
        closed_form = CLOSED_FORM_COL0_GRAD and X.is_cuda
        if closed_form:
            with torch.no_grad():
                M = torch.cdist(X, Y, p=2)
        else:
            M = torch.cdist(X, Y, p=2)
        self.join_pending_solve()
        if OVERLAP_SINKHORN and M.is_cuda and stopThr < 0:
            # (with a stop rule the solver reads the marginal error back on the host: nothing to overlap)
            main, side = torch.cuda.current_stream(dev), _side_stream(dev)
            Md = M.detach()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                sinkhorn(a, b, Md, reg=0.01, numItermax=numItermax, stopThr=stopThr, return_plan=False)
            for t in (Md, a, b):
                t.record_stream(side)
            self._pending_solve = side
        else:
            T, _ = sinkhorn(a, b, M.detach(), reg=0.01, numItermax=numItermax, stopThr=stopThr, return_plan=False)
        # newT = one-hot(argmax(zeros)) = column 0 of every row (reference :221-222)
        if closed_form:
            return _Column0Loss.apply(X, Y, M[:, 0].contiguous())
        return torch.sum(M[:, 0].to(torch.float64))

    def join_pending_solve(self):
        """Make the caller's stream wait for a Sinkhorn solve still running on the side stream (end of a step)."""
        side = getattr(self, "_pending_solve", None)
        if side is not None:
            torch.cuda.current_stream(side.device).wait_stream(side)
            self._pending_solve = None


def _gw_loss(self, outputs, data, bsz, *, max_iter=1000, epsilon=0.01, sample=None):
    """models/models_ea.py:226-244 (call sites are commented out upstream, run/train_unsup_ea.py:100): L1 cost,
    Gromov-Wasserstein plan by iterative projection, then — same quirk as the Wasserstein loss — the one-hot is
    taken from argmax of a zero tensor, so the value is sum_i ||X_i - Y_0||_1."""
    from ..SinkhornOT import gw_iterative_1
    dev = outputs.device
    if sample is None:
        e1, e2 = data['e1'], data['e2']
        index1, index2 = data['index1'], data['index2']
        L = _take(index1, np.random.permutation(e1)[:bsz])
        R = _take(index2, np.random.permutation(e2)[:bsz])
        sample = (_host_to_device(L, dev), _host_to_device(R, dev))
    X, Y = outputs[sample[0]], outputs[sample[1]]
    a, b = torch.ones(bsz, device=dev), torch.ones(bsz, device=dev)
    M = torch.cdist(X, Y, p=1)
    C1 = torch.cdist(X, X, p=1).detach()
    C2 = torch.cdist(Y, Y, p=1).detach()
    gw_iterative_1(C1, C2, a, b, epsilon=epsilon, max_iter=max_iter)
    return torch.sum(M[:, 0].to(torch.float64))


UEAModel.get_loss_gromove_wassertein = _gw_loss


def _take(index, positions):
    """np.array([index[i] for i in positions]) of models_ea.py:211-212 — as one fancy-indexing call when `index` is an
    array (same values, same RNG consumption by the caller; ~1 ms of interpreter loop per side at bsz = 3000 otherwise)."""
    if isinstance(index, np.ndarray):
        return index[positions]
    return np.array([index[i] for i in positions])


def _host_to_device(arr, dev):
    """Host index array -> CUDA tensor through pinned memory (async H2D)."""
    staged = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64)).pin_memory()
    return staged.to(dev, non_blocking=True)
