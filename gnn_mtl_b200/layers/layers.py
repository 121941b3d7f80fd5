"""GCN / highway-GCN layers on the eagraft SpMM kernels.

Same class names, constructor signatures, attribute names and state_dict keys as
the reference's layers/layers.py (GraphConvolution :19-42, HighWayGraphConvolution
:45-80, Linear :83-96, get_dim_act :8-16), so models/encoders.py-style glue works
unchanged.  What differs is the execution: the sparse aggregation, activation and
highway blend run as ONE CUDA kernel (eg_spmm), and the backward is a fused
element-wise kernel followed by the same SpMM on the transposed CSR.  The dense
products of a layer (x·Wᵀ + b, x·G + c, dx, dW/db) run on the tcgen05 3xTF32
GEMM kernels (eg_gemm_nt_3xtf32 / eg_gemm_tn_3xtf32); see ``_DenseProducts``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.modules.module import Module

from .. import _lib, ops
from ..adjacency import resolve


def get_dim_act(args):
    """layers/layers.py:8-16: (num_layers-1) layers, dims [feat_dim] + [dim]*(L-1)."""
    act = (lambda x: x) if not args.act else getattr(F, args.act)
    n = args.num_layers - 1
    return [args.feat_dim] + [args.dim] * n, [act] * n


def classify_activation(act):
    """Map the callable the reference passes around to a fused-epilogue code.
    Returns ACT_IDENTITY / ACT_RELU, or None for anything else (un-fused path)."""
    if act is None:
        return _lib.ACT_IDENTITY
    if act in (F.relu, torch.relu):
        return _lib.ACT_RELU
    # Anything else is probed on a wide range (so clamped variants such as relu6 / hardtanh differ from relu)
    # with a fresh clone per call (an in-place activation must not alias the comparison value).
    probe = torch.tensor([-1e4, -100.0, -7.0, -2.0, -0.5, -1e-3, 0.0, 1e-3, 0.75, 3.0, 6.5, 100.0, 1e4])
    try:
        got = act(probe.clone())
    except Exception:
        return None
    if not torch.is_tensor(got) or got.shape != probe.shape:
        return None
    if torch.equal(got, probe):
        return _lib.ACT_IDENTITY
    if torch.equal(got, torch.relu(probe)):
        return _lib.ACT_RELU
    return None


# tests set this to a list to receive act(S) of every fused aggregation (to compare ReLU branch decisions)
CAPTURE_ACT = None


class _Aggregate(torch.autograd.Function):
    """out = epilogue(A · hidden);  backward: dH = Aᵀ · dS  (layers/layers.py:35,64 + autograd)."""

    @staticmethod
    def forward(ctx, hidden, gate_pre, x_res, adjacency, act_code):
        needs_grad = any(ctx.needs_input_grad[:3])
        save_act = needs_grad and (act_code == _lib.ACT_RELU or gate_pre is not None)
        if getattr(adjacency, "sharded", False):     # row-partitioned graph: fetch every rank's feature rows
            hidden = adjacency.gather(hidden)
        out, act_out = ops.spmm(adjacency.csr, hidden, act_code, gate_pre, x_res, save_act=save_act)
        if CAPTURE_ACT is not None:
            CAPTURE_ACT.append(act_out)
        ctx.adjacency, ctx.act_code = adjacency, act_code
        ctx.has_gate = gate_pre is not None
        if ctx.has_gate:
            ctx.save_for_backward(act_out, gate_pre, x_res)
        else:
            ctx.save_for_backward(act_out)
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.has_gate:
            act_out, gate_pre, x_res = ctx.saved_tensors
        else:
            (act_out,) = ctx.saved_tensors
            gate_pre = x_res = None
        need_h, need_g, need_x = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        d_gate = d_xres = None
        if ctx.has_gate or ctx.act_code == _lib.ACT_RELU:
            dS, d_gate, d_xres = ops.epilogue_bwd(dout, act_out, gate_pre, x_res, ctx.act_code, need_g, need_x)
        else:
            dS = dout
        dH = None
        if need_h:
            if getattr(ctx.adjacency, "sharded", False):
                dS = ctx.adjacency.gather(dS, transposed=True)
            dH, _ = ops.spmm(ctx.adjacency.csr_t, dS, _lib.ACT_IDENTITY)
        return dH, d_gate, d_xres, None, None


USE_TCGEN05_GEMM = True   # dense layer products on the 3xTF32 tcgen05 tiles; False -> everything on cuBLAS fp32
USE_TCGEN05_DW = True     # dW = dHᵀ·x on the MN-major split-K tcgen05 kernel; False -> cuBLAS fp32 batched split-K
USE_CHAINED_GEMM = True   # ReLU-feeding x·Wᵀ + b on the short-chain tcgen05 kernel; False -> cuBLAS fp32 (SIMT)


def _weight_grad(d_hidden, x):
    """dW = dHᵀ·x ([out, n]·[n, in]) on cuBLAS fp32.  The plain call tiles only the 300x300 output
    (15 CTAs on 148 SMs); batching the reduction dimension gives cuBLAS a split-K it does not pick itself."""
    n = x.shape[0]
    x = x.contiguous()
    for parts in (64, 32, 16):
        if n % parts == 0 and n // parts >= 1024:
            return torch.bmm(d_hidden.view(parts, n // parts, -1).transpose(1, 2), x.view(parts, n // parts, -1)).sum(0)
    return d_hidden.t() @ x


class _DenseProducts(torch.autograd.Function):
    """hidden = x·Wᵀ + b  and (highway) gate_pre = x·G + c  (layers/layers.py:61,69).

    3xTF32 on tcgen05 is accurate to ~2e-6 relative, cuBLAS fp32 to ~2e-7.  That is irrelevant for
    the smooth consumers (sigmoid gate, identity activation, the linear input gradient) but not
    for a product that feeds a ReLU: the activation acts on S = A·hidden, and a relative error of 2e-6 in
    hidden moves ~1e-5 of the near-zero entries of S across zero (emulated on the benchmark graph: ~90 of 6e7 per layer,
    against ~4 for an error of fp32 size — profiles/README.md); each flip moves gradient entries of
    that row by ~5e-3 of the max-norm — outside the 1e-4 parity bar.  (A sign-safe epilogue on `hidden` itself
    was tried in round 2 and is the wrong place: the discontinuity sits behind the aggregation.)  So `hidden` stays on cuBLAS fp32 when the layer's activation is ReLU
    (`exact_hidden`), and goes through the tensor-core kernel otherwise; `gate_pre` and
    dx = [dH | d_gate]·[Wᵀ | G]ᵀ always do.  dW = dHᵀ·x runs on the MN-major split-K tensor-core kernel from the
    same hi/lo splits (x's is kept from the forward pass, dH's is shared with dx); db is a reduction."""

    @staticmethod
    def forward(ctx, x, weight, bias, gate_w, gate_b, exact_hidden):
        n_out = weight.shape[0]
        gate_pre = None
        x_split = None                      # hi/lo split of x: made once, reused by dW = dHᵀ·x in backward
        if exact_hidden and USE_CHAINED_GEMM:
            # short accumulation chains (6 MMAs, folded in fp32 registers): fp32-SIMT-level error, so the ReLU
            # behind the aggregation sees the same branches as an fp32 product
            # (the gate is a smooth consumer: it takes the long-chain kernel, 0.255 ms against 0.327 ms at 300 columns)
            hidden, sp = ops.gemm_nt([x], weight, bias, return_splits=True, chained=True)
            x_split = sp[0]
            if gate_w is not None:
                # raw-operand CTA-pair kernel: 0.211 ms against 0.251 ms for the split-operand kernel fed with
                # x_split (200k x 300 -> 300); same bits
                gate_pre = ops.gemm_nt_raw([x], gate_w.t().contiguous(), gate_b)
        elif exact_hidden:
            # cuBLAS fp32; the bias is added in place afterwards (cublasLt's own bias pass for this shape is a
            # separate 0.34 ms kernel, the in-place add 0.16 ms; same roundings: fl(fl(x·Wᵀ) + b))
            hidden = torch.mm(x, weight.t())
            if bias is not None:
                hidden.add_(bias)
            if gate_w is not None:
                gate_pre, sp = ops.gemm_nt([x], gate_w.t().contiguous(), gate_b, return_splits=True)
                x_split = sp[0]
        elif gate_w is None:
            hidden, sp = ops.gemm_nt([x], weight, bias, return_splits=True)
            x_split = sp[0]
        else:
            b0 = bias if bias is not None else torch.zeros(n_out, device=x.device)
            (hidden, gate_pre), sp = ops.gemm_nt([x], torch.cat([weight, gate_w.t()], 0), torch.cat([b0, gate_b]),
                                                 n1=n_out, return_splits=True)
            x_split = sp[0]
        keep_split = USE_TCGEN05_DW and x_split is not None and ctx.needs_input_grad[1]
        # db = dHᵀ·1 rides along with dW = dHᵀ·x: the first padding column of x's hi part is set to one (the
        # forward products above are done with it; their weight operand is zero over the padding)
        ctx.ones_col = bool(keep_split and bias is not None and x_split[0].shape[1] > x.shape[1])
        if ctx.ones_col:
            x_split[0][:, x.shape[1]] = 1.0
        empty = weight.new_empty(0)
        ctx.save_for_backward(x, weight, gate_w if gate_w is not None else empty,
                              x_split[0] if keep_split else empty, x_split[1] if keep_split else empty)
        ctx.has_gate = gate_w is not None
        ctx.has_bias = bias is not None
        return hidden, gate_pre

    @staticmethod
    def backward(ctx, d_hidden, d_gate):
        x, weight, gate_w, x_hi, x_lo = ctx.saved_tensors
        dx = dW = db = None
        if d_hidden is None:
            d_hidden = torch.zeros(x.shape[0], weight.shape[0], device=x.device)
        d_hidden = d_hidden.contiguous()
        use_tc_dw = USE_TCGEN05_DW and ctx.needs_input_grad[1] and weight.shape[0] % 4 == 0 and x.shape[1] % 4 == 0
        dh_split = (ops.split_tf32(d_hidden, ops._pad16(d_hidden.shape[1]))
                    if (use_tc_dw or (ctx.needs_input_grad[0] and not (ctx.has_gate and d_gate is not None))) else None)
        if ctx.needs_input_grad[0]:
            if ctx.has_gate and d_gate is not None:
                # raw-operand kernel (hi/lo split inside the kernel): d_gate is used here only, so its split pass
                # (a launch + 0.7 GB of HBM traffic) disappears; same bits as the split-operand product
                dx = ops.gemm_nt_raw([d_hidden, d_gate.contiguous()], torch.cat([weight.t(), gate_w], 1))
            else:
                dx = ops.gemm_nt([d_hidden], weight.t().contiguous(), a_splits=[dh_split])
        if ctx.needs_input_grad[1]:
            if use_tc_dw:
                if x_hi.numel() == 0:
                    x_hi, x_lo = ops.split_tf32(x.contiguous(), ops._pad16(x.shape[1]))
                if ctx.ones_col and ctx.needs_input_grad[2]:
                    both = ops.gemm_tn(dh_split, weight.shape[0], (x_hi, x_lo), x.shape[1] + 1)
                    dW, db = both[:, :x.shape[1]].contiguous(), both[:, x.shape[1]].contiguous()
                else:
                    dW = ops.gemm_tn(dh_split, weight.shape[0], (x_hi, x_lo), x.shape[1])
            else:
                dW = _weight_grad(d_hidden, x)
        if db is None and ctx.has_bias and ctx.needs_input_grad[2]:
            db = d_hidden.sum(0)
        return dx, dW, db, None, None, None


def _tc_gemm_ok(x, linear, gate_w=None):
    # in_features % 4: the backward dx product writes rows of in_features floats through the same kernel
    return (USE_TCGEN05_GEMM and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2
            and linear.out_features % 4 == 0 and x.shape[1] % 4 == 0 and x.shape[0] > 0
            and (gate_w is None or gate_w.shape[1] % 4 == 0))


def _dense(x):
    # the reference stores the (fully dense) feature matrix as sparse COO (utils/data_utils.py:358,397)
    return x.to_dense() if x.is_sparse else x


class GraphConvolution(Module):
    """act(A · dropout(x Wᵀ + b)).  layers/layers.py:19-42."""

    def __init__(self, in_features, out_features, dropout, act, use_bias):
        super(GraphConvolution, self).__init__()
        self.dropout = dropout
        self.linear = nn.Linear(in_features, out_features, use_bias)
        self.act = act
        self.in_features = in_features
        self.out_features = out_features
        self._act_code = classify_activation(act)

    def _aggregate(self, hidden, adj, gate_pre=None, x_res=None):
        if torch.is_tensor(adj) and adj.layout == torch.strided:
            # dense adjacency branch of the reference (:36-37); not the EA path
            support = self.act(torch.mm(adj, hidden))
            if gate_pre is None:
                return support
            t = torch.sigmoid(gate_pre)
            return t * support + (1.0 - t) * x_res
        adjacency = resolve(adj)
        if self._act_code is None:
            # activation the epilogue does not know: aggregate fused, activate/blend outside
            support = self.act(_Aggregate.apply(hidden, None, None, adjacency, _lib.ACT_IDENTITY))
            if gate_pre is None:
                return support
            t = torch.sigmoid(gate_pre)
            return t * support + (1.0 - t) * x_res
        return _Aggregate.apply(hidden, gate_pre, x_res, adjacency, self._act_code)

    def forward(self, input):
        x, adj = input
        x = _dense(x)
        if _tc_gemm_ok(x, self.linear):
            hidden, _ = _DenseProducts.apply(x, self.linear.weight, self.linear.bias, None, None,
                                             self._act_code != _lib.ACT_IDENTITY)
        else:
            hidden = self.linear.forward(x)
        hidden = F.dropout(hidden, self.dropout, training=self.training)
        return self._aggregate(hidden, adj), adj

    def extra_repr(self):
        return 'input_dim={}, output_dim={}'.format(self.in_features, self.out_features)


class HighWayGraphConvolution(GraphConvolution):
    """t·act(A·(xWᵀ+b)) + (1-t)·x with t = sigmoid(x G + c).  layers/layers.py:45-80.

    kernel_gate / bias_gate are plain tensors as in the reference (not
    Parameters, not in state_dict, moved by hand): initialised after nn.Linear
    from the global RNG with U(±sqrt(6/2d)) / zeros (:51-54)."""

    def __init__(self, in_features, out_features, dropout, act, use_bias, cuda, device):
        super(HighWayGraphConvolution, self).__init__(in_features, out_features, dropout, act, use_bias)
        assert (self.in_features == self.out_features)
        d = self.in_features
        init_range = np.sqrt(6.0 / (d + d))
        self.kernel_gate = torch.FloatTensor(d, d).uniform_(-init_range, init_range)
        self.bias_gate = torch.zeros([d])
        if not cuda == -1:
            self.kernel_gate = self.kernel_gate.to(device)
            self.bias_gate = self.bias_gate.to(device)

    def forward(self, input):
        x, adj = input
        x = _dense(x)
        if self.kernel_gate.device != x.device:      # reference moves them in __init__ only
            self.kernel_gate = self.kernel_gate.to(x.device)
            self.bias_gate = self.bias_gate.to(x.device)
        if _tc_gemm_ok(x, self.linear, self.kernel_gate):
            hidden, gate_pre = _DenseProducts.apply(x, self.linear.weight, self.linear.bias, self.kernel_gate,
                                                    self.bias_gate, self._act_code != _lib.ACT_IDENTITY)
        else:
            hidden = self.linear.forward(x)
            gate_pre = torch.addmm(self.bias_gate, x, self.kernel_gate)
        hidden = F.dropout(hidden, self.dropout, training=self.training)
        return self._aggregate(hidden, adj, gate_pre, x), adj


class Linear(Module):
    """act(dropout(x Wᵀ + b)).  layers/layers.py:83-96 (dense; not a changed subsystem)."""

    def __init__(self, in_features, out_features, dropout, act, use_bias):
        super(Linear, self).__init__()
        self.dropout = dropout
        self.linear = nn.Linear(in_features, out_features, use_bias)
        self.act = act

    def forward(self, x):
        hidden = self.linear.forward(x)
        hidden = F.dropout(hidden, self.dropout, training=self.training)
        return self.act(hidden)
