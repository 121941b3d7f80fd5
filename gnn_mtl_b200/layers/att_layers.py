"""Attention layers: the reference's layers/att_layers.py (:8-93) over the eagraft GAT kernels.

``SpGraphAttentionLayer`` keeps the reference's parameters (``W`` [in, out], ``a`` [1, 2·out], same
initialisation) and forward semantics; the per-edge gather / exp / two SpMMs / divide of :39-59 run as one
gather kernel (``eg_gat_fwd``) that never materialises the edge weights, with an explicit backward
(``eg_gat_bwd_edges`` + the transposed SpMM).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..adjacency import resolve


class SpGraphAttentionLayer(nn.Module):
    """Sparse GAT layer (layers/att_layers.py:8-64)."""

    def __init__(self, in_features, out_features, dropout, alpha, activation):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.alpha = alpha
        self.W = nn.Parameter(torch.zeros(size=(in_features, out_features)))
        nn.init.xavier_normal_(self.W.data, gain=1.414)
        self.a = nn.Parameter(torch.zeros(size=(1, 2 * out_features)))
        nn.init.xavier_normal_(self.a.data, gain=1.414)
        self.dropout = nn.Dropout(dropout)
        self.leakyrelu = nn.LeakyReLU(self.alpha)
        self.act = activation

    def forward(self, input, adj):
        adjacency = resolve(adj)
        h = torch.mm(input, self.W)                                  # :33
        # a·[h_i ‖ h_j] = h_i·a[:D] + h_j·a[D:]  (:38-41), so two length-N projections replace the 2D×E gather
        D = self.out_features
        s1 = torch.mv(h, self.a[0, :D])
        s2 = torch.mv(h, self.a[0, D:])
        edge_scale = None
        if self.training and self.dropout.p > 0:                     # edge dropout after the row sum (:50)
            edge_scale = self.dropout(torch.ones(adjacency.csr.nnz, device=h.device))
        h_prime = ops.gat_aggregate(h, s1, s2, adjacency, self.alpha, edge_scale)
        return self.act(h_prime)

    def __repr__(self):
        return self.__class__.__name__ + ' (' + str(self.in_features) + ' -> ' + str(self.out_features) + ')'


class GraphAttentionLayer(nn.Module):
    """Multi-head wrapper (layers/att_layers.py:67-93): concat or mean over heads."""

    def __init__(self, input_dim, output_dim, dropout, activation, alpha, nheads, concat):
        super().__init__()
        self.dropout = dropout
        self.output_dim = output_dim
        self.attentions = [SpGraphAttentionLayer(input_dim, output_dim, dropout=dropout, alpha=alpha,
                                                 activation=activation) for _ in range(nheads)]
        self.concat = concat
        for i, attention in enumerate(self.attentions):
            self.add_module('attention_{}'.format(i), attention)

    def forward(self, input):
        x, adj = input
        x = F.dropout(x, self.dropout, training=self.training)
        if self.concat:
            h = torch.cat([att(x, adj) for att in self.attentions], dim=1)
        else:
            h_cat = torch.cat([att(x, adj).view((-1, self.output_dim, 1)) for att in self.attentions], dim=2)
            h = torch.mean(h_cat, dim=2)
        h = F.dropout(h, self.dropout, training=self.training)
        return (h, adj)
