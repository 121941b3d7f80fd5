"""Graph decoders: models/decoders.py wiring (:6-81) over the eagraft layers."""
import torch.nn as nn
import torch.nn.functional as F

from ..layers.att_layers import GraphAttentionLayer
from ..layers.layers import GraphConvolution, HighWayGraphConvolution, Linear


def _identity(x):
    return x


class Decoder(nn.Module):
    """decode(x, adj) (models/decoders.py:12-18)."""

    decode_adj = True

    def decode(self, x, adj):
        if self.decode_adj:
            probs, _ = self.cls.forward((x, adj))
            return probs
        return self.cls.forward(x)


class GCNDecoder(Decoder):
    def __init__(self, args):
        super().__init__()
        self.cls = GraphConvolution(args.dim, args.n_classes, args.dropout, _identity, args.bias)


class GATDecoder(Decoder):
    """One single-head ELU attention layer (models/decoders.py:31-37)."""

    def __init__(self, args):
        super().__init__()
        self.cls = GraphAttentionLayer(args.dim, args.n_classes, args.dropout, F.elu, args.alpha, 1, True)


class HGCNDecoder(Decoder):
    """One identity-activation highway layer (models/decoders.py:40-47)."""

    def __init__(self, args):
        super().__init__()
        self.cls = HighWayGraphConvolution(args.dim, args.n_classes, args.dropout, _identity, args.bias,
                                           args.cuda, args.device)


class MLPDecoder(Decoder):
    """dim -> dim -> dim -> n_classes, relu / relu / identity (models/decoders.py:50-63)."""

    decode_adj = False

    def __init__(self, args):
        super().__init__()
        dims = [args.dim, args.dim, args.dim, args.n_classes]
        acts = [F.relu, F.relu, _identity]
        self.cls = nn.Sequential(*[Linear(dims[i], dims[i + 1], args.dropout, acts[i], args.bias)
                                   for i in range(3)])


model2decoder = {'GCN': MLPDecoder, 'GAT': MLPDecoder, 'HGCN': HGCNDecoder, 'Distill': HGCNDecoder}
