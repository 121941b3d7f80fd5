"""Graph encoders: the reference's models/encoders.py wiring (:6-94) over the
eagraft layers."""
import torch.nn as nn

from ..layers.att_layers import GraphAttentionLayer
from ..layers.layers import GraphConvolution, HighWayGraphConvolution, Linear, get_dim_act


class Encoder(nn.Module):
    """encode(x, adj): thread (x, adj) through nn.Sequential (models/encoders.py:12-18)."""

    encode_graph = True

    def encode(self, x, adj):
        if self.encode_graph:
            output, _ = self.layers.forward((x, adj))
            return output
        return self.layers.forward(x)

    def _stack(self, args, make):
        assert args.num_layers > 0
        dims, acts = get_dim_act(args)
        self.layers = nn.Sequential(*[make(dims[i], dims[i + 1], acts[i]) for i in range(len(dims) - 1)])


class MLP(Encoder):
    encode_graph = False

    def __init__(self, args):
        super().__init__()
        self._stack(args, lambda i, o, a: Linear(i, o, args.dropout, a, args.bias))


class GCN(Encoder):
    def __init__(self, args):
        super().__init__()
        self._stack(args, lambda i, o, a: GraphConvolution(i, o, args.dropout, a, args.bias))


class GAT(Encoder):
    """Graph attention encoder (models/encoders.py:69-86): n_heads heads of width dim / n_heads, concatenated."""

    def __init__(self, args):
        super().__init__()
        assert args.num_layers > 0
        dims, acts = get_dim_act(args)
        gat_layers = []
        for i in range(len(dims) - 1):
            assert dims[i + 1] % args.n_heads == 0
            gat_layers.append(GraphAttentionLayer(dims[i], dims[i + 1] // args.n_heads, args.dropout, acts[i],
                                                  args.alpha, args.n_heads, True))
        self.layers = nn.Sequential(*gat_layers)


class HGCN(Encoder):
    def __init__(self, args):
        super().__init__()
        self._stack(args, lambda i, o, a: HighWayGraphConvolution(i, o, args.dropout, a, args.bias,
                                                                   args.cuda, args.device))


model2encoder = {'GCN': GCN, 'GAT': GAT, 'HGCN': HGCN, 'Distill': HGCN, 'MLP': MLP}
