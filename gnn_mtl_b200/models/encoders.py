"""Graph encoders: the reference's models/encoders.py wiring (:6-94) over the
eagraft layers.  GAT is out of scope."""
import torch.nn as nn

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


class HGCN(Encoder):
    def __init__(self, args):
        super().__init__()
        self._stack(args, lambda i, o, a: HighWayGraphConvolution(i, o, args.dropout, a, args.bias,
                                                                   args.cuda, args.device))


model2encoder = {'GCN': GCN, 'HGCN': HGCN, 'Distill': HGCN, 'MLP': MLP}
