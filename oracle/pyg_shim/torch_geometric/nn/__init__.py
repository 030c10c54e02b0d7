"""TEST INFRASTRUCTURE ONLY - GCNConv and GraphNorm with PyG 2.3.1 default semantics.

Call sites in the reference: TwoWL/model/model.py:37-38,54. Restated from the PyG 2.3.1
documentation of ``GCNConv`` / ``gcn_norm`` / ``add_remaining_self_loops`` / ``GraphNorm``
(defaults: add_self_loops=True, normalize=True, bias=True, cached=False, improved=False;
GraphNorm eps=1e-5, batch=None). Parameter names match PyG so state_dicts interchange.
"""
import math

import torch
from torch import nn


class _PygLinear(nn.Module):
    """``torch_geometric.nn.dense.linear.Linear(bias=False, weight_initializer='glorot')``."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        a = math.sqrt(6.0 / (in_channels + out_channels))
        nn.init.uniform_(self.weight, -a, a)

    def forward(self, x):
        return x @ self.weight.t()


def gcn_norm(edge_index, num_nodes, dtype=torch.float32):
    """Drop (i,i) columns, append N self-loops of weight 1, symmetric-normalise by the
    in-degree (histogram of row 1)."""
    keep = edge_index[0] != edge_index[1]
    loop = torch.arange(num_nodes, dtype=torch.long, device=edge_index.device)
    ei = torch.cat([edge_index[:, keep], loop.unsqueeze(0).repeat(2, 1)], dim=1)
    w = torch.ones(ei.size(1), dtype=dtype, device=ei.device)
    deg = torch.zeros(num_nodes, dtype=dtype, device=ei.device).scatter_add_(0, ei[1], w)
    dis = deg.pow(-0.5)
    dis = dis.masked_fill(dis == float("inf"), 0.0)
    return ei, dis[ei[0]] * w * dis[ei[1]]


class GCNConv(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _PygLinear(in_channels, out_channels)
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, edge_index):
        ei, w = gcn_norm(edge_index, x.size(0), x.dtype)
        z = self.lin(x)
        msg = w.view(-1, 1) * z.index_select(0, ei[0])
        out = torch.zeros_like(z).index_add_(0, ei[1], msg)
        return out + self.bias


class GraphNorm(nn.Module):
    def __init__(self, in_channels, eps=1e-5):
        super().__init__()
        self.in_channels, self.eps = in_channels, eps
        self.weight = nn.Parameter(torch.ones(in_channels))
        self.bias = nn.Parameter(torch.zeros(in_channels))
        self.mean_scale = nn.Parameter(torch.ones(in_channels))

    def forward(self, x, batch=None):
        assert batch is None, "the TwoWL path never passes a batch vector"
        mean = x.mean(dim=0, keepdim=True)
        out = x - mean * self.mean_scale
        var = out.pow(2).mean(dim=0, keepdim=True)
        std = (var + self.eps).sqrt()
        return self.weight * out / std + self.bias
