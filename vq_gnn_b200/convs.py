"""`OurGCNConv` / `OurGATConv` — drop-ins for vq_gnn_v{1,2}/convs.py:26-101, 124-266.

Parameter names, shapes, initialisers and RNG consumption follow the reference (which inherits them
from PyG 1.7.2 `GCNConv` / `GATConv`) so `state_dict()` keys match.  As in the reference, the GCN
conv's `weight` / `bias` and the v2 GAT conv's `lin_l` exist but are never used by `forward`
(convs.py:69-99 are commented out).

`forward(x, adj)` computes the plain (non-VQ) operator on an explicit CSR adjacency through the same
CUDA kernels the VQ layers use (a plan with no tail entries):
    GCN/SAGE:  out = adj @ x                                         (convs.py:65-101)
    GAT:       out[i] = sum_j adj[i,j] exp(lrelu((a_l[j]+a_r[i])/s)) x[j]   (convs.py:165-266,
               vq_softmax.py:41-57: un-normalised exp, no max-subtraction)
The VQ layers do not call these `forward`s: they fuse the codeword gather into the same kernels
(see models.py: `VQConvFunction`).
"""
from __future__ import annotations

import math

import torch
from torch import nn
from torch.nn import Linear, Parameter

Tensor = torch.Tensor


def glorot(t: Tensor):
    stdv = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    t.data.uniform_(-stdv, stdv)


class OurGCNConv(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, improved: bool = False, cached: bool = False,
                 add_self_loops: bool = True, normalize: bool = True, bias: bool = True, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached = improved, cached
        self.add_self_loops, self.normalize = add_self_loops, normalize
        self.weight = Parameter(torch.empty(in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        glorot(self.weight)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x: Tensor, edge_index, edge_weight=None) -> Tensor:
        from .models import plain_propagate
        return plain_propagate(x, edge_index, None, None)


class OurGATConv(nn.Module):
    def __init__(self, in_channels, out_channels: int, heads: int = 1, concat: bool = True,
                 negative_slope: float = 0.2, dropout: float = 0.0, add_self_loops: bool = True,
                 bias: bool = True, version: str = 'v2', **kwargs):
        super().__init__()
        if heads != 1:
            raise NotImplementedError("the VQ-GNN layers use a single head")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.add_self_loops = add_self_loops
        self.version = version
        if version == 'v2':   # v2/convs.py:141-143 creates lin_l (unused); v1/convs.py:160-165 does not
            self.lin_l = Linear(in_channels, heads * out_channels, bias=False)
            self.lin_r = self.lin_l
        self.att_l = Parameter(torch.empty(1, heads, out_channels))
        self.att_r = Parameter(torch.empty(1, heads, out_channels))
        if bias and concat:
            self.bias = Parameter(torch.empty(heads * out_channels))
        elif bias and not concat:
            self.bias = Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.version == 'v2':      # inherited GATConv.reset_parameters (PyG 1.7.2)
            glorot(self.lin_l.weight)
            glorot(self.lin_r.weight)
        glorot(self.att_l)
        glorot(self.att_r)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x: Tensor, edge_index, size=None, return_attention_weights=None) -> Tensor:
        if isinstance(edge_index, torch.Tensor):
            raise NotImplementedError  # vq_softmax.py:54-55: only the CSR (SparseTensor) path exists
        from .models import plain_propagate
        return plain_propagate(x, edge_index, self.att_l.view(-1), self.att_r.view(-1), self.negative_slope)
