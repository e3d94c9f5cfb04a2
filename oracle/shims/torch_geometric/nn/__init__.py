"""Minimal `MessagePassing` / `GCNConv` / `SAGEConv` / `GATConv` with the PyG 1.7.2 behaviour the
VQ-GNN reference relies on (v{1,2}/convs.py:26,95,124-131,229).  TEST INFRASTRUCTURE ONLY.

Restated behaviour (PyG 1.7.2 `MessagePassing.propagate` with a `torch_sparse.SparseTensor`):
  * adjacency convention: row = target i, col = source j ("adj_t"): out[i] = sum_j A[i,j] msg(j -> i);
  * a conv that implements `message_and_aggregate` (GCNConv, SAGEConv) takes the fused path
    `matmul(adj_t, x, reduce=aggr)`;
  * otherwise (GATConv) `__collect__` lifts `*_j` tensors with `index_select(node_dim, col)`, `*_i`
    tensors with `gather_csr(rowptr)`, passes `edge_weight = adj.storage.value()`, `index = row`,
    `ptr = rowptr`, `size_i = n_rows`, then `message(...)` and `aggregate = segment_csr(msg, ptr, aggr)`.
Parameter names/initialisers follow PyG 1.7.2 (`GCNConv.weight/.bias`, glorot/zeros;
`GATConv.lin_l/.att_l/.att_r`, glorot) so `state_dict()` keys and RNG consumption match.
"""
import inspect
import math

import torch
from torch import Tensor
from torch.nn import Linear, Parameter
from torch_scatter import gather_csr, segment_csr
from torch_sparse import SparseTensor, matmul


def glorot(tensor):
    if tensor is not None:
        stdv = math.sqrt(6.0 / (tensor.size(-2) + tensor.size(-1)))
        tensor.data.uniform_(-stdv, stdv)


def zeros(tensor):
    if tensor is not None:
        tensor.data.fill_(0)


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2):
        super().__init__()
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim
        self.fuse = type(self).message_and_aggregate is not MessagePassing.message_and_aggregate
        self._msg_params = [p for p in inspect.signature(self.message).parameters]

    def message_and_aggregate(self, adj_t, **kwargs):
        raise NotImplementedError

    def propagate(self, edge_index, size=None, **kwargs):
        if not isinstance(edge_index, SparseTensor):
            raise NotImplementedError("oracle shim: only SparseTensor adjacency is supported")
        if self.fuse:
            x = kwargs["x"]
            return self.message_and_aggregate(edge_index, x=x)
        rowptr, col, value = edge_index.csr()
        row = edge_index.storage.row()
        coll = {"edge_weight": value, "index": row, "ptr": rowptr,
                "size_i": edge_index.sparse_size(0), "size_j": edge_index.sparse_size(1)}
        for name in self._msg_params:
            if name.endswith("_j") or name.endswith("_i"):
                data = kwargs.get(name[:-2])
                if isinstance(data, (tuple, list)):
                    data = data[0] if name.endswith("_j") else data[1]
                if data is None:
                    coll[name] = None
                elif name.endswith("_j"):
                    coll[name] = data.index_select(0, col)
                else:
                    coll[name] = gather_csr(data, rowptr)
        msg = self.message(**{k: coll[k] for k in self._msg_params if k in coll})
        return segment_csr(msg, rowptr, reduce="sum")


class GCNConv(MessagePassing):
    def __init__(self, in_channels, out_channels, improved=False, cached=False, add_self_loops=True,
                 normalize=True, bias=True, **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(**kwargs)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached = improved, cached
        self.add_self_loops, self.normalize = add_self_loops, normalize
        self._cached_edge_index = self._cached_adj_t = None
        self.weight = Parameter(torch.Tensor(in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        glorot(self.weight)
        zeros(self.bias)

    def message(self, x_j, edge_weight):
        return x_j if edge_weight is None else edge_weight.view(-1, 1) * x_j

    def message_and_aggregate(self, adj_t, x):
        return matmul(adj_t, x, reduce=self.aggr)


class SAGEConv(MessagePassing):
    def __init__(self, in_channels, out_channels, normalize=False, root_weight=True, bias=True, **kwargs):
        kwargs.setdefault("aggr", "mean")
        super().__init__(**kwargs)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.normalize, self.root_weight = normalize, root_weight
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.lin_l = Linear(in_channels[0], out_channels, bias=bias)
        if self.root_weight:
            self.lin_r = Linear(in_channels[1], out_channels, bias=False)

    def message(self, x_j):
        return x_j

    def message_and_aggregate(self, adj_t, x):
        raise NotImplementedError("oracle shim: SAGEConv is never instantiated by the reference")


class GATConv(MessagePassing):
    # OurGATConv overrides __init__/forward/message and calls
    # super(GATConv, self).__init__(node_dim=0, aggr='add'); only reset_parameters is inherited (v2).
    def reset_parameters(self):
        glorot(self.lin_l.weight)
        glorot(self.lin_r.weight)
        glorot(self.att_l)
        glorot(self.att_r)
        zeros(self.bias)

    def message(self, x_j, alpha_j, alpha_i, index, ptr, size_i):
        raise NotImplementedError
