"""torch_geometric.utils subset (PyG 1.7.2 semantics), oracle shim."""
import torch

from .num_nodes import maybe_num_nodes


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    return edge_index[:, mask], (None if edge_attr is None else edge_attr[mask])


def add_self_loops(edge_index, edge_weight=None, fill_value=1., num_nodes=None):
    N = maybe_num_nodes(edge_index, num_nodes)
    loop = torch.arange(N, dtype=torch.long, device=edge_index.device).unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        edge_weight = torch.cat([edge_weight, edge_weight.new_full((N,), fill_value)])
    return torch.cat([edge_index, loop], dim=1), edge_weight


def add_remaining_self_loops(edge_index, edge_weight=None, fill_value=1., num_nodes=None):
    N = maybe_num_nodes(edge_index, num_nodes)
    row, col = edge_index[0], edge_index[1]
    mask = row != col
    loop_index = torch.arange(N, dtype=row.dtype, device=row.device).unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        inv = ~mask
        loop_weight = torch.full((N,), fill_value, dtype=edge_weight.dtype, device=edge_index.device)
        remaining = edge_weight[inv]
        if remaining.numel() > 0:
            loop_weight[row[inv]] = remaining
        edge_weight = torch.cat([edge_weight[mask], loop_weight])
    return torch.cat([edge_index[:, mask], loop_index], dim=1), edge_weight


def softmax(src, index=None, ptr=None, num_nodes=None, dim=0):
    N = maybe_num_nodes(index, num_nodes)
    mx = torch.full((N,) + tuple(src.shape[1:]), float('-inf'), dtype=src.dtype)
    mx = mx.scatter_reduce(0, index.view(-1, *([1] * (src.dim() - 1))).expand_as(src), src, 'amax')
    out = (src - mx[index]).exp()
    den = torch.zeros_like(mx).index_add_(0, index, out)
    return out / (den[index] + 1e-16)


def dense_to_sparse(adj):
    idx = adj.nonzero().t().contiguous()
    return idx, adj[idx[0], idx[1]]


def to_undirected(edge_index, num_nodes=None):
    N = maybe_num_nodes(edge_index, num_nodes)
    row, col = torch.cat([edge_index[0], edge_index[1]]), torch.cat([edge_index[1], edge_index[0]])
    key = torch.unique(row * N + col)
    return torch.stack([torch.div(key, N, rounding_mode='floor'), key % N])


from .subgraph import subgraph  # noqa: E402  (rebinding the submodule name to the function, as PyG does)
