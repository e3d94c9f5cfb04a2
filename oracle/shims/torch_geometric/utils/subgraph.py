import torch

from .num_nodes import maybe_num_nodes


def subgraph(subset, edge_index, edge_attr=None, relabel_nodes=False, num_nodes=None):
    N = maybe_num_nodes(edge_index, num_nodes)
    mask = torch.zeros(N, dtype=torch.bool)
    mask[subset] = True
    emask = mask[edge_index[0]] & mask[edge_index[1]]
    edge_index = edge_index[:, emask]
    if relabel_nodes:
        pos = torch.full((N,), -1, dtype=torch.long)
        pos[subset] = torch.arange(int(mask.sum()) if subset.dtype == torch.bool else subset.numel())
        edge_index = pos[edge_index]
    return edge_index, (None if edge_attr is None else edge_attr[emask])
