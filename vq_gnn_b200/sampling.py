"""Mini-batch graph construction: the producers of the hot path's `batch_A` input.

Restates, in device-agnostic torch (runs on the GPU in bench.py, on CPU in tests):
  * v2 `_k_hop_subgraph` (vq_gnn_v2/dataloader.py:98-148): subset = [B ; B'] with the batch nodes
    first, train keeps every edge inside B u B', eval keeps only rows in B, nodes relabelled;
  * v1 `__collate__` tail (vq_gnn_v1/utils/dataloader.py:64-86): `(deg_inv[B], A_BN, A_BB, A_NB_v, batch_idx)`;
  * the `node` / `cluster` / `cont` / `rw` / `edge` samplers (vq_gnn_v2/dataloader.py:52-96); METIS itself is not
    rebuilt: contiguous id blocks of a planted-partition graph stand in for its parts.
The order of the B' nodes is unspecified in the reference (`unique(sorted=False)`); here it is ascending.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .graph import CSRAdj
from .synth import Graph

Tensor = torch.Tensor


def _row_gather(g: Graph, node_idx: Tensor):
    """All stored entries of rows `node_idx` (in order): (local_row, col, edge_id)."""
    start, end = g.rowptr[node_idx], g.rowptr[node_idx + 1]
    deg = end - start
    total = int(deg.sum())
    local_row = torch.repeat_interleave(torch.arange(node_idx.numel(), device=deg.device), deg)
    first = torch.cumsum(deg, 0) - deg
    eid = torch.arange(total, device=deg.device) - torch.repeat_interleave(first, deg) \
        + torch.repeat_interleave(start, deg)
    return local_row, g.col[eid], eid


def k_hop_batch_v2(g: Graph, node_idx: Tensor, train_flag: bool = True):
    """-> (batch_idx, subset, CSRAdj[(B+B')^2]) == v2 `prepare_batch_input`'s `batch_A`
    (vq_gnn_v2/utils/misc.py:57-75) without the host round trip."""
    N, dev = g.N, g.col.device
    node_idx = node_idx.to(dev)
    B = node_idx.numel()
    lrow, ncol, eid = _row_gather(g, node_idx)
    in_batch = torch.zeros(N, dtype=torch.bool, device=dev)
    in_batch[node_idx] = True
    rest = torch.unique(ncol[~in_batch[ncol]])
    subset = torch.cat([node_idx, rest])
    pos = torch.full((N,), -1, dtype=torch.long, device=dev)
    pos[subset] = torch.arange(subset.numel(), device=dev)
    if train_flag and rest.numel() > 0:
        lrow2, ncol2, eid2 = _row_gather(g, rest)
        keep2 = pos[ncol2] >= 0
        rows = torch.cat([lrow, lrow2[keep2] + B])
        cols = torch.cat([pos[ncol], pos[ncol2[keep2]]])
        vals = torch.cat([g.val[eid], g.val[eid2[keep2]]])
    else:
        rows, cols, vals = lrow, pos[ncol], g.val[eid]
    dim = subset.numel()
    adj = CSRAdj.from_coo(rows, cols, vals, (dim, dim))
    return node_idx, subset, adj


def collate_batch_v1(g: Graph, node_idx: Tensor, train_flag: bool = True, recovery_flag: bool = True):
    """-> (deg_inv[B], A_BN (r,c,v), A_BB (r,c,v)|None, A_NB_v|None, batch_idx)
    (vq_gnn_v1/utils/dataloader.py:64-86)."""
    N, dev = g.N, g.col.device
    node_idx = node_idx.to(dev)
    lrow, ncol, eid = _row_gather(g, node_idx)
    val = g.val[eid]
    A_BN = (lrow, ncol, val)
    A_BB = None
    if recovery_flag and train_flag:
        pos = torch.full((N,), -1, dtype=torch.long, device=dev)
        pos[node_idx] = torch.arange(node_idx.numel(), device=dev)
        keep = pos[ncol] >= 0
        A_BB = (lrow[keep], pos[ncol[keep]], val[keep])
    A_NB_v = None
    if g.conv_type != 'GCN' and train_flag:
        # deg[node_idx].view(-1,1) * A_BN * deg_inv.view(1,-1)   (dataloader.py:77-78)
        A_NB_v = g.deg[node_idx][lrow] * val * g.deg_inv[ncol]
    return g.deg_inv[node_idx], A_BN, A_BB, A_NB_v, node_idx


def cont_sampler(g: Graph, seeds: Tensor, walk_length: int, batch_size: int,
                 generator: Optional[torch.Generator] = None) -> List[Tensor]:
    """`cont` sampler (vq_gnn_v2/dataloader.py:77-88, sliding window 1): seeds, then `walk_length`
    rounds of (x3 replicate -> 1-step random walk -> unique -> first batch_size)."""
    dev = g.col.device
    out, cur = [seeds.to(dev)], seeds.to(dev)
    for _ in range(walk_length):
        cur = torch.cat([cur] * 3)
        deg = g.rowptr[cur + 1] - g.rowptr[cur]
        r = (torch.rand(cur.numel(), generator=generator, device=dev) * deg.clamp(min=1)).long()
        nxt = g.col[(g.rowptr[cur] + r).clamp(max=g.nnz - 1)]
        cur = torch.unique(torch.where(deg > 0, nxt, cur))[:batch_size]
        out.append(cur)
    return out


def random_walk(g: Graph, start: Tensor, walk_length: int,
                generator: Optional[torch.Generator] = None) -> Tensor:
    """`adj_t.random_walk(start, walk_length)` of torch_sparse as used by the loaders: uniform next-neighbour steps,
    a node without neighbours stays where it is.  -> [len(start), walk_length + 1] node ids (column 0 = start)."""
    dev = g.col.device
    cur = start.to(dev)
    out = [cur]
    for _ in range(walk_length):
        deg = g.rowptr[cur + 1] - g.rowptr[cur]
        r = (torch.rand(cur.numel(), generator=generator, device=dev) * deg.clamp(min=1)).long()
        nxt = g.col[(g.rowptr[cur] + r).clamp(max=max(g.nnz - 1, 0))]
        cur = torch.where(deg > 0, nxt, cur)
        out.append(cur)
    return torch.stack(out, 1)


def rw_sampler(g: Graph, seeds: Tensor, walk_length: int,
               generator: Optional[torch.Generator] = None) -> Tensor:
    """`rw` sampler (vq_gnn_v2/dataloader.py:74-75; v1 utils/dataloader.py:44-45): every node visited by one random
    walk of `walk_length` steps from each seed (the loader passes batch_size // (walk_length + 1) seeds)."""
    return torch.unique(random_walk(g, seeds, walk_length, generator).reshape(-1))


def edge_sampler(g: Graph, seeds: Tensor, generator: Optional[torch.Generator] = None) -> Tensor:
    """`edge` sampler (vq_gnn_v2/dataloader.py:71-72): the end points of one random edge per seed
    (the loader passes batch_size // 2 seeds)."""
    return torch.unique(random_walk(g, seeds, 1, generator).reshape(-1))


def cluster_batch(N: int, num_parts: int, parts: Tensor) -> Tensor:
    """Concatenate contiguous-range parts (the planted-partition stand-in for METIS clusters)."""
    bs = (N + num_parts - 1) // num_parts
    chunks = [torch.arange(int(p) * bs, min((int(p) + 1) * bs, N)) for p in parts]
    return torch.cat(chunks)
