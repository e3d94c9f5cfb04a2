"""Seeded synthetic graphs shaped like the reference's datasets (no dataset is available offline).

Plain torch, device-agnostic (CPU in tests, CUDA in bench.py).  Edge VALUES follow the
reference's normalisation exactly, because the message-passing kernels consume them:
  * v2: `norm_adj` (vq_gnn_v2/utils/misc.py:14-34): GCN  D~^-1/2 (A+I) D~^-1/2,
        SAGE D^-1 A (no self loop), GAT D~^-1 (A+I);
  * v1: `norm_adj` (vq_gnn_v1/main_node.py:323-349): degrees use deg+1 for GCN/GAT but NO diagonal
        entry is stored (the self loop is re-added per batch by `mapper` with value `deg_inv`).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

Tensor = torch.Tensor


@dataclass
class Graph:
    """Row-sorted COO/CSR of a normalised adjacency (`adj_t` convention: row = target)."""
    N: int
    rowptr: Tensor   # int64 [N+1]
    row: Tensor      # int64 [nnz]
    col: Tensor      # int64 [nnz]
    val: Tensor      # fp32  [nnz]
    deg: Optional[Tensor] = None       # v1 only (vq_gnn_v1/main_node.py:326,334,342)
    deg_inv: Optional[Tensor] = None   # v1 only
    conv_type: str = 'GCN'
    version: str = 'v2'

    @property
    def nnz(self) -> int:
        return int(self.col.numel())

    @property
    def col32(self) -> Tensor:
        """int32 mirror of `col` for the device-side batch builders (csrc/khop.cu); built once."""
        c = self.__dict__.get('_col32')
        if c is None or c.device != self.col.device:
            c = self.__dict__['_col32'] = self.col.to(torch.int32)
        return c

    def to(self, device) -> "Graph":
        mv = lambda t: None if t is None else t.to(device)
        return Graph(self.N, mv(self.rowptr), mv(self.row), mv(self.col), mv(self.val),
                     mv(self.deg), mv(self.deg_inv), self.conv_type, self.version)


def _csr_from_keys(key: Tensor, N: int):
    key = torch.unique(key)  # sorted, deduplicated (row-major)
    row = torch.div(key, N, rounding_mode='floor')
    col = key - row * N
    rowptr = torch.zeros(N + 1, dtype=torch.long, device=key.device)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=N), 0)
    return rowptr, row, col


def random_edges(N: int, num_undirected: int, seed: int = 0, power_law: float = 0.0,
                 num_blocks: int = 0, intra_frac: float = 0.6, device='cpu',
                 weights: bool = False):
    """Symmetric edge set without self loops.  power_law > 0: Chung-Lu endpoints with Pareto(power_law)
    weights; num_blocks > 0: planted partition over contiguous id ranges (stands in for METIS parts,
    vq_gnn_v2/utils/misc.py:93-130), `intra_frac` of the edges inside a block."""
    g = torch.Generator(device=device).manual_seed(seed)
    E = int(num_undirected)
    if power_law > 0:
        u = torch.rand(N, generator=g, device=device).clamp_(min=1e-6)
        # the CDF is accumulated in fp64 on the HOST: a parallel fp32 prefix sum on the GPU is not reproducible from
        # process to process, and every rank (and both bench arms) must build the SAME graph
        w = u.double().cpu().pow(-1.0 / power_law)
        cdf = torch.cumsum(w / w.sum(), 0).to(device)
        src = torch.searchsorted(cdf, torch.rand(E, generator=g, device=device).double()).clamp_(max=N - 1)
        dst = torch.searchsorted(cdf, torch.rand(E, generator=g, device=device).double()).clamp_(max=N - 1)
    else:
        src = torch.randint(0, N, (E,), generator=g, device=device)
        dst = torch.randint(0, N, (E,), generator=g, device=device)
    if num_blocks > 0:
        bs = (N + num_blocks - 1) // num_blocks
        intra = torch.rand(E, generator=g, device=device) < intra_frac
        off = torch.randint(0, bs, (E,), generator=g, device=device)
        dst_in = (torch.div(src, bs, rounding_mode='floor') * bs + off).clamp_(max=N - 1)
        dst = torch.where(intra, dst_in, dst)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    key = torch.cat([src * N + dst, dst * N + src])
    return _csr_from_keys(key, N)


def normalized_graph(N: int, rowptr: Tensor, row: Tensor, col: Tensor, conv_type: str,
                     version: str = 'v2', edge_weight: Optional[Tensor] = None) -> Graph:
    dev = col.device
    w = torch.ones(col.numel(), device=dev) if edge_weight is None else edge_weight.float()
    if version == 'v2':
        if conv_type in ('GCN', 'GAT'):     # set_diag(): drop existing diagonal, add 1.0 (misc.py:16,27)
            keep = row != col
            d = torch.arange(N, device=dev)
            key = torch.cat([row[keep] * N + col[keep], d * N + d])
            w = torch.cat([w[keep], torch.ones(N, device=dev)])
            order = torch.argsort(key)
            key, w = key[order], w[order]
            row = torch.div(key, N, rounding_mode='floor')
            col = key - row * N
            rowptr = torch.zeros(N + 1, dtype=torch.long, device=dev)
            rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=N), 0)
        deg = torch.zeros(N, device=dev).index_add_(0, row, w)
        if conv_type == 'GCN':
            dis = deg.pow(-0.5)
            dis[dis == float('inf')] = 0
            val = dis[row] * w * dis[col]
        else:
            di = deg.pow(-1)
            di[di == float('inf')] = 0
            val = di[row] * w
        return Graph(N, rowptr, row, col, val, None, None, conv_type, 'v2')
    # v1 (vq_gnn_v1/main_node.py:323-349)
    deg = torch.zeros(N, device=dev).index_add_(0, row, w)
    if conv_type in ('GCN', 'GAT'):
        deg = deg + 1
    deg_inv = deg.pow(-1)
    if conv_type == 'GCN':
        dis = deg.pow(-0.5)
        dis[dis == float('inf')] = 0
        val = dis[row] * w * dis[col]
    else:
        deg_inv[deg_inv == float('inf')] = 0
        val = deg_inv[row] * w
    return Graph(N, rowptr, row, col, val, deg, deg_inv, conv_type, 'v1')


def make_graph(N: int, num_undirected: int, conv_type: str, version: str = 'v2', seed: int = 0,
               power_law: float = 0.0, num_blocks: int = 0, intra_frac: float = 0.6,
               device='cpu') -> Graph:
    rowptr, row, col = random_edges(N, num_undirected, seed, power_law, num_blocks, intra_frac, device)
    return normalized_graph(N, rowptr, row, col, conv_type, version)


# Named shapes of BASELINE.json's configs.  `scale` shrinks N and E together for tests.
CONFIG_SHAPES = {
    'c1_arxiv':    dict(N=169_343, E=1_166_243, C=128, classes=40, M=256, D=4, conv='GCN', version='v2',
                        num_blocks=80, power_law=0.0),
    'c2_reddit':   dict(N=232_965, E=57_307_946, C=602, classes=41, M=1024, D=4, conv='SAGE', version='v1',
                        num_blocks=0, power_law=2.2),
    'c3_ppi':      dict(N=44_906, E=323_000, C=50, classes=121, M=4096, D=4, conv='GAT',
                        version='v2', num_blocks=20, power_law=0.0),
    'c4_collab':   dict(N=235_868, E=1_285_465, C=128, classes=128, M=1024, D=4, conv='GCN', version='v2',
                        num_blocks=0, power_law=2.5),
    'c5_products': dict(N=2_449_029, E=61_859_140, C=100, classes=47, M=4096, D=4, conv='GCN', version='v2',
                        num_blocks=64, power_law=0.0),
}
