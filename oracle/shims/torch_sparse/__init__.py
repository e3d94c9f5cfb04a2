"""Pure-torch stand-in for the subset of `torch_sparse` the VQ-GNN reference imports.

TEST INFRASTRUCTURE ONLY (oracle/): lets the *unmodified* reference files
(`vq_gnn_v{1,2}/convs.py`, `models.py`, `utils/dataloader.py`) import and run on
CPU in the builder container, where the real `torch_sparse` (un-vendored third
party dependency, pytorch_sparse 0.6.10-0.6.12 implied by the reference's
`torch-geometric >= 1.7.2`, README.md:13-15) is not installed.

Semantics restated from the published pytorch_sparse behaviour:
  * `SparseTensor(row, col, value, sparse_sizes)` sorts entries by (row, col)
    (stable) and does NOT merge duplicates.
  * `coalesce(index, value, m, n)` sorts by (row, col) and SUMS duplicates.
  * `to_symmetric()` returns A + A^T with duplicate coordinates summed
    (so entries present in both directions, and diagonal entries, are doubled).
  * `matmul(A, x, reduce='sum')` is a CSR SpMM: out[i] = sum_j A[i, j] * x[j].
Call sites in the reference: v1/utils/dataloader.py:144-192 (mapper),
v2/utils/misc.py:14-34,57-75, v{1,2}/convs.py:9,95.
"""
from typing import Optional

import torch
from torch import Tensor


def _ind2ptr(row: Tensor, n: int) -> Tensor:
    counts = torch.bincount(row, minlength=n)
    ptr = torch.zeros(n + 1, dtype=torch.long, device=row.device)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr


class _Storage:
    def __init__(self, row, col, value, sparse_sizes, rowptr):
        self._row, self._col, self._value = row, col, value
        self._sparse_sizes, self._rowptr = tuple(sparse_sizes), rowptr

    def row(self):
        return self._row

    def col(self):
        return self._col

    def value(self):
        return self._value

    def rowptr(self):
        return self._rowptr

    def sparse_sizes(self):
        return self._sparse_sizes


class SparseTensor:
    def __init__(self, row: Optional[Tensor] = None, rowptr: Optional[Tensor] = None,
                 col: Optional[Tensor] = None, value: Optional[Tensor] = None,
                 sparse_sizes=None, is_sorted: bool = False):
        assert row is not None and col is not None
        row, col = row.to(torch.long), col.to(torch.long)
        if sparse_sizes is None:
            sparse_sizes = (int(row.max()) + 1 if row.numel() else 0,
                            int(col.max()) + 1 if col.numel() else 0)
        m, n = int(sparse_sizes[0]), int(sparse_sizes[1])
        if not is_sorted and row.numel() > 0:
            key = row * n + col
            perm = torch.argsort(key, stable=True)
            row, col = row[perm], col[perm]
            if value is not None:
                value = value[perm]
        self.storage = _Storage(row, col, value, (m, n), _ind2ptr(row, m))

    # ---- constructors / conversions -------------------------------------------------
    @classmethod
    def from_edge_index(cls, edge_index, edge_attr=None, sparse_sizes=None):
        return cls(row=edge_index[0], col=edge_index[1], value=edge_attr, sparse_sizes=sparse_sizes)

    @classmethod
    def from_dense(cls, mat: Tensor):
        idx = mat.nonzero().t()
        return cls(row=idx[0], col=idx[1], value=mat[idx[0], idx[1]], sparse_sizes=mat.shape,
                   is_sorted=True)

    def coo(self):
        s = self.storage
        return s.row(), s.col(), s.value()

    def csr(self):
        s = self.storage
        return s.rowptr(), s.col(), s.value()

    def sparse_sizes(self):
        return self.storage.sparse_sizes()

    def sparse_size(self, dim):
        return self.storage.sparse_sizes()[dim]

    def size(self, dim):
        return self.storage.sparse_sizes()[dim]

    def sizes(self):
        return list(self.storage.sparse_sizes())

    def nnz(self):
        return self.storage.col().numel()

    def has_value(self):
        return self.storage.value() is not None

    def device(self):
        return self.storage.col().device

    def to(self, device, *_, **__):
        s = self.storage
        v = s.value()
        return SparseTensor(row=s.row().to(device), col=s.col().to(device),
                            value=None if v is None else v.to(device),
                            sparse_sizes=s.sparse_sizes(), is_sorted=True)

    def cpu(self):
        return self.to('cpu')

    def to_dense(self, dtype=None):
        row, col, value = self.coo()
        if value is None:
            value = torch.ones(row.numel(), dtype=dtype or torch.float, device=row.device)
        out = torch.zeros(self.sparse_sizes(), dtype=value.dtype, device=row.device)
        out.index_put_((row, col), value, accumulate=True)
        return out

    def set_value(self, value, layout=None):
        s = self.storage
        return SparseTensor(row=s.row(), col=s.col(), value=value, sparse_sizes=s.sparse_sizes(),
                            is_sorted=True)

    def fill_value(self, fill_value, dtype=None):
        v = torch.full((self.nnz(),), fill_value, dtype=dtype or torch.float, device=self.device())
        return self.set_value(v)

    def t(self):
        row, col, value = self.coo()
        m, n = self.sparse_sizes()
        return SparseTensor(row=col, col=row, value=value, sparse_sizes=(n, m))

    # ---- reductions / elementwise ----------------------------------------------------
    def sum(self, dim=None):
        row, col, value = self.coo()
        if value is None:
            value = torch.ones(row.numel(), device=row.device)
        if dim is None:
            return value.sum()
        m, n = self.sparse_sizes()
        if dim in (1, -1):
            return torch.zeros(m, dtype=value.dtype, device=row.device).index_add_(0, row, value)
        return torch.zeros(n, dtype=value.dtype, device=row.device).index_add_(0, col, value)

    def mul(self, other: Tensor):
        return mul(self, other)

    def __mul__(self, other):
        return mul(self, other)

    def __rmul__(self, other):
        return mul(self, other)

    # ---- structure -------------------------------------------------------------------
    def set_diag(self, values: Optional[Tensor] = None, k: int = 0):
        return set_diag(self, values, k)

    def to_symmetric(self, reduce: str = "sum"):
        assert reduce in ("sum", "add")
        row, col, value = self.coo()
        N = max(self.sparse_sizes())
        r2, c2 = torch.cat([row, col]), torch.cat([col, row])
        v2 = None if value is None else torch.cat([value, value])
        idx, v2 = coalesce(torch.stack([r2, c2]), v2, N, N)
        return SparseTensor(row=idx[0], col=idx[1], value=v2, sparse_sizes=(N, N), is_sorted=True)

    def __getitem__(self, index):
        # row selection by a LongTensor of row ids (v1/utils/dataloader.py:69)
        if isinstance(index, Tensor) and index.dtype == torch.long:
            return self.index_select(0, index)
        raise NotImplementedError

    def index_select(self, dim: int, idx: Tensor):
        assert dim == 0
        rowptr, col, value = self.csr()
        deg = rowptr[idx + 1] - rowptr[idx]
        new_row = torch.repeat_interleave(torch.arange(idx.numel()), deg)
        start = torch.repeat_interleave(rowptr[idx], deg)
        off = torch.arange(int(deg.sum())) - torch.repeat_interleave(torch.cumsum(deg, 0) - deg, deg)
        perm = start + off
        return SparseTensor(row=new_row, col=col[perm], value=None if value is None else value[perm],
                            sparse_sizes=(idx.numel(), self.sparse_sizes()[1]), is_sorted=True)

    def saint_subgraph(self, node_idx: Tensor):
        """A[node_idx][:, node_idx] with relabelled columns; returns (adj, edge_ids)."""
        row, col, value = self.coo()
        n = self.sparse_sizes()[1]
        pos = torch.full((n,), -1, dtype=torch.long)
        pos[node_idx] = torch.arange(node_idx.numel())
        sub = self.index_select(0, node_idx)
        r, c, v = sub.coo()
        keep = pos[c] >= 0
        out = SparseTensor(row=r[keep], col=pos[c[keep]], value=None if v is None else v[keep],
                           sparse_sizes=(node_idx.numel(), node_idx.numel()))
        return out, keep.nonzero().flatten()

    def random_walk(self, start: Tensor, walk_length: int):
        rowptr, col, _ = self.csr()
        cur = start.to(torch.long)
        out = [cur]
        for _ in range(walk_length):
            deg = rowptr[cur + 1] - rowptr[cur]
            r = (torch.rand(cur.numel()) * deg.clamp(min=1)).to(torch.long)
            nxt = col[(rowptr[cur] + r).clamp(max=max(col.numel() - 1, 0))]
            cur = torch.where(deg > 0, nxt, cur)
            out.append(cur)
        return torch.stack(out, dim=1)

    def __matmul__(self, other):
        return matmul(self, other)


def coalesce(index: Tensor, value: Optional[Tensor], m: int, n: int, op: str = "add"):
    assert op == "add"
    row, col = index[0].to(torch.long), index[1].to(torch.long)
    key = row * n + col
    ukey, inv = torch.unique(key, sorted=True, return_inverse=True)
    new_index = torch.stack([torch.div(ukey, n, rounding_mode='floor'), ukey % n])
    if value is None:
        return new_index, None
    out = torch.zeros((ukey.numel(),) + tuple(value.shape[1:]), dtype=value.dtype, device=value.device)
    out.index_add_(0, inv, value)
    return new_index, out


def matmul(src: SparseTensor, other: Tensor, reduce: str = "sum"):
    assert reduce in ("sum", "add")
    row, col, value = src.coo()
    msg = other.index_select(0, col)
    if value is not None:
        msg = msg * value.view(-1, *([1] * (other.dim() - 1))).to(other.dtype)
    out = torch.zeros((src.sparse_sizes()[0],) + tuple(other.shape[1:]), dtype=other.dtype,
                      device=other.device)
    return out.index_add(0, row, msg)


def set_diag(src: SparseTensor, values: Optional[Tensor] = None, k: int = 0):
    assert k == 0
    row, col, value = src.coo()
    m, n = src.sparse_sizes()
    keep = row != col
    d = torch.arange(min(m, n), device=row.device)
    if value is not None:
        dv = torch.ones(d.numel(), dtype=value.dtype, device=row.device) if values is None else values
        value = torch.cat([value[keep], dv])
    return SparseTensor(row=torch.cat([row[keep], d]), col=torch.cat([col[keep], d]), value=value,
                        sparse_sizes=(m, n))


def fill_diag(src: SparseTensor, fill_value: float, k: int = 0):
    m, n = src.sparse_sizes()
    return set_diag(src, torch.full((min(m, n),), float(fill_value)), k)


def sum(src: SparseTensor, dim=None):  # noqa: A001  (name mandated by torch_sparse)
    return src.sum(dim)


def mul(src: SparseTensor, other: Tensor):
    row, col, value = src.coo()
    if value is None:
        value = torch.ones(row.numel(), dtype=other.dtype, device=row.device)
    m, n = src.sparse_sizes()
    if other.dim() == 2 and other.size(0) == m and other.size(1) == 1:
        value = value * other.view(-1)[row]
    elif other.dim() == 2 and other.size(0) == 1 and other.size(1) == n:
        value = value * other.view(-1)[col]
    else:
        raise ValueError("shim mul: only [m,1] / [1,n] broadcasting is supported")
    return src.set_value(value)
