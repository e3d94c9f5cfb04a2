"""Batch-graph containers and the kernel-ready "plan" for one mini-batch.

`CSRAdj` duck-types the slice of `torch_sparse.SparseTensor` the reference's layers touch
(`.csr()`, `.coo()`, `.sparse_sizes()`, `.to()`; vq_gnn_v2/utils/misc.py:73-74), so a real SparseTensor,
the oracle's shim, or this container can be handed to `LowRankGNNLayer.forward` interchangeably.

`BatchPlan` is what the CUDA message-passing kernels consume.  It is built ONCE per mini-batch (the
reference rebuilds/sorts a (B+M)^2 matrix per branch per layer in `mapper`,
vq_gnn_v1/utils/dataloader.py:144-192) and covers both formulations with one layout:

  forward CSR over R output rows; column ids  c <  B : batch row c (dense features)
                                              c >= B : "tail" entry t = c - B, a node that is only
                                                       known through its codewords; its global node
                                                       id is tail_node[t] (v2: subset[B:]) or t (v1)
  v2 (vq_gnn_v2/models.py:144-231): R = B + B' in training (rows >= B only feed `info_backward`), B in eval.
  v1 (vq_gnn_v1/models.py:143-233 + mapper): R = B; every out-of-batch neighbour j of row i is a tail
      entry with value A[i,j] and reverse value rval = A_NB_v (SAGE/GAT) or A[i,j] (GCN, to_symmetric);
      the (B+M)^2 matrix and its per-codeword sums Q[i,m] are never materialised: the kernel adds
      val * codebook[code[j]] per edge, which equals sum_m Q[i,m] * codebook[m].
  backward CSR ("transposed"): for batch column j < B the list of (source row i, value) pairs.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch

Tensor = torch.Tensor


class CSRAdj:
    """Minimal CSR sparse matrix (row = target, col = source; the reference's `adj_t` convention)."""

    def __init__(self, rowptr: Tensor, col: Tensor, value: Tensor, sparse_sizes: Tuple[int, int]):
        self._rowptr, self._col, self._value = rowptr, col, value
        self._sizes = (int(sparse_sizes[0]), int(sparse_sizes[1]))
        self._row = None

    @classmethod
    def from_coo(cls, row: Tensor, col: Tensor, value: Tensor, sparse_sizes) -> "CSRAdj":
        m, n = int(sparse_sizes[0]), int(sparse_sizes[1])
        row, col = row.long(), col.long()
        order = torch.argsort(row * n + col, stable=True)
        row, col, value = row[order], col[order], value[order]
        rowptr = torch.zeros(m + 1, dtype=torch.long, device=row.device)
        rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=m), 0)
        out = cls(rowptr, col, value, (m, n))
        out._row = row
        return out

    @classmethod
    def from_dense(cls, mat: Tensor) -> "CSRAdj":
        idx = mat.nonzero().t()
        return cls.from_coo(idx[0], idx[1], mat[idx[0], idx[1]], mat.shape)

    def csr(self):
        return self._rowptr, self._col, self._value

    def coo(self):
        if self._row is None:
            deg = self._rowptr[1:] - self._rowptr[:-1]
            self._row = torch.repeat_interleave(
                torch.arange(self._sizes[0], device=self._col.device), deg)
        return self._row, self._col, self._value

    def sparse_sizes(self):
        return self._sizes

    def size(self, dim):
        return self._sizes[dim]

    def nnz(self):
        return int(self._col.numel())

    def to(self, device, non_blocking: bool = False) -> "CSRAdj":
        out = CSRAdj(self._rowptr.to(device, non_blocking=non_blocking),
                     self._col.to(device, non_blocking=non_blocking),
                     self._value.to(device, non_blocking=non_blocking), self._sizes)
        return out

    def pin_memory(self) -> "CSRAdj":
        return CSRAdj(self._rowptr.pin_memory(), self._col.pin_memory(), self._value.pin_memory(),
                      self._sizes)

    def to_dense(self) -> Tensor:
        row, col, val = self.coo()
        out = torch.zeros(self._sizes, dtype=val.dtype, device=val.device)
        return out.index_put_((row, col), val, accumulate=True)


@dataclass
class BatchPlan:
    version: str
    conv_type: str
    B: int
    R: int                       # forward output rows
    T: int                       # number of tail entries (v2: B'; v1: N, identity)
    N: int
    batch_idx: Tensor            # int32 [B]  global node ids of the batch rows
    fwd_rowptr: Tensor           # int32 [R+1]
    fwd_col: Tensor              # int32 [nnz]
    fwd_val: Tensor              # fp32  [nnz]
    fwd_rval: Optional[Tensor]   # fp32  [nnz] reverse values (v1) or None
    tail_node: Optional[Tensor]  # int32 [T] or None (identity)
    bwd_rowptr: Tensor           # int32 [B+1]
    bwd_col: Tensor              # int32 [nnzT]   source row ids (< R)
    bwd_val: Tensor              # fp32  [nnzT]
    bwd_eid: Optional[Tensor] = None  # int32 [nnzT] position of the same entry in the forward CSR (GAT)
    training: bool = True
    extras: dict = field(default_factory=dict)

    @property
    def device(self):
        return self.fwd_col.device

    @property
    def nnz(self) -> int:
        return int(self.fwd_col.numel())

    def chunk_rows(self, which: str) -> Tensor:
        """nnz-balanced work partition of the forward ('fwd') or transposed ('bwd') CSR for the CUDA
        kernels (include/vqgnn.h: vqgnn_mp_chunk_rows).  Built lazily on the device, once per plan."""
        key = 'chunk_rows_' + which
        t = self.extras.get(key)
        if t is None:
            from . import _lib
            rowptr, nnz, rows = ((self.fwd_rowptr, self.nnz, self.R) if which == 'fwd'
                                 else (self.bwd_rowptr, int(self.bwd_col.numel()), self.B))
            _lib.require_device(rowptr)
            lib = _lib.load()
            n = int(lib.vqgnn_mp_num_chunks(nnz, MP_CHUNK))
            t = torch.empty(max(n, 1), dtype=torch.int32, device=rowptr.device)
            _lib.check(lib.vqgnn_mp_chunk_rows(_lib.ptr(rowptr), rows, nnz, MP_CHUNK, _lib.ptr(t), _lib.stream()))
            self.extras[key] = t
        return t

    def warm(self, split: bool = True) -> "BatchPlan":
        """Build every lazily-derived piece now (on the current stream): the work partitions and, for v1
        plans with long rows, the in-batch / tail split.  `LowRankGNN.prepare` calls this so that a prefetching
        loader pays for it (and its host syncs) off the training stream."""
        self.chunk_rows('fwd')
        self.chunk_rows('bwd')
        if (split and self.version == 'v1' and self.fwd_rval is not None
                and self.nnz >= TAIL_MIN_AVG_DEGREE * self.B):
            self.split_v1()
        return self

    def split_v1(self):
        """v1 plans: the forward CSR split into its in-batch part (dense rows, generic kernel) and its tail
        part (out-of-batch neighbours, shared-memory codebook kernel).  Built lazily, once per plan.
        -> dict(inb=(rowptr, col, val, chunk_row, nnz), tail=(rowptr, node, val, rval, chunk_row, nnz))"""
        sp = self.extras.get('split')
        if sp is None:
            from . import _lib
            assert self.version == 'v1' and self.fwd_rval is not None
            lib = _lib.load()
            dev = self.fwd_col.device
            B = self.B
            deg = (self.fwd_rowptr[1:] - self.fwd_rowptr[:-1]).long()
            rows = torch.repeat_interleave(torch.arange(B, device=dev), deg)
            is_tail = self.fwd_col >= B

            def csr(mask, chunk):
                r = rows[mask]
                ptr = torch.zeros(B + 1, dtype=torch.int32, device=dev)
                ptr[1:] = torch.cumsum(torch.bincount(r, minlength=B), 0).to(torch.int32)
                nnz = int(r.numel())
                n = int(lib.vqgnn_mp_num_chunks(nnz, chunk))
                cr = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
                _lib.check(lib.vqgnn_mp_chunk_rows(_lib.ptr(ptr), B, nnz, chunk, _lib.ptr(cr), _lib.stream()))
                return ptr, cr, nnz

            tptr, tcr, tnnz = csr(is_tail, TAIL_CHUNK)
            iptr, icr, innz = csr(~is_tail, MP_CHUNK)
            sp = dict(
                tail=(tptr, (self.fwd_col[is_tail] - B).contiguous(), self.fwd_val[is_tail].contiguous(),
                      self.fwd_rval[is_tail].contiguous(), tcr, tnnz),
                inb=(iptr, self.fwd_col[~is_tail].contiguous(), self.fwd_val[~is_tail].contiguous(), icr, innz))
            self.extras['split'] = sp
        return sp


MP_CHUNK = 256   # CSR entries per warp task (multiple of 32)


TAIL_CHUNK = 512  # entries per warp task of the shared-memory tail kernel (csrc/mp_tail.cu)
TAIL_MIN_AVG_DEGREE = 32   # below this the per-row reduce of the lane=entry kernel does not pay


def _i32(t: Tensor) -> Tensor:
    return t.to(torch.int32).contiguous()


def _transpose_lt(rows: Tensor, cols: Tensor, vals: Tensor, B: int):
    """CSR over columns < B: returns (rowptr[B+1], src_row, val, eid)."""
    sel = (cols < B).nonzero().flatten()
    c, r, v = cols[sel], rows[sel], vals[sel]
    order = torch.argsort(c, stable=True)
    c, r, v, eid = c[order], r[order], v[order], sel[order]
    ptr = torch.zeros(B + 1, dtype=torch.long, device=cols.device)
    if c.numel() > 0:
        ptr[1:] = torch.cumsum(torch.bincount(c, minlength=B), 0)
    return _i32(ptr), _i32(r), v.float().contiguous(), _i32(eid)


def plan_from_v2(batch_A, conv_type: str, N: int, training: bool, device) -> BatchPlan:
    """batch_A = (batch_idx[B], subset[B+B'], adj (B+B')^2)  (vq_gnn_v2/models.py:157)."""
    batch_idx, subset, adj = batch_A
    rowptr, col, val = adj.csr()
    dim = int(adj.sparse_sizes()[0])
    B = int(batch_idx.shape[0])
    rowptr, col = rowptr.to(device).long(), col.to(device).long()
    val = val.to(device).float()
    R = dim if training else B       # eval: rows >= B are never read (info_backward unused)
    if R < dim:
        nnz = int(rowptr[R])
        rowptr, col, val = rowptr[:R + 1], col[:nnz], val[:nnz]
    deg = rowptr[1:] - rowptr[:-1]
    rows = torch.repeat_interleave(torch.arange(R, device=device), deg)
    bptr, bcol, bval, beid = _transpose_lt(rows, col, val, B)
    return BatchPlan('v2', conv_type, B, R, dim - B, N, _i32(batch_idx.to(device)),
                     _i32(rowptr), _i32(col), val.contiguous(), None, _i32(subset.to(device)[B:]),
                     bptr, bcol, bval, beid, training)


def plan_from_v1(batch_A, conv_type: str, N: int, training: bool, device) -> BatchPlan:
    """batch_A = (deg_inv[B], A_BN(r,c,v), A_BB(r,c,v)|None, A_NB_v|None, batch_idx[B])
    (vq_gnn_v1/utils/dataloader.py:86, mapper :144-192).

    Assumes, as the reference's loader guarantees (dataloader.py:69-73), that A_BB is the restriction
    of A_BN to in-batch columns with the same values; then mapper's "+v on the codeword column, -v
    cancellation, drop <= 0" (:149-180) is exactly "out-of-batch entries only", which is what is built.
    """
    deg_inv, A_BN, A_BB, A_NB_v, batch_idx = batch_A
    B = int(batch_idx.shape[0])
    dev = device
    r, c, v = (t.to(dev) for t in A_BN)
    r, c, v = r.long(), c.long(), v.float()
    batch_idx = batch_idx.to(dev).long()
    if conv_type == 'GCN':
        rv = v                                   # to_symmetric(): M->B block is the transpose (:189-190)
    elif A_NB_v is not None:
        rv = A_NB_v.to(dev).float()              # (:153-154)
    else:
        rv = torch.zeros_like(v)                 # eval: no reverse block
    rows, cols, vals, rvals = [], [], [], []
    if A_BB is not None:
        pos = torch.full((N,), -1, dtype=torch.long, device=dev)
        pos[batch_idx] = torch.arange(B, device=dev)
        out = pos[c] < 0
        rows.append(r[out]), cols.append(c[out] + B), vals.append(v[out]), rvals.append(rv[out])
        br, bc, bv = (t.to(dev) for t in A_BB)
        br, bc, bv = br.long(), bc.long(), bv.float()
        if conv_type == 'GCN':                   # S + S^T on the in-batch block
            br, bc, bv = torch.cat([br, bc]), torch.cat([bc, br]), torch.cat([bv, bv])
        rows.append(br), cols.append(bc), vals.append(bv), rvals.append(torch.zeros_like(bv))
    else:
        rows.append(r), cols.append(c + B), vals.append(v), rvals.append(rv)
    if conv_type != 'SAGE':                      # self loops, value deg_inv (:182-185); doubled by to_symmetric
        d = torch.arange(B, device=dev)
        dv = deg_inv.to(dev).float() * (2.0 if conv_type == 'GCN' else 1.0)
        rows.append(d), cols.append(d), vals.append(dv), rvals.append(torch.zeros_like(dv))
    rows, cols = torch.cat(rows), torch.cat(cols)
    vals, rvals = torch.cat(vals), torch.cat(rvals)
    order = torch.argsort(rows * (B + N) + cols, stable=True)
    rows, cols, vals, rvals = rows[order], cols[order], vals[order], rvals[order]
    rowptr = torch.zeros(B + 1, dtype=torch.long, device=dev)
    rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=B), 0)
    bptr, bcol, bval, beid = _transpose_lt(rows, cols, vals, B)
    return BatchPlan('v1', conv_type, B, B, N, N, _i32(batch_idx), _i32(rowptr), _i32(cols),
                     vals.contiguous(), rvals.contiguous(), None, bptr, bcol, bval, beid, training)


def build_plan(batch_A, conv_type: str, N: int, training: bool, device) -> BatchPlan:
    if isinstance(batch_A, BatchPlan):
        return batch_A
    if len(batch_A) == 3:
        return plan_from_v2(batch_A, conv_type, N, training, device)
    if len(batch_A) == 5:
        return plan_from_v1(batch_A, conv_type, N, training, device)
    raise ValueError("batch_A must be the v2 (batch_idx, subset, adj) or the v1 "
                     "(deg_inv, A_BN, A_BB, A_NB_v, batch_idx) tuple")
