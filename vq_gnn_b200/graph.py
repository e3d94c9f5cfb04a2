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


class BatchPlan:
    """Kernel-ready description of one mini-batch graph (see the module docstring).

    The merged forward CSR (`fwd_rowptr / fwd_col / fwd_val / fwd_rval`) may be LAZY: the v1 builder produces
    the in-batch / tail split that the fast kernels consume directly and only sorts the merged form when some
    consumer asks for it (generic kernel, GAT, tests)."""

    def __init__(self, version: str, conv_type: str, B: int, R: int, T: int, N: int, batch_idx: Tensor,
                 fwd_rowptr: Optional[Tensor], fwd_col: Optional[Tensor], fwd_val: Optional[Tensor],
                 fwd_rval: Optional[Tensor], tail_node: Optional[Tensor], bwd_rowptr: Optional[Tensor],
                 bwd_col: Optional[Tensor], bwd_val: Optional[Tensor], bwd_eid: Optional[Tensor] = None,
                 training: bool = True, extras: Optional[dict] = None, merged_builder=None,
                 nnz: Optional[int] = None, has_rval: Optional[bool] = None, bwd_builder=None):
        self.version, self.conv_type = version, conv_type
        self.B, self.R, self.T, self.N = B, R, T, N
        self.batch_idx = batch_idx             # int32 [B]  global node ids of the batch rows
        self._fwd = None if fwd_rowptr is None else (fwd_rowptr, fwd_col, fwd_val, fwd_rval)
        self._merged_builder = merged_builder  # () -> (rowptr int32 [R+1], col int32, val fp32, rval fp32 | None)
        self.tail_node = tail_node             # int32 [T] or None (identity)
        # transposed CSR over the batch columns; may be DEFERRED (device-built v2 plans: the entry count is read back
        # from the device only when the backward first needs it, long after the kernels that produced it finished)
        self._bwd = None if bwd_rowptr is None else (bwd_rowptr, bwd_col, bwd_val)
        self._bwd_builder = bwd_builder          # () -> (rowptr int32 [B+1], src row int32 [nnzT], val fp32 [nnzT])
        self.bwd_eid = bwd_eid
        self.training = training
        self.extras = {} if extras is None else extras
        self._nnz = nnz
        self._has_rval = has_rval
        # v1: the in-batch block (forward split + transposed CSR) is small (~2 x 10^5 entries at the Reddit shape);
        # finer chunks give its kernels enough warps to fill the machine
        self.small_chunk = small_chunk(int(bwd_col.numel())) if (version == 'v1' and bwd_col is not None) else MP_CHUNK

    def _merged(self):
        if self._fwd is None:
            self._fwd = self._merged_builder()
            self._nnz = int(self._fwd[1].numel())     # exact from now on (a device-built plan only had a bound)
            self.extras.pop('chunk_rows_fwd', None)
        return self._fwd

    def _bwd_get(self):
        if self._bwd is None:
            self._bwd = self._bwd_builder()
        return self._bwd

    bwd_rowptr = property(lambda self: self._bwd_get()[0])  # int32 [B+1]
    bwd_col = property(lambda self: self._bwd_get()[1])     # int32 [nnzT] source row ids (< R)
    bwd_val = property(lambda self: self._bwd_get()[2])     # fp32  [nnzT]
    fwd_rowptr = property(lambda self: self._merged()[0])   # int32 [R+1]
    fwd_col = property(lambda self: self._merged()[1])      # int32 [nnz]: < B dense row, >= B tail entry
    fwd_val = property(lambda self: self._merged()[2])      # fp32  [nnz]
    fwd_rval = property(lambda self: self._merged()[3])     # fp32  [nnz] reverse values (v1) or None

    @property
    def nnz_B(self) -> int:
        """Entries of the first B rows of the forward CSR (v2: the batch rows; rows >= B only feed info_backward)."""
        n = self.extras.get('nnz_B')
        if n is None:
            n = self.extras['nnz_B'] = int(self.fwd_rowptr[self.B])
        return n

    def entry_rows(self) -> Tensor:
        """int32 [nnz]: row id of every forward-CSR entry of the rows >= B (entries of batch rows are not filled);
        what vqgnn_mp_info streams instead of walking rowptr.  Built once per plan on the device."""
        t = self.extras.get('erow')
        if t is None:
            from . import _lib
            t = torch.empty(max(self.nnz, 1) + 4, dtype=torch.int32, device=self.device)
            _lib.check(_lib.load().vqgnn_csr_expand_rows(_lib.ptr(self.fwd_rowptr), self.B, self.R, _lib.ptr(t),
                                                         _lib.stream()))
            self.extras['erow'] = t
        return t

    @property
    def has_rval(self) -> bool:
        return self._has_rval if self._has_rval is not None else self.fwd_rval is not None

    @property
    def device(self):
        return self.batch_idx.device

    @property
    def nnz(self) -> int:
        return self._nnz if self._nnz is not None else int(self.fwd_col.numel())

    def tensors(self):
        """Every tensor the plan currently holds (for stream bookkeeping)."""
        out = [self.batch_idx, self.tail_node, self.bwd_eid]
        if self._fwd is not None:
            out += list(self._fwd)
        if self._bwd is not None:
            out += list(self._bwd)

        def rec(o):
            if isinstance(o, torch.Tensor):
                out.append(o)
            elif isinstance(o, (tuple, list)):
                for u in o:
                    rec(u)
            elif isinstance(o, dict):
                for u in o.values():
                    rec(u)
        rec(self.extras)
        return [t for t in out if isinstance(t, torch.Tensor)]

    def chunk_rows(self, which: str) -> Tensor:
        """nnz-balanced work partition of the forward ('fwd') or transposed ('bwd') CSR for the CUDA
        kernels (include/vqgnn.h: vqgnn_mp_chunk_rows).  Built lazily on the device, once per plan."""
        key = 'chunk_rows_' + which
        t = self.extras.get(key)
        if t is None:
            from . import _lib
            if which == 'fwd':
                rowptr = self.fwd_rowptr          # materialises a lazy merged CSR (and makes nnz exact) first
                nnz, rows = self.nnz, self.R
            elif which == 'fwdB':                 # the batch rows only (v2 split forward)
                rowptr, nnz, rows = self.fwd_rowptr, self.nnz_B, self.B
            else:
                rowptr, nnz, rows = self.bwd_rowptr, int(self.bwd_col.numel()), self.B
            chunk = MP_CHUNK if which in ('fwd', 'fwdB') else self.small_chunk
            _lib.require_device(rowptr)
            lib = _lib.load()
            n = int(lib.vqgnn_mp_num_chunks(nnz, chunk))
            t = torch.empty(max(n, 1), dtype=torch.int32, device=rowptr.device)
            _lib.check(lib.vqgnn_mp_chunk_rows(_lib.ptr(rowptr), rows, nnz, chunk, _lib.ptr(t), _lib.stream()))
            self.extras[key] = t
        return t

    def warm(self, split: bool = True) -> "BatchPlan":
        """Build every lazily-derived piece now (on the current stream): the work partitions and, for v1
        plans with long rows, the in-batch / tail split.  `LowRankGNN.prepare` calls this so that a prefetching
        loader pays for it (and its host syncs) off the training stream."""
        if self._bwd is not None:       # a deferred transposed CSR is finished by the first backward instead
            self.chunk_rows('bwd')
        if (split and self.version == 'v1' and self.has_rval and self.conv_type != 'GAT'
                and self.nnz >= TAIL_MIN_AVG_DEGREE * self.B):
            self.split_v1()
        else:
            self.chunk_rows('fwd')
        return self

    def split_v1(self):
        """v1 plans: the forward CSR split into its in-batch part (dense rows, generic kernel) and its tail
        part (out-of-batch neighbours, shared-memory codebook kernel).  Built once per plan.
        -> dict(inb=(rowptr, col, val, chunk_row, nnz), tail=(rowptr, node, val, rval, chunk_row, nnz))"""
        sp = self.extras.get('split')
        if sp is None:
            from . import _lib
            assert self.version == 'v1' and self.has_rval
            lib = _lib.load()
            B = self.B

            def chunks(ptr, nnz, chunk):
                n = int(lib.vqgnn_mp_num_chunks(nnz, chunk))
                cr = torch.empty(max(n, 1), dtype=torch.int32, device=ptr.device)
                _lib.check(lib.vqgnn_mp_chunk_rows(_lib.ptr(ptr), B, nnz, chunk, _lib.ptr(cr), _lib.stream()))
                return cr

            raw = self.extras.get('split_raw')
            if raw is None:      # derive from the merged CSR
                dev = self.fwd_col.device
                deg = (self.fwd_rowptr[1:] - self.fwd_rowptr[:-1]).long()
                rows = torch.repeat_interleave(torch.arange(B, device=dev), deg)
                is_tail = self.fwd_col >= B

                def ptr_of(mask):
                    ptr = torch.zeros(B + 1, dtype=torch.int32, device=dev)
                    ptr[1:] = torch.cumsum(torch.bincount(rows[mask], minlength=B), 0).to(torch.int32)
                    return ptr
                raw = dict(tail=(ptr_of(is_tail), (self.fwd_col[is_tail] - B).contiguous(),
                                 self.fwd_val[is_tail].contiguous(), self.fwd_rval[is_tail].contiguous()),
                           inb=(ptr_of(~is_tail), self.fwd_col[~is_tail].contiguous(),
                                self.fwd_val[~is_tail].contiguous()))
            tptr, tnode, tval, trval = raw['tail']
            iptr, icol, ival = raw['inb']
            tn, inn = int(tnode.numel()), int(icol.numel())
            sp = dict(tail=(tptr, tnode, tval, trval, chunks(tptr, tn, TAIL_CHUNK), tn),
                      inb=(iptr, icol, ival, chunks(iptr, inn, self.small_chunk), inn))
            self.extras['split'] = sp
        return sp


MP_CHUNK = 256   # CSR entries per warp task (multiple of 32)


def small_chunk(nnz: int) -> int:
    """Chunk size for a small CSR: 64 entries per warp task below 2^20 entries."""
    return 64 if nnz < (1 << 20) else MP_CHUNK


DEVICE_PLAN_BUILDER = True   # plans on CUDA are built by csrc/plan.cu (False: the torch builders below)
import os as _os
TAIL_CHUNK = int(_os.environ.get('VQGNN_TAIL_CHUNK', '512'))  # entries per warp task of the shared-memory tail kernel (csrc/mp_tail.cu)
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


def plan_from_v2_device(batch_A, conv_type: str, N: int, device) -> BatchPlan:
    """Training-mode `plan_from_v2` with the transposed CSR built on the device (csrc/plan.cu:
    vqgnn_csr_transpose_lt) and NO host synchronisation: the number of transposed entries is copied to pinned host
    memory asynchronously and only read when the backward first asks for the structure."""
    from . import _lib
    batch_idx, subset, adj = batch_A
    lib, st = _lib.load(), _lib.stream()
    dev = torch.device(device)
    rowptr, col, val = adj.csr()
    dim = int(adj.sparse_sizes()[0])
    B = int(batch_idx.shape[0])
    rowptr, col = _i32(rowptr.to(dev)), _i32(col.to(dev))
    val = val.to(dev).float().contiguous()
    _lib.require_device(val)
    nnz = int(col.numel())
    browptr = torch.empty(B + 1, dtype=torch.int32, device=dev)
    brow = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    bval = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    ws = torch.empty(int(lib.vqgnn_csr_transpose_workspace_bytes(B, nnz)), dtype=torch.uint8, device=dev)
    _lib.check(lib.vqgnn_csr_transpose_lt(_lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), dim, nnz, B,
                                          _lib.ptr(browptr), _lib.ptr(brow), _lib.ptr(bval), _lib.ptr(count),
                                          _lib.ptr(ws), st))
    host_count = torch.empty(1, dtype=torch.int32).pin_memory()    # owned by this plan (kept alive by the closure)
    host_count.copy_(count, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()

    def bwd():
        ev.synchronize()
        n = int(host_count[0])
        return browptr, brow[:n], bval[:n]

    return BatchPlan('v2', conv_type, B, dim, dim - B, N, _i32(batch_idx.to(dev)), rowptr, col, val, None,
                     _i32(subset.to(dev)[B:]), None, None, None, None, True,
                     extras={'_keep': (ws, count, brow, bval, browptr)}, bwd_builder=bwd)


def _csr_ptr(rows: Tensor, n: int) -> Tensor:
    ptr = torch.zeros(n + 1, dtype=torch.int32, device=rows.device)
    if rows.numel() > 0:
        ptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0).to(torch.int32)
    return ptr


def plan_from_v1_device(batch_A, conv_type: str, N: int, training: bool, device) -> BatchPlan:
    """`plan_from_v1` built by libvqgnn on the device (csrc/plan.cu: vqgnn_plan_v1_build): no host
    synchronisation and no torch sort/mask ops, so a loader can enqueue it behind the H2D copy of the batch on a
    side stream from the training thread itself.  The number of tail entries stays on the device."""
    from . import _lib
    deg_inv, A_BN, A_BB, A_NB_v, batch_idx = batch_A
    lib, st = _lib.load(), _lib.stream()
    dev = torch.device(device)
    B = int(batch_idx.shape[0])
    r, c, v = (t.to(dev) for t in A_BN)
    r, c, v = r.long().contiguous(), c.long().contiguous(), v.float().contiguous()
    _lib.require_device(v)
    nnz = int(r.numel())
    if conv_type == 'GCN':
        rv = v
    elif A_NB_v is not None:
        rv = A_NB_v.to(dev).float().contiguous()
    else:
        rv = None
    if A_BB is not None:
        br, bc, bv = (t.to(dev) for t in A_BB)
        br, bc, bv = br.long().contiguous(), bc.long().contiguous(), bv.float().contiguous()
        nbb = int(br.numel())
        if nbb == 0:   # keep the "A_BB present" semantics with valid (non-NULL) pointers
            br = bc = torch.zeros(1, dtype=torch.long, device=dev)
            bv = torch.zeros(1, device=dev)
    else:
        br = bc = bv = None
        nbb = 0
    sym, loops = int(conv_type == 'GCN'), int(conv_type != 'SAGE')
    nin = nbb * (1 + sym) + loops * B
    bidx = batch_idx.to(dev).long().contiguous()
    dinv = deg_inv.to(dev).float().contiguous()
    i32 = lambda n: torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    f32 = lambda n: torch.empty(max(n, 1), dtype=torch.float32, device=dev)
    n_tc = int(lib.vqgnn_mp_num_chunks(nnz, TAIL_CHUNK))
    ichunk = small_chunk(nin)
    n_ic = int(lib.vqgnn_mp_num_chunks(nin, ichunk))
    tptr, tnode, tval, trval, tcount, tcr = i32(B + 1), i32(nnz), f32(nnz), f32(nnz), i32(1), i32(n_tc)
    iptr, icol, ival, icr = i32(B + 1), i32(nin), f32(nin), i32(n_ic)
    bptr, brow, bval, bcr = i32(B + 1), i32(nin), f32(nin), i32(n_ic)
    ws = torch.empty(int(lib.vqgnn_plan_v1_workspace_bytes(N, B, nnz, nin)), dtype=torch.uint8, device=dev)
    _lib.check(lib.vqgnn_plan_v1_build(
        _lib.ptr(r), _lib.ptr(c), _lib.ptr(v), _lib.ptr(rv), nnz, _lib.ptr(br), _lib.ptr(bc), _lib.ptr(bv), nbb,
        _lib.ptr(bidx), _lib.ptr(dinv), B, N, sym, loops, TAIL_CHUNK, ichunk,
        _lib.ptr(tptr), _lib.ptr(tnode), _lib.ptr(tval), _lib.ptr(trval), _lib.ptr(tcount), _lib.ptr(tcr),
        _lib.ptr(iptr), _lib.ptr(icol), _lib.ptr(ival), _lib.ptr(icr),
        _lib.ptr(bptr), _lib.ptr(brow), _lib.ptr(bval), _lib.ptr(bcr), _lib.ptr(ws), st))

    def merged():   # generic-kernel / GAT / test consumers: needs the tail count on the host
        nt = int(tcount.item())
        rows_t = torch.repeat_interleave(torch.arange(B, device=dev), (tptr[1:] - tptr[:-1]).long())
        rows_i = torch.repeat_interleave(torch.arange(B, device=dev), (iptr[1:] - iptr[:-1]).long())
        mr = torch.cat([rows_t, rows_i])
        mc = torch.cat([tnode[:nt].long() + B, icol[:nin].long()])
        order = torch.argsort(mr * (B + N) + mc, stable=True)
        mv = torch.cat([tval[:nt], ival[:nin]])[order]
        mrv = torch.cat([trval[:nt], torch.zeros(nin, device=dev)])[order]
        return _csr_ptr(mr, B), _i32(mc[order]), mv.contiguous(), mrv.contiguous()

    extras = {'split': dict(tail=(tptr, tnode, tval, trval, tcr, nnz, tcount),
                            inb=(iptr, icol[:nin], ival[:nin], icr, nin)),
              'chunk_rows_bwd': bcr, '_keep': (r, c, v, rv, br, bc, bv, bidx, dinv, ws)}
    return BatchPlan('v1', conv_type, B, B, N, N, _i32(bidx), None, None, None, None, None, bptr, brow[:nin],
                     bval[:nin], None, training, extras=extras, merged_builder=merged, nnz=nnz + nin, has_rval=True)


def plan_from_v1(batch_A, conv_type: str, N: int, training: bool, device) -> BatchPlan:
    """batch_A = (deg_inv[B], A_BN(r,c,v), A_BB(r,c,v)|None, A_NB_v|None, batch_idx[B])
    (vq_gnn_v1/utils/dataloader.py:86, mapper :144-192).

    Assumes, as the reference's loader guarantees (dataloader.py:69-73), that A_BB is the restriction
    of A_BN to in-batch columns with the same values; then mapper's "+v on the codeword column, -v
    cancellation, drop <= 0" (:149-180) is exactly "out-of-batch entries only", which is what is built.

    Produced directly in split form -- tail entries (a stable compaction of A_BN, no sort) and the small in-batch
    block (A_BB, + its transpose for GCN, + self loops) -- because that is what the fast kernels read; the merged,
    (row, col)-sorted CSR is assembled lazily for the consumers that want it."""
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
    # ---- tail part ---------------------------------------------------------------------------------
    if A_BB is not None:
        pos = torch.full((N,), -1, dtype=torch.long, device=dev)
        pos[batch_idx] = torch.arange(B, device=dev)
        out = pos[c] < 0
        tr, tc, tv, trv = r[out], c[out], v[out], rv[out]
    else:
        tr, tc, tv, trv = r, c, v, rv
    if tr.numel() > 1 and not bool((tr[1:] >= tr[:-1]).all()):   # the reference's loader emits row-sorted COO
        order = torch.argsort(tr, stable=True)
        tr, tc, tv, trv = tr[order], tc[order], tv[order], trv[order]
    tail = (_csr_ptr(tr, B), _i32(tc), tv.contiguous(), trv.contiguous())
    # ---- in-batch part (small) -----------------------------------------------------------------------
    rows, cols, vals = [], [], []
    if A_BB is not None:
        br, bc, bv = (t.to(dev) for t in A_BB)
        br, bc, bv = br.long(), bc.long(), bv.float()
        if conv_type == 'GCN':                   # S + S^T on the in-batch block
            br, bc, bv = torch.cat([br, bc]), torch.cat([bc, br]), torch.cat([bv, bv])
        rows.append(br), cols.append(bc), vals.append(bv)
    if conv_type != 'SAGE':                      # self loops, value deg_inv (:182-185); doubled by to_symmetric
        d = torch.arange(B, device=dev)
        rows.append(d), cols.append(d)
        vals.append(deg_inv.to(dev).float() * (2.0 if conv_type == 'GCN' else 1.0))
    if rows:
        ir, ic, iv = torch.cat(rows), torch.cat(cols), torch.cat(vals)
        order = torch.argsort(ir * B + ic, stable=True)
        ir, ic, iv = ir[order], ic[order], iv[order]
    else:
        ir = ic = torch.zeros(0, dtype=torch.long, device=dev)
        iv = torch.zeros(0, device=dev)
    inb = (_csr_ptr(ir, B), _i32(ic), iv.contiguous())
    bptr, bcol, bval, _ = _transpose_lt(ir, ic, iv, B)
    nnz = int(tc.numel() + ic.numel())

    def merged():
        mr = torch.cat([tr, ir])
        mc = torch.cat([tc + B, ic])
        order = torch.argsort(mr * (B + N) + mc, stable=True)
        mv = torch.cat([tv, iv])[order]
        mrv = torch.cat([trv, torch.zeros_like(iv)])[order]
        return _csr_ptr(mr, B), _i32(mc[order]), mv.contiguous(), mrv.contiguous()

    return BatchPlan('v1', conv_type, B, B, N, N, _i32(batch_idx), None, None, None, None, None, bptr, bcol, bval,
                     None, training, extras={'split_raw': dict(tail=tail, inb=inb)}, merged_builder=merged,
                     nnz=nnz, has_rval=True)


def _pinned_i32(n: int = 1) -> Tensor:
    return torch.empty(n, dtype=torch.int32).pin_memory()


def plan_from_graph_v2(g, node_idx: Tensor, conv_type: str, training: bool = True,
                       batch_rows_only: bool = False) -> BatchPlan:
    """v2 batch plan straight from the device-resident graph `g` (rowptr int64, col / col32, val) and the batch's
    node ids: csrc/khop.cu restates `_k_hop_subgraph` + `prepare_batch_input` (vq_gnn_v2/dataloader.py:98-148,
    utils/misc.py:57-75) on the device -- subset = [batch ; ascending out-of-batch 1-hop neighbours], train keeps
    every edge inside the subset, eval the batch rows only -- and hands the relabelled int32 CSR to the kernels
    without ever forming the reference's int64 COO.  Two small device->host reads (B' and nnz) size the outputs.
    batch_rows_only: keep only the batch rows (the loader's train_flag=False structure) in a plan that is still
    marked `training` -- what the reference's init() feeds a model in train mode (main_node.py:17-37 uses the
    TEST loader)."""
    from . import _lib
    lib, st = _lib.load(), _lib.stream()
    dev = g.col.device
    N = int(g.N)
    ids = node_idx.to(dev).long().contiguous()
    _lib.require_device(g.val)
    B = int(ids.numel())
    rowptr, col, val = g.rowptr.contiguous(), g.col32, g.val.contiguous()
    ws = torch.empty(int(lib.vqgnn_khop_workspace_bytes(N, N)), dtype=torch.uint8, device=dev)
    bidx = torch.empty(B, dtype=torch.int32, device=dev)
    tail_all = torch.empty(N, dtype=torch.int32, device=dev)
    cnt = torch.empty(2, dtype=torch.int32, device=dev)
    _lib.check(lib.vqgnn_khop_mark(_lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(ids), B, N, _lib.ptr(bidx),
                                   _lib.ptr(tail_all), _lib.ptr(cnt), _lib.ptr(ws), st))
    host = _pinned_i32(2)
    host[:1].copy_(cnt[:1], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    T = int(host[0])
    tail_node = tail_all[:T]
    R = B + T if (training and not batch_rows_only) else B
    out_rowptr = torch.empty(R + 1, dtype=torch.int32, device=dev)
    _lib.check(lib.vqgnn_khop_count(_lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(ids), _lib.ptr(tail_node), B, R, N,
                                    _lib.ptr(out_rowptr), _lib.ptr(cnt[1:]), _lib.ptr(ws), st))
    host[1:].copy_(cnt[1:], non_blocking=True)
    host_b = _pinned_i32(1)
    host_b.copy_(out_rowptr[B:B + 1], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    nnz = int(host[1])
    out_col = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    out_val = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
    _lib.check(lib.vqgnn_khop_fill(_lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), _lib.ptr(ids), _lib.ptr(tail_node),
                                   B, R, N, _lib.ptr(out_rowptr), _lib.ptr(out_col), _lib.ptr(out_val), _lib.ptr(ws),
                                   st))
    out_col, out_val = out_col[:nnz], out_val[:nnz]
    # transposed CSR over the batch columns (backward structure), entry count read lazily
    browptr = torch.empty(B + 1, dtype=torch.int32, device=dev)
    brow = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    bval = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    tws = torch.empty(int(lib.vqgnn_csr_transpose_workspace_bytes(B, nnz)), dtype=torch.uint8, device=dev)
    _lib.check(lib.vqgnn_csr_transpose_lt(_lib.ptr(out_rowptr), _lib.ptr(out_col), _lib.ptr(out_val), R, nnz, B,
                                          _lib.ptr(browptr), _lib.ptr(brow), _lib.ptr(bval), _lib.ptr(count),
                                          _lib.ptr(tws), st))
    host_count = _pinned_i32(1)
    host_count.copy_(count, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()

    def bwd():
        ev.synchronize()
        n = int(host_count[0])
        return browptr, brow[:n], bval[:n]

    return BatchPlan('v2', conv_type, B, R, T, N, bidx, out_rowptr, out_col, out_val, None, tail_node, None, None,
                     None, None, training,
                     extras={'_keep': (ws, tws, count, brow, bval, browptr, tail_all, ids), 'nnz_B': int(host_b[0])},
                     bwd_builder=bwd)


def batch_from_graph_v1(g, node_idx: Tensor, train_flag: bool = True, recovery_flag: bool = True):
    """The reference's v1 batch tuple (deg_inv[B], A_BN (r, c, v), A_BB (r, c, v) | None, A_NB_v | None, batch_idx)
    (vq_gnn_v1/utils/dataloader.py:64-86) assembled on the device from the resident graph (csrc/khop.cu), in the
    reference's own COO format: what `mapper` / vqgnn_plan_v1_build consume.  One device->host read (the two
    entry counts)."""
    from . import _lib
    lib, st = _lib.load(), _lib.stream()
    dev = g.col.device
    N = int(g.N)
    ids = node_idx.to(dev).long().contiguous()
    _lib.require_device(g.val)
    B = int(ids.numel())
    rowptr, col, val = g.rowptr.contiguous(), g.col32, g.val.contiguous()
    with_bb = bool(recovery_flag and train_flag)
    with_nb = bool(g.conv_type != 'GCN' and train_flag)
    ws = torch.empty(int(lib.vqgnn_khop_workspace_bytes(N, N)), dtype=torch.uint8, device=dev)
    off_bn = torch.empty(B + 1, dtype=torch.int32, device=dev)
    off_bb = torch.empty(B + 1, dtype=torch.int32, device=dev)
    cnt = torch.empty(2, dtype=torch.int32, device=dev)
    _lib.check(lib.vqgnn_collate_v1_count(_lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(ids), B, N, int(with_bb),
                                          _lib.ptr(off_bn), _lib.ptr(off_bb), _lib.ptr(cnt), _lib.ptr(ws), st))
    host = _pinned_i32(2)
    host.copy_(cnt, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    nnz, nbb = int(host[0]), int(host[1])
    i64 = lambda n: torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    f32 = lambda n: torch.empty(max(n, 1), dtype=torch.float32, device=dev)
    bn_r, bn_c, bn_v = i64(nnz), i64(nnz), f32(nnz)
    nb_v = f32(nnz) if with_nb else None
    bb = (i64(nbb), i64(nbb), f32(nbb)) if with_bb else (None, None, None)
    deg_inv = f32(B)
    _lib.check(lib.vqgnn_collate_v1_fill(
        _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), _lib.ptr(g.deg), _lib.ptr(g.deg_inv), _lib.ptr(ids), B, N,
        _lib.ptr(off_bn), _lib.ptr(off_bb), _lib.ptr(bn_r), _lib.ptr(bn_c), _lib.ptr(bn_v), _lib.ptr(nb_v),
        _lib.ptr(bb[0]), _lib.ptr(bb[1]), _lib.ptr(bb[2]), _lib.ptr(deg_inv), _lib.ptr(ws), st))
    A_BN = (bn_r[:nnz], bn_c[:nnz], bn_v[:nnz])
    A_BB = (bb[0][:nbb], bb[1][:nbb], bb[2][:nbb]) if with_bb else None
    A_NB_v = nb_v[:nnz] if with_nb else None
    return deg_inv[:B], A_BN, A_BB, A_NB_v, ids


def build_plan(batch_A, conv_type: str, N: int, training: bool, device) -> BatchPlan:
    if isinstance(batch_A, BatchPlan):
        return batch_A
    if len(batch_A) == 3:
        if training and torch.device(device).type == 'cuda' and DEVICE_PLAN_BUILDER:
            return plan_from_v2_device(batch_A, conv_type, N, device)
        return plan_from_v2(batch_A, conv_type, N, training, device)
    if len(batch_A) == 5:
        if torch.device(device).type == 'cuda' and DEVICE_PLAN_BUILDER:
            return plan_from_v1_device(batch_A, conv_type, N, training, device)
        return plan_from_v1(batch_A, conv_type, N, training, device)
    raise ValueError("batch_A must be the v2 (batch_idx, subset, adj) or the v1 "
                     "(deg_inv, A_BN, A_BB, A_NB_v, batch_idx) tuple")
