"""`LowRankGNNBlock` / `LowRankGNNLayer` / `LowRankGNN` — drop-ins for vq_gnn_v{1,2}/models.py.

Constructor and forward signatures, return tuples, module tree and `state_dict()` keys follow the
reference (v2: models.py:11-63, 66-231, 234-374; v1: models.py:23-233, 236-367, 370-536).  One extra
keyword, `version` ('v2' default, 'v1'), selects which formulation a layer implements; `batch_A` is
accepted in the matching reference form:
    v2: (batch_idx[B], subset[B+B'], adj (B+B')^2)                      (vq_gnn_v2/models.py:157)
    v1: (deg_inv[B], A_BN(r,c,v), A_BB(r,c,v)|None, A_NB_v|None, batch_idx[B])   (vq_gnn_v1/utils/dataloader.py:86)
or as a prebuilt `graph.BatchPlan`.

What differs is the execution: the per-branch Python loops (gather -> cat -> conv -> hook, ~10^4 tiny
launches per step) become one fused CUDA launch per layer per direction over all branches, through
libvqgnn's C-ABI.  The VQ hook keeps v1 semantics: it runs inside the layer's backward on
(d loss / d conv-output, detached layer input) and only mutates quantiser state for the NEXT step.
`literal_v2_hooks=True` reproduces v2's dangling-slice behaviour (the hook never fires; SURVEY.md
Appendix B.1).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .convs import OurGATConv, OurGCNConv
from .graph import MP_CHUNK, TAIL_CHUNK, TAIL_MIN_AVG_DEGREE, BatchPlan, CSRAdj, build_plan

import os as _os
INFO_SLAB = int(_os.environ.get('VQGNN_INFO_SLAB', '16'))     # columns per slab of the split info kernel (16/32/64)
USE_ROWS_KERNEL = _os.environ.get('VQGNN_MPFWD_ROWS', '1') != '0'   # TMA row-gather forward (csrc/mp_rows.cuh)
INFO_SPLIT_MIN_ENTRIES = 1 << 20     # below this the one-kernel forward is as fast
from .vq import VectorQuantizerEMA, VQBank

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------------
# fused message passing as an autograd Function
# --------------------------------------------------------------------------------------------------
_TRIGGERS = {}


def _trigger(dev) -> Tensor:
    t = _TRIGGERS.get(dev)
    if t is None:
        t = _TRIGGERS[dev] = torch.zeros((), device=dev, requires_grad=True)
    return t


def _mp_ws(dev, nnz: int, chunk: int, C: int) -> Tensor:
    """Scratch of vqgnn_mp_fwd / vqgnn_mp_bwd: ordered info partials + the piece buffers of rows cut >= 2 times."""
    return torch.empty(int(_lib.load().vqgnn_mp_workspace_bytes(int(nnz), int(chunk), int(C))), dtype=torch.uint8,
                       device=dev)


def _tail_table(plan: BatchPlan, C: int) -> Tensor:
    """[T, C] table of materialised codeword rows, allocated with B spare rows IN FRONT (`.head_rows`, a [B, C] view):
    the row-gather kernel addresses the rows of both operands -- batch rows and codeword rows -- by 32-bit offsets from
    one base, so the batch rows are copied next to the table (x_input = cat(x, codewords) of vq_gnn_v2/models.py:161-179,
    literally) instead of hoping that two separate allocations of a 180 GB device land within one 64 GB window."""
    full = torch.empty(plan.B + plan.T, C, device=plan.device)
    t = full[plan.B:]
    t.head_rows = full[:plan.B]
    return t


def _rows_operand(x: Tensor, tail: Optional[Tensor]) -> Tensor:
    """The batch-row operand of vqgnn_mp_fwd_rows: x copied into the spare rows in front of `tail` when it has them."""
    head = getattr(tail, 'head_rows', None)
    if head is None or head.shape != x.shape:
        return x
    head.copy_(x)
    return head


def _rows_kernel_ok(x: Tensor, tail: Optional[Tensor], C: int) -> bool:
    """vqgnn_mp_fwd_rows takes this call: wide enough rows, 16 B aligned, and both tables inside one 64 GB window (the
    kernel addresses rows by 32-bit offsets in 16 B units from the lower of the two base pointers)."""
    if not USE_ROWS_KERNEL or C < 16 or C % 4 or x.stride(0) % 4 or x.data_ptr() % 16:
        return False
    if tail is None:
        return True
    lo = min(x.data_ptr(), tail.data_ptr())
    hi = max(x.data_ptr() + x.numel() * 4, tail.data_ptr() + tail.numel() * 4)
    return tail.data_ptr() % 16 == 0 and hi - lo < (1 << 36) - (1 << 20)


class _TallLinear(torch.autograd.Function):
    """F.linear for a tall input ([B, Cin], B >> Cin, Cout) whose WEIGHT gradient dW = dY^T X is a [Cout, Cin] product
    with a B-long contraction: cuBLAS picks a 64x64-tile SIMT kernel without split-K for it, i.e. four CTAs on a 148-SM
    part (88 us for 0.65 GFLOP at the products shape, 7 % of the step).  Here the contraction is split into
    `_TALL_SPLITS` batched products that are then added (a fixed order: deterministic).  Forward and input gradient are
    the library calls F.linear makes."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return F.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dy @ weight if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            B = x.shape[0]
            S = _TALL_SPLITS
            Bs = (B // S) * S
            xs = x[:Bs].reshape(S, Bs // S, x.shape[1])
            ds = dy[:Bs].reshape(S, Bs // S, dy.shape[1])
            dw = torch.bmm(ds.transpose(1, 2), xs).sum(0)
            if Bs < B:
                dw = dw + dy[Bs:].t() @ x[Bs:]
        db = dy.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db


_TALL_SPLITS = 32
_TALL_MIN_ROWS = 4096


def _linear(mod: nn.Linear, x: Tensor) -> Tensor:
    """`mod(x)`; tall CUDA inputs in training take the split-K weight gradient."""
    if (x.is_cuda and x.dim() == 2 and x.shape[0] >= _TALL_MIN_ROWS and torch.is_grad_enabled()
            and mod.weight.requires_grad and x.is_contiguous()):
        return _TallLinear.apply(x, mod.weight, mod.bias)
    return mod(x)


class VQConvFunction(torch.autograd.Function):
    """Y[:B], info_backward = conv([x ; codewords], adj)  for GCN / SAGE-Mean.

    forward : vqgnn_mp_fwd   (vq_gnn_v2/models.py:161-198 ; vq_gnn_v1/models.py:170-223 + mapper)
    backward: vqgnn_mp_bwd   (adj^T dY + d info_backward / dx), then the VQ hook
              (vq_gnn_v2/models.py:39-56 / vq_gnn_v1/models.py:71-125) via `bank.run`.
    """

    @staticmethod
    def forward(ctx, x: Tensor, trigger: Tensor, layer: "LowRankGNNLayer", plan: BatchPlan, wu: float,
                fire_hook: bool):
        # `trigger` is a dummy scalar that requires grad, so backward (and with it the VQ hook) runs even
        # when x itself needs no gradient (first layer) -- the reference gets the same effect from
        # X_output_B.requires_grad_() (vq_gnn_v1/models.py:202).
        _lib.require_device(x)
        lib, st = _lib.load(), _lib.stream()
        bank = layer.bank
        B, C = x.shape
        dev = x.device
        y = torch.empty(B, C, device=dev)
        need_info = plan.training
        info = torch.zeros((), device=dev)
        v1 = plan.version == 'v1'
        gq = torch.empty(B, C, device=dev) if (v1 and plan.has_rval) else None
        codes_g = None
        if v1 and gq is not None and layer.use_tail_kernel and (
                layer.use_tail_kernel == 'force' or plan.nnz >= TAIL_MIN_AVG_DEGREE * B):
            codes_g = bank.grouped_codes()
        if codes_g is not None:
            # in-batch entries through the generic kernel (zero-initialises y), then the out-of-batch entries
            # through the shared-memory codebook kernel, which accumulates on top (csrc/mp_tail.cu)
            sp = plan.split_v1()
            iptr, icol, ival, icr, innz = sp['inb']
            tptr, tnode, tval, trval, tcr, tnnz = sp['tail'][:6]
            tcount = sp['tail'][6] if len(sp['tail']) > 6 else None   # entry count on the device (plan.cu)
            _lib.check(lib.vqgnn_mp_fwd(
                _lib.ptr(iptr), _lib.ptr(icol), _lib.ptr(ival), None, _lib.ptr(icr), plan.small_chunk, innz, B, B,
                _lib.ptr(x), x.stride(0), None, _lib.ptr(bank.codes), _lib.ptr(bank.O), bank.nb, bank.M, bank.D,
                bank.Wp, None, 0, 0, 1.0, 1.0, _lib.ptr(y), y.stride(0), None, 0, None,
                _lib.ptr(_mp_ws(dev, innz, plan.small_chunk, C)), st))
            ws = torch.empty(int(lib.vqgnn_mp_fwd_tail_workspace_bytes(tnnz, TAIL_CHUNK, B, C)), dtype=torch.uint8,
                             device=dev)
            _lib.check(lib.vqgnn_mp_fwd_tail(
                _lib.ptr(tptr), _lib.ptr(tnode), _lib.ptr(tval), _lib.ptr(trval), _lib.ptr(tcr), TAIL_CHUNK, tnnz,
                _lib.ptr(tcount), B, _lib.ptr(x), x.stride(0), _lib.ptr(codes_g), codes_g.shape[1], _lib.ptr(bank.O), bank.nb,
                bank.M, bank.D, bank.Wp, float(wu), float(wu), _lib.ptr(y), y.stride(0), _lib.ptr(gq),
                gq.stride(0), _lib.ptr(info) if need_info else None, _lib.ptr(ws), st))
        elif (not v1 and need_info and plan.T > 0 and bank.D == 4 and bank.Wp == 8 and layer.split_info
              and (layer.split_info == 'force' or plan.nnz - plan.nnz_B >= INFO_SPLIT_MIN_ENTRIES)):
            # v2 training, large batch graph: the rows >= B (most of the entries) only feed the info_backward scalar.
            # Batch rows through the generic kernel (y), the rest through the slab-ordered SDDMM-shaped kernel whose
            # gathers stay L2-resident (csrc/mp_info.cu).
            slab = INFO_SLAB
            plan.extras['split_fwd'] = True
            nslab = (C + slab - 1) // slab
            tfS = torch.empty(nslab, plan.T, slab, device=dev)
            tgS = torch.empty(nslab, plan.T, slab, device=dev)
            _lib.check(lib.vqgnn_tail_materialize_slab(
                _lib.ptr(plan.tail_node), plan.T, _lib.ptr(bank.codes), _lib.ptr(bank.O), bank.nb, bank.M, bank.D,
                bank.Wp, slab, _lib.ptr(tfS), _lib.ptr(tgS), st))
            nB = plan.nnz_B
            _lib.check(lib.vqgnn_mp_fwd(
                _lib.ptr(plan.fwd_rowptr), _lib.ptr(plan.fwd_col), _lib.ptr(plan.fwd_val), None,
                _lib.ptr(plan.chunk_rows('fwdB')), MP_CHUNK, nB, B, B, _lib.ptr(x), x.stride(0),
                _lib.ptr(plan.tail_node), _lib.ptr(bank.codes), _lib.ptr(bank.O), bank.nb, bank.M, bank.D, bank.Wp,
                _lib.ptr(tfS), plan.T, slab, 1.0, float(wu), _lib.ptr(y), y.stride(0), None, 0, None,
                _lib.ptr(_mp_ws(dev, nB, MP_CHUNK, C)), st))
            ctx.tail_grad, ctx.tail_slab = tgS, slab      # the backward reads the gradient codewords slab-major
            iws = torch.empty(int(lib.vqgnn_mp_info_workspace_bytes(plan.nnz, C, slab)), dtype=torch.uint8, device=dev)
            _lib.check(lib.vqgnn_mp_info(
                _lib.ptr(plan.entry_rows()), _lib.ptr(plan.fwd_col), _lib.ptr(plan.fwd_val), nB, plan.nnz, B, plan.R,
                _lib.ptr(x), x.stride(0), _lib.ptr(tfS), _lib.ptr(tgS), C, slab, float(wu), _lib.ptr(info),
                _lib.ptr(iws), st))
        else:
            tail_feat = None
            if layer.wants_dense_tail(plan):
                # v2: every tail node is referenced by many edges -- gather its codewords once (both halves; the
                # gradient half is kept for the backward) and let the kernels read coalesced dense rows
                pre = layer._prefetched
                if pre is not None and pre[0] is plan:        # issued ahead on the side stream (LowRankGNN.forward)
                    _, tail_feat, tail_grad, ev = pre
                    layer._prefetched = None
                    torch.cuda.current_stream(dev).wait_event(ev)
                else:
                    tail_feat, tail_grad = layer.materialize_tail_rows(plan, need_info)
                ctx.tail_grad = tail_grad
            xr = _rows_operand(x, tail_feat) if (tail_feat is not None and not v1 and USE_ROWS_KERNEL) else x
            if (tail_feat is not None and not v1 and _rows_kernel_ok(xr, tail_feat, C)
                    and (tail_grad is not None or not need_info)):
                # materialised rows: lean asynchronous row gathers (csrc/mp_rows.cuh)
                _lib.check(lib.vqgnn_mp_fwd_rows(
                    _lib.ptr(plan.fwd_rowptr), _lib.ptr(plan.fwd_col), _lib.ptr(plan.fwd_val),
                    _lib.ptr(plan.chunk_rows('fwd')), MP_CHUNK, plan.nnz, plan.R, B, _lib.ptr(xr), xr.stride(0),
                    _lib.ptr(tail_feat), plan.T, 1.0, None, _lib.ptr(tail_grad), C, C, float(wu), _lib.ptr(y), y.stride(0),
                    _lib.ptr(info) if need_info else None, _lib.ptr(_mp_ws(dev, plan.nnz, MP_CHUNK, C)), st))
            else:
                _lib.check(lib.vqgnn_mp_fwd(
                    _lib.ptr(plan.fwd_rowptr), _lib.ptr(plan.fwd_col), _lib.ptr(plan.fwd_val), _lib.ptr(plan.fwd_rval),
                    _lib.ptr(plan.chunk_rows('fwd')), MP_CHUNK, plan.nnz, plan.R, B, _lib.ptr(x), x.stride(0),
                    _lib.ptr(plan.tail_node), _lib.ptr(bank.codes),
                    _lib.ptr(bank.O), bank.nb, bank.M, bank.D, bank.Wp, _lib.ptr(tail_feat), C, 0,
                    float(wu) if v1 else 1.0, float(wu),
                    _lib.ptr(y), y.stride(0), _lib.ptr(gq), gq.stride(0) if gq is not None else 0,
                    _lib.ptr(info) if need_info else None, _lib.ptr(_mp_ws(dev, plan.nnz, MP_CHUNK, C)), st))
        ctx.layer, ctx.plan, ctx.wu, ctx.fire_hook = layer, plan, float(wu), fire_hook
        if not hasattr(ctx, 'tail_grad'):
            ctx.tail_grad = None
        if not hasattr(ctx, 'tail_slab'):
            ctx.tail_slab = 0
        ctx.save_for_backward(x, gq)
        return y, info

    @staticmethod
    def backward(ctx, dy: Tensor, dinfo: Tensor):
        x, gq = ctx.saved_tensors
        layer, plan, wu = ctx.layer, ctx.plan, ctx.wu
        lib, st = _lib.load(), _lib.stream()
        bank = layer.bank
        dy = dy.contiguous()
        dinfo = dinfo.contiguous().float()
        B, C = x.shape
        dx = None
        v1 = plan.version == 'v1'
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B, C, device=x.device)
            nnz_t = int(plan.bwd_col.numel())
            rows_cand = not v1 and gq is None and ctx.tail_grad is not None and not ctx.tail_slab and USE_ROWS_KERNEL
            dyr = _rows_operand(dy, ctx.tail_grad) if rows_cand else dy
            if rows_cand and _rows_kernel_ok(dyr, ctx.tail_grad, C):
                # v2: the same lean row-gather kernel over the transposed CSR (dY rows / gradient codeword rows)
                _lib.check(lib.vqgnn_mp_fwd_rows(
                    _lib.ptr(plan.bwd_rowptr), _lib.ptr(plan.bwd_col), _lib.ptr(plan.bwd_val),
                    _lib.ptr(plan.chunk_rows('bwd')), plan.small_chunk, nnz_t, B, B, _lib.ptr(dyr), dyr.stride(0),
                    _lib.ptr(ctx.tail_grad), plan.T, float(wu), _lib.ptr(dinfo), None, C, C, 1.0, _lib.ptr(dx),
                    dx.stride(0),
                    None, _lib.ptr(_mp_ws(x.device, nnz_t, plan.small_chunk, C)), st))
            else:
                _lib.check(lib.vqgnn_mp_bwd(
                    _lib.ptr(plan.bwd_rowptr), _lib.ptr(plan.bwd_col), _lib.ptr(plan.bwd_val),
                    _lib.ptr(plan.chunk_rows('bwd')), plan.small_chunk, int(plan.bwd_col.numel()), B, _lib.ptr(dy),
                    dy.stride(0), _lib.ptr(plan.tail_node), _lib.ptr(bank.codes), _lib.ptr(bank.O), bank.nb,
                    bank.M, bank.D, bank.Wp, _lib.ptr(ctx.tail_grad), plan.T if ctx.tail_slab else C, ctx.tail_slab,
                    0.0 if v1 else wu, _lib.ptr(gq), gq.stride(0) if gq is not None else 0,
                    wu, _lib.ptr(dinfo), _lib.ptr(dx), dx.stride(0),
                    _lib.ptr(_mp_ws(x.device, nnz_t, plan.small_chunk, C)), st))
        if ctx.fire_hook:
            # the reference's hook(grad): vq.update(X_B, grad) ; c_indices[batch] = idx ; return grad
            bank.update(x, dy, plan.batch_idx)
        return dx, None, None, None, None, None


def plain_propagate(x: Tensor, adj, att_l: Optional[Tensor], att_r: Optional[Tensor],
                    negative_slope: float = 0.2) -> Tensor:
    """`conv.forward(x, adj)` of the reference on an explicit adjacency with no codeword rows (forward only):
    GCN/SAGE  out = adj @ x                                                        (convs.py:65-101)
    GAT       out[i] = sum_j adj[i,j] exp(lrelu((a_l[j] + a_r[i]) / sigma)) x[j]   (convs.py:165-266), all columns
              of x taking part in the scores (att_* of length C) -- the un-normalised sum the reference returns."""
    _lib.require_device(x)
    n = x.shape[0]
    rowptr, col, val = adj.csr()
    assert int(adj.sparse_sizes()[0]) == n
    C = x.shape[1]
    lib, st = _lib.load(), _lib.stream()
    xc = x.detach().contiguous().float()
    y = torch.empty_like(xc)
    codes = torch.zeros(1, 1, dtype=torch.int16, device=x.device)
    O = torch.zeros(1, 1, 8, device=x.device)
    D = 4 if C % 4 == 0 else 1
    rowptr, col, val = rowptr.to(torch.int32), col.to(torch.int32), val.float().contiguous()
    nnz = int(col.numel())
    chunks = torch.empty(max(int(lib.vqgnn_mp_num_chunks(nnz, MP_CHUNK)), 1), dtype=torch.int32, device=x.device)
    _lib.check(lib.vqgnn_mp_chunk_rows(_lib.ptr(rowptr), n, nnz, MP_CHUNK, _lib.ptr(chunks), st))
    if att_l is None and _rows_kernel_ok(xc, None, C):     # wide rows: the lean row-gather kernel (no codeword rows)
        _lib.check(lib.vqgnn_mp_fwd_rows(
            _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), _lib.ptr(chunks), MP_CHUNK, nnz, n, n, _lib.ptr(xc),
            xc.stride(0), None, 0, 1.0, None, None, 0, C, 1.0, _lib.ptr(y), y.stride(0), None,
            _lib.ptr(_mp_ws(x.device, nnz, MP_CHUNK, C)), st))
        return y
    if att_l is None:
        _lib.check(lib.vqgnn_mp_fwd(
            _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), None, _lib.ptr(chunks), MP_CHUNK, nnz,
            n, n, _lib.ptr(xc), xc.stride(0), None, _lib.ptr(codes), _lib.ptr(O), C // D, 1, D, 8, None, 0, 0, 1.0, 1.0,
            _lib.ptr(y), y.stride(0), None, 0, None, _lib.ptr(_mp_ws(x.device, nnz, MP_CHUNK, C)), st))
        return y
    # GAT: the fused kernels carry an implicit ones column (coefficient att[C]); a zero coefficient removes it from
    # the scores, and multiplying the normalised output back by its denominator gives the reference's plain sum
    zero = torch.zeros(1, device=x.device)
    al = torch.cat([att_l.detach().reshape(-1).float(), zero])
    ar = torch.cat([att_r.detach().reshape(-1).float(), zero])
    assert al.numel() == C + 1, "att_l / att_r must have one entry per column of x"
    a_l, a_r, stat = torch.empty(n, device=x.device), torch.empty(n, device=x.device), torch.empty(2, device=x.device)
    den = torch.empty(n, device=x.device)
    _lib.check(lib.vqgnn_gat_scores(n, n, _lib.ptr(xc), xc.stride(0), None, _lib.ptr(codes), _lib.ptr(O), C // D, 1, D,
                                    8, None, 0, _lib.ptr(al), _lib.ptr(ar), _lib.ptr(a_l), _lib.ptr(a_r),
                                    _lib.ptr(stat), st))
    _lib.check(lib.vqgnn_gat_fwd(_lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), _lib.ptr(chunks), MP_CHUNK, nnz, n, n,
                                 _lib.ptr(xc), xc.stride(0), None, _lib.ptr(codes), _lib.ptr(O), C // D, 1, D, 8,
                                 None, 0, _lib.ptr(a_l), _lib.ptr(a_r), _lib.ptr(stat), float(negative_slope), 1.0,
                                 _lib.ptr(y), y.stride(0), _lib.ptr(den), None, None, st))
    return y * (den + 1e-16).unsqueeze(1)


# --------------------------------------------------------------------------------------------------
# modules
# --------------------------------------------------------------------------------------------------
class LowRankGNNBlock(nn.Module):
    """Per-branch state holder (vq_gnn_v2/models.py:11-63; v1 adds the per-branch conv, :42-49)."""

    def __init__(self, in_channels, hidden_channels, num_M, num_D, num_N, num_branch, cluster, kmeans_iter,
                 EMA_flag, kmeans_init, use_gcn, commitment_cost, grad_normalize_scale, hook_flag, warm_up_flag,
                 momentum, conv_type, transformer_flag, version: str = 'v2'):
        super().__init__()
        self.num_M, self.num_D, self.num_N, self.EMA_flag = num_M, num_D, num_N, EMA_flag
        self.commitment_cost, self.hook_flag = commitment_cost, hook_flag
        self.grad_normalize_scale, self.conv_type = grad_normalize_scale, conv_type
        self.transformer_flag, self.version = transformer_flag, version
        if transformer_flag:
            raise NotImplementedError("--transformer-flag is outside the hot path (SURVEY.md §2 #6)")
        if not EMA_flag:
            raise ValueError('Not EMA vq not studied')                     # v1/models.py:58-59
        c = torch.randint(0, num_M, (num_N,), dtype=torch.short)           # models.py:27
        self.register_buffer('c_indices', c)
        add_flag = False
        if version == 'v1':                                                # v1/models.py:42-53
            if conv_type != 'GAT':
                self.conv = OurGCNConv(in_channels, in_channels, normalize=False)
            else:
                self.conv = OurGATConv(in_channels + 1, in_channels + 1, bias=False, add_self_loops=False,
                                       version='v1')
            add_flag = conv_type == 'GAT'
        self.vq = VectorQuantizerEMA(num_M, num_D, commitment_cost=commitment_cost,
                                     grad_normalize_scale=grad_normalize_scale, warm_up_flag=warm_up_flag,
                                     momentum=momentum, add_flag=add_flag)
        self.kmeans_init = kmeans_init
        self.grad_kmeans_init = kmeans_init
        self.inited = False
        if version == 'v1':
            self.ln = nn.LayerNorm(in_channels, elementwise_affine=False)  # v1/models.py:65 (unused)

    def init(self, X_B, batch_indices):
        """models.py:61-63 for a single branch (the layer normally initialises all branches at once)."""
        bank, i = self.vq.bank, self.vq._branch
        x = X_B.detach().contiguous().float()
        idx = bank.run(x, None, batch_indices.to(torch.int32), self.vq.training, k0=i, nbc=1, local=True)
        return idx

    def hook(self, grad):
        raise RuntimeError("the per-branch hook is fused into LowRankGNNLayer's backward "
                           "(VQConvFunction.backward); it is not called directly")


class LowRankGNNLayer(nn.Module):
    def __init__(self, in_channels, out_channels, dropout,
                 num_M, num_D, num_N, num_branch, cluster, ln_para, no_second_fc,
                 kmeans_iter, EMA_flag, split, kmeans_init, dropbranch, skip, use_gcn, commitment_cost,
                 grad_normalize_scale, hook, weight_ahead, warm_up_flag, momentum, conv_type, transformer_flag,
                 version: str = 'v2', literal_v2_hooks: bool = False):
        super().__init__()
        self.weight_ahead = weight_ahead
        if self.weight_ahead:
            if out_channels % num_D != 0:
                raise ValueError('Cannot fully split')
            self.num_branch = int(out_channels / num_D)
        else:
            if in_channels % num_D != 0:
                raise ValueError('Cannot fully split')
            self.num_branch = int(in_channels / num_D)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.no_second_fc, self.EMA_flag, self.split = no_second_fc, EMA_flag, split
        self.num_D, self.num_M, self.num_N = num_D, num_M, num_N
        self.dropbranch, self.skip = dropbranch, skip
        self.conv_type, self.transformer_flag = conv_type, transformer_flag
        self.version, self.literal_v2_hooks = version, literal_v2_hooks
        self.hook_flag = hook
        if dropbranch and dropbranch > 0:
            raise NotImplementedError("dropbranch > 0 is unused by every reference configuration")
        if transformer_flag:
            raise NotImplementedError("--transformer-flag is outside the hot path (SURVEY.md §2 #6)")
        if conv_type not in ('GCN', 'SAGE', 'GAT'):
            raise ValueError('GNN conv type not supported')

        if version == 'v2':                                               # v2/models.py:93-97
            if conv_type != 'GAT':
                self.conv = OurGCNConv(in_channels, in_channels, normalize=False)
            else:
                self.conv = OurGATConv(in_channels + 1, in_channels + 1, bias=False, add_self_loops=False)
        self.linear_k, self.linear_v = nn.ModuleList(), nn.ModuleList()
        self.gnn_block, self.transformer_block = nn.ModuleList(), nn.ModuleList()
        for _ in range(self.num_branch):
            if no_second_fc:
                self.gnn_block.append(LowRankGNNBlock(
                    None if version == 'v2' else num_D, None if version == 'v2' else out_channels,
                    num_M, num_D, num_N, self.num_branch, cluster, kmeans_iter, EMA_flag, kmeans_init,
                    use_gcn, commitment_cost, grad_normalize_scale, hook, warm_up_flag, momentum, conv_type,
                    False, version=version))
            else:
                raise ValueError('second fc not studied')
        if self.skip:
            self.linear_skip = nn.Linear(in_channels, out_channels)
        self.gnn_transform = nn.Linear(in_channels, out_channels)
        self.batch_norm = nn.BatchNorm1d(out_channels, affine=False)
        if self.conv_type == 'SAGE':
            self.fc_sage = nn.Linear(in_channels, out_channels)

        add_flag = (version == 'v1' and conv_type == 'GAT')
        bank = VQBank(self.num_branch, num_M, num_D, grad_normalize_scale=grad_normalize_scale,
                      warm_up_flag=warm_up_flag, momentum=momentum, add_flag=add_flag, num_N=num_N)
        object.__setattr__(self, 'bank', bank)
        self.sync_status = False
        # shared-memory codebook kernel for the v1 out-of-batch entries: True = when the average row is long
        # enough (graph.TAIL_MIN_AVG_DEGREE), 'force' = whenever the shape allows, False = never
        self.use_tail_kernel = True
        # v2: gather the out-of-batch nodes' codewords once per step into dense rows (vqgnn_tail_materialize).
        # True = when tail nodes are referenced >= 4x on average (measured per layer: arxiv 1.61 -> 1.50 ms, collab
        # 1.59 -> 1.49, products forward 2.09 -> 1.17 + 0.18 ms), 'force' = always, False = never
        self.materialize_tail = True
        # also materialise the gradient codewords (read by the backward's ~B*deg out-of-batch entries); False: the
        # backward gathers them per entry from the codebook instead
        self.materialize_grad = _os.environ.get('VQGNN_MATERIALIZE_GRAD', '1') != '0'
        # v2 training: batch rows and out-of-batch rows (info_backward only) through separate kernels (csrc/mp_info.cu:
        # slab-major tables, L2-resident slices).  True = when the latter hold >= INFO_SPLIT_MIN_ENTRIES entries,
        # 'force' = always, False (default) = never: measured at the products shape both forms run at the same
        # ~5.5 TB/s of on-chip gather traffic (1.50 vs 1.60 ms per layer) and the split pays a second small launch for
        # the batch rows (step 9.2 vs 8.5 ms), so the one-kernel forward stays the default
        self.split_info = False
        self._prefetched = None
        self._restack()

    # ---- dense copies of the out-of-batch codewords (v2) -----------------------------------------
    def wants_dense_tail(self, plan: BatchPlan) -> bool:
        bank = self.bank
        if plan.version != 'v2' or plan.T == 0 or bank.D != 4 or bank.Wp != 8 or self.conv_type == 'GAT':
            return False
        return self.materialize_tail == 'force' or bool(self.materialize_tail and plan.nnz >= 4 * plan.T)

    def materialize_tail_rows(self, plan: BatchPlan, with_grad: bool, out=None):
        """tail_feat / tail_grad [T, C]: the feature / gradient codewords of every out-of-batch node (vqgnn_tail_materialize)
        on the current stream (into `out` = (tail_feat, tail_grad) when given)."""
        bank, dev = self.bank, plan.device
        C = bank.nb * bank.D
        if out is None:
            tail_feat = _tail_table(plan, C)
            tail_grad = _tail_table(plan, C) if (with_grad and self.materialize_grad) else None
        else:
            tail_feat, tail_grad = out
        _lib.check(_lib.load().vqgnn_tail_materialize(
            _lib.ptr(plan.tail_node), plan.T, _lib.ptr(bank.codes), _lib.ptr(bank.O), bank.nb, bank.M, bank.D,
            bank.Wp, _lib.ptr(tail_feat), _lib.ptr(tail_grad), C, _lib.stream()))
        return tail_feat, tail_grad

    def prefetch_tail_rows(self, plan: BatchPlan, side: "torch.cuda.Stream") -> None:
        """Issue this layer's materialisation on `side` now (it depends only on the quantiser state and the plan, not
        on the layer input), to be picked up by the layer's forward: the HBM-write-bound copy then overlaps the
        gather-bound forward kernels of the layers before it."""
        self._prefetched = None
        if not (self.training and plan.training and self.inited and self.wants_dense_tail(plan)
                and not (self.split_info and self.split_info == 'force')):
            return
        self.bank.join()                                   # a pending side-stream VQ update writes codes / O
        # the buffers are allocated on the CONSUMER's stream (its allocator pool owns them; the side stream's writes are
        # ordered before the consumer's reads by the event), so no cross-stream record_stream bookkeeping is needed
        C = self.bank.nb * self.bank.D
        tf = _tail_table(plan, C)
        tg = _tail_table(plan, C) if self.materialize_grad else None
        side.wait_stream(torch.cuda.current_stream(plan.device))
        with torch.cuda.stream(side):
            self.materialize_tail_rows(plan, True, out=(tf, tg))
            ev = torch.cuda.Event()
            ev.record(side)
        self._prefetched = (plan, tf, tg, ev)

    # ---- stacked storage <-> per-branch reference buffers -------------------------------------
    def _restack(self):
        """(Re)build the stacked bank from the per-branch buffers and alias them as views of it.
        Runs after construction and after every `.to()` / `.cuda()`."""
        blocks = list(self.gnn_block)
        dev = blocks[0].c_indices.device
        bank = self.bank
        if bank.device != dev:
            bank.to(dev)
        for i, b in enumerate(blocks):
            b.vq._pull_into_bank(bank, i)
            bank.codes[:, i] = b.c_indices
        for i, b in enumerate(blocks):
            b.vq._owns_bank = False
            b.vq._bind_views(bank, i)
            b._buffers['c_indices'] = bank.codes[:, i]
        bank.mark_codes_dirty()

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._restack()
        return out

    def _load_from_state_dict(self, *a, **k):
        self.bank.mark_codes_dirty()
        return super()._load_from_state_dict(*a, **k)   # buffers are views: copy_ writes through

    # ---- reference bookkeeping ----------------------------------------------------------------
    @property
    def inited(self) -> bool:
        return all(b.inited for b in self.gnn_block)

    def set_inited(self, flag: bool = True):
        for b in self.gnn_block:
            b.inited = flag
            b.kmeans_init = False
            b.grad_kmeans_init = False

    def check_status(self):
        self.bank.check_status()

    # ---- forward ------------------------------------------------------------------------------
    def forward(self, x, batch_A, warm_up_rate, unlabeled):
        nb = self.num_branch
        errors, X_B_norms, quantized_norms = [0] * nb, [0] * nb, [0] * nb
        losses, info_backwards, hookeds = 0, 0, []
        _lib.require_device(x)
        plan = build_plan(batch_A, self.conv_type, self.num_N, self.training, x.device)
        if plan.training != self.training:
            raise ValueError("BatchPlan was built for a different train/eval mode")
        x = x.float()
        xc = x if x.is_contiguous() else x.contiguous()
        inited = self.inited
        self.bank.join()             # a pending side-stream update of this layer's quantisers (bank.async_update)
        self.bank.poll_status()      # deferred 'Bad Init!' check of an earlier update: no stream stall
        do_init = (not inited or unlabeled) and (self.training or self.version == 'v2')
        if do_init:          # models.py:165-166 (v2) / v1 models.py:164-165: feature-only warm start
            self.bank.run(xc.detach(), None, plan.batch_idx, self.training)
            if self.sync_status:
                self.bank.check_status()
        fire = (inited and self.training and not unlabeled and self.hook_flag
                and not (self.version == 'v2' and self.literal_v2_hooks) and torch.is_grad_enabled())
        if self.conv_type == 'GAT':
            from .gat import gat_conv
            y, info = gat_conv(self, xc, plan, float(warm_up_rate), fire)
        else:
            y, info = VQConvFunction.apply(xc, _trigger(xc.device) if fire else None, self, plan,
                                           float(warm_up_rate), fire)
        if self.training:
            info_backwards = info_backwards + info                      # models.py:199-200
        out = _linear(self.gnn_transform, y)                            # models.py:202
        if self.conv_type == 'SAGE':
            out = out + _linear(self.fc_sage, x)                        # :203-204
        if self.skip:
            out = out + _linear(self.linear_skip, x)                    # :228-229
        return out, errors, X_B_norms, quantized_norms, losses, info_backwards, hookeds


class LowRankGNN(nn.Module):
    def __init__(self, in_channels, hidden_channels, out_channels, num_layers, dropout,
                 num_M, num_D, num_N, num_branch=0, cluster='vq', ln_para=True, no_second_fc=False,
                 kmeans_iter=100, EMA_flag=True, split=True, kmeans_init=False, dropbranch=0, skip=True,
                 use_gcn=False, commitment_cost=0.5, grad_scale=(1, 1), act='relu', weight_ahead=False,
                 bn_flag=False, warm_up_flag=False, momentum=0.1, conv_type='GCN', transformer_flag=False,
                 alpha_dropout_flag=False, version: str = 'v2', literal_v2_hooks: bool = False):
        super().__init__()
        self.num_layers, self.skip, self.dropout = num_layers, skip, dropout
        self.bn_flag, self.alpha_dropout_flag = bn_flag, alpha_dropout_flag
        self.version, self.conv_type, self.num_N = version, conv_type, num_N
        self.prefetch_tails = True      # v2 training: see LowRankGNNLayer.prefetch_tail_rows
        if self.alpha_dropout_flag:
            self.alpha_dropout = nn.AlphaDropout(p=self.dropout)
        self.convs, self.batch_norms = nn.ModuleList(), nn.ModuleList()

        def mk(cin, cout, db):
            return LowRankGNNLayer(cin, cout, dropout, num_M, num_D, num_N, num_branch=num_branch,
                                   cluster=cluster, ln_para=ln_para, no_second_fc=no_second_fc,
                                   kmeans_iter=kmeans_iter, EMA_flag=EMA_flag, split=split,
                                   kmeans_init=kmeans_init, dropbranch=db, skip=skip, use_gcn=use_gcn,
                                   commitment_cost=commitment_cost, grad_normalize_scale=grad_scale, hook=True,
                                   weight_ahead=weight_ahead, warm_up_flag=warm_up_flag, momentum=momentum,
                                   conv_type=conv_type, transformer_flag=transformer_flag, version=version,
                                   literal_v2_hooks=literal_v2_hooks)
        self.convs.append(mk(in_channels, hidden_channels, 0))
        self.batch_norms.append(nn.BatchNorm1d(hidden_channels, affine=False))
        for _ in range(num_layers - 2):
            self.convs.append(mk(hidden_channels, hidden_channels, dropbranch))
            self.batch_norms.append(nn.BatchNorm1d(hidden_channels, affine=False))
        self.convs.append(mk(hidden_channels, out_channels, dropbranch))
        self.transform = nn.Linear(out_channels, out_channels)
        self.ln = nn.LayerNorm(hidden_channels, elementwise_affine=False)
        if act == 'relu':
            self.act_f = F.relu
        elif act == 'elu':
            self.act_f = F.elu
        elif act == 'leaky_gelu':
            self.act_f = lambda x: 0.1 * x + 0.9 * F.gelu(x)
        else:
            raise ValueError('Activation not supported!')

    def prepare(self, batch_A, device=None) -> BatchPlan:
        """Build the kernel plan once per mini-batch (shared by all layers)."""
        device = device or next(self.parameters()).device
        plan = build_plan(batch_A, self.conv_type, self.num_N, self.training, device)
        return plan.warm() if plan.device.type == 'cuda' else plan

    def prepare_from_graph(self, graph, node_idx: Tensor, recovery_flag: bool = True) -> BatchPlan:
        """Mini-batch plan straight from the device-resident normalised graph (`synth.Graph` / any object with
        rowptr, col, val[, deg, deg_inv]) and the batch's node ids: the device-side equivalent of the reference's
        loader + `prepare_batch_input` (vq_gnn_v2/dataloader.py:98-148 + utils/misc.py:57-75; v1:
        vq_gnn_v1/utils/dataloader.py:64-86) without the int64 COO round trip through host memory -- only the
        node ids cross PCIe."""
        from . import graph as G
        if self.version == 'v2':
            plan = G.plan_from_graph_v2(graph, node_idx, self.conv_type, self.training)
            return plan.warm()
        batch_A = G.batch_from_graph_v1(graph, node_idx, train_flag=self.training, recovery_flag=recovery_flag)
        return self.prepare(batch_A, device=graph.col.device)

    def forward(self, batch, warm_up_rate=1, unlabeled=False):
        losses_full, info_backwards_full = 0, 0
        errors_full, X_B_norms_full, quantized_norms_full = [], [], []
        x, batch_A = batch
        batch_A = build_plan(batch_A, self.conv_type, self.num_N, self.training, x.device)
        if (self.prefetch_tails and self.training and not unlabeled and x.is_cuda and self.version == 'v2'
                and len(self.convs) > 1):
            # layers 2..L: materialise their out-of-batch codeword rows on a side stream while layer 1 runs
            side = VQBank.side_stream(x.device, 1)
            for conv in self.convs[1:]:
                conv.prefetch_tail_rows(batch_A, side)
        for i, conv in enumerate(self.convs[:-1]):
            x, errors, X_B_norms, quantized_norms, losses, info_backwards, _ = \
                conv(x, batch_A, warm_up_rate, unlabeled)
            if self.bn_flag:
                x = self.batch_norms[i](x)
            x = self.act_f(x)
            if self.alpha_dropout_flag:
                x = self.alpha_dropout(x)
            else:
                x = F.dropout(x, p=self.dropout, training=self.training)
            losses_full += losses
            info_backwards_full += info_backwards
            errors_full.append(errors), X_B_norms_full.append(X_B_norms)
            quantized_norms_full.append(quantized_norms)
        x, errors, X_B_norms, quantized_norms, losses, info_backwards, _ = \
            self.convs[-1](x, batch_A, warm_up_rate, unlabeled)
        losses_full += losses
        info_backwards_full += info_backwards
        errors_full.append(errors), X_B_norms_full.append(X_B_norms)
        quantized_norms_full.append(quantized_norms)
        self.errors, self.X_B_norms, self.quantized_norms = errors_full, X_B_norms_full, quantized_norms_full
        return x, losses_full, info_backwards_full

    def init(self, batch, layer_idx):
        """Codebook warm start (models.py:370-374)."""
        x, batch_A = batch
        batch_A = build_plan(batch_A, self.conv_type, self.num_N, self.training, x.device)
        for i, conv in enumerate(self.convs[:layer_idx]):
            x, _, _, _, _, _, _ = conv(x, batch_A, 1, False)
            x = self.act_f(x)

    @torch.no_grad()
    def warm_start(self, graph, X: Tensor, batch_size: int, node_order: Optional[Tensor] = None) -> int:
        """The reference's `init(data, model, device, test_loader)` (vq_gnn_v2/main_node.py:17-37, v1 alike) as a
        streaming pass over the device-resident graph: for layer_idx = 1..L, every node batch of the TEST loader
        (consecutive ranges of `node_order`, batch rows only) runs `model.init(batch, layer_idx)` -- a feature-only
        codebook update of each visited layer -- then every block is marked inited.  L(L+1)/2 layer passes over the
        whole graph, all on the device: node ids never leave it.  Returns the number of batches per pass."""
        from . import graph as G
        was_training = self.training
        self.train()
        dev = graph.col.device
        N = int(graph.N)
        order = torch.arange(N, device=dev) if node_order is None else node_order.to(dev)
        nb = 0
        # every rank streams the WHOLE graph and (the update being deterministic) arrives at identical replicas:
        # no collective is needed, and summing identical statistics over ranks would only rescale them
        dist_flags = [layer.bank.distributed for layer in self.convs]
        for layer in self.convs:
            layer.bank.distributed = False
        for layer_idx in range(1, self.num_layers + 1):
            nb = 0
            for lo in range(0, N, batch_size):
                ids = order[lo:lo + batch_size]
                if self.version == 'v2':
                    plan = G.plan_from_graph_v2(graph, ids, self.conv_type, True, batch_rows_only=True)
                else:
                    bA = G.batch_from_graph_v1(graph, ids, train_flag=False, recovery_flag=False)
                    plan = build_plan(bA, self.conv_type, self.num_N, True, dev)
                self.init((X[ids], plan), layer_idx)
                nb += 1
        for layer, flag in zip(self.convs, dist_flags):
            layer.bank.distributed = flag
        self.set_inited(True)
        self.train(was_training)
        return nb

    def set_inited(self, flag: bool = True):
        """What main's init() does after the warm-start passes (main_node.py:30-37)."""
        for layer in self.convs:
            layer.set_inited(flag)

    def check_status(self):
        for layer in self.convs:
            layer.check_status()

    def set_async_vq_updates(self, flag: bool = True, per_layer_streams: bool = False):
        """Run the hook's VQ updates on a side stream (VQBank.async_update): they only prepare the NEXT step.
        per_layer_streams: one side stream per layer, so the three updates of a step overlap each other as well (multi-GPU:
        give every layer's bank its own `process_group` first)."""
        for li, layer in enumerate(self.convs):
            layer.bank.async_update = bool(flag)
            layer.bank.side_lane = li if per_layer_streams else 0

    def join_vq_updates(self):
        """Order the current stream after every pending side-stream VQ update.  Each layer's next forward does this by
        itself; call it explicitly before reading quantiser state or before a CUDA-graph capture of the step ends."""
        for layer in self.convs:
            layer.bank.join()

    def inference(self, x, A):
        raise NotImplementedError("LowRankGNN.inference is broken in the reference v2 (models.py:355) "
                                  "and is not part of the hot path")
