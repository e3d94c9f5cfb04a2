"""GAT message passing of the VQ layers (v2 "B+B'" formulation) as one autograd Function over libvqgnn.

Reference: `LowRankGNNLayer.forward` with `OurGATConv` (vq_gnn_v2/models.py:161-198, vq_gnn_v2/convs.py:165-266,
vq_gnn_v2/utils/vq_softmax.py:41-57).  The reference concatenates a ones column, runs PyG's per-edge
`message` (materialising nnz x (C+1) messages) and divides by the aggregated ones column afterwards; here
the codeword gather, the edge weights, the denominator and the normalisation are fused into
`vqgnn_gat_scores` + `vqgnn_gat_fwd`, and the whole backward (including the gradient through the max-based
"Trick 1" scale, which the reference leaves inside the autograd graph) is `vqgnn_gat_bwd`.
"""
from __future__ import annotations

import torch

from . import _lib
from .graph import MP_CHUNK, BatchPlan

Tensor = torch.Tensor


class VQGATFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, att_l: Tensor, att_r: Tensor, layer, plan: BatchPlan, wu: float,
                fire_hook: bool, slope: float):
        _lib.require_device(x)
        lib, st = _lib.load(), _lib.stream()
        bank = layer.bank
        B, C = x.shape
        R, dev = plan.R, x.device
        # scores cover ALL B + B' nodes of x_input in both modes: the eval plan keeps only the batch ROWS (R = B)
        # but its columns still reach the out-of-batch nodes, and Trick 1's max runs over every row of x_input
        # (vq_gnn_v2/convs.py:188-211)
        Rs = B + plan.T
        al, ar = att_l.detach().reshape(-1).contiguous(), att_r.detach().reshape(-1).contiguous()
        assert al.numel() == C + 1 and ar.numel() == C + 1
        a_l, a_r = torch.empty(Rs, device=dev), torch.empty(Rs, device=dev)
        stat = torch.empty(2, device=dev)
        tail_feat = tail_grad = None
        auto = plan.nnz >= 4 * max(plan.T, 1)
        if (plan.T > 0 and bank.D == 4 and bank.Wp == 8
                and (layer.materialize_tail == 'force' or (layer.materialize_tail and auto))):
            # every out-of-batch node's codewords gathered once into dense rows (see models.VQConvFunction)
            from .models import _tail_table           # [T, C] with B spare rows in front (see models._tail_table)
            tail_feat = _tail_table(plan, C)
            tail_grad = _tail_table(plan, C) if plan.training else None     # its spare rows receive dYn in the backward
            _lib.check(lib.vqgnn_tail_materialize(
                _lib.ptr(plan.tail_node), plan.T, _lib.ptr(bank.codes), _lib.ptr(bank.O), bank.nb, bank.M, bank.D,
                bank.Wp, _lib.ptr(tail_feat), _lib.ptr(tail_grad), C, st))
        ctx.tail = (tail_feat, tail_grad)
        _lib.check(lib.vqgnn_gat_scores(
            Rs, B, _lib.ptr(x), x.stride(0), _lib.ptr(plan.tail_node), _lib.ptr(bank.codes), _lib.ptr(bank.O),
            bank.nb, bank.M, bank.D, bank.Wp, _lib.ptr(tail_feat), C, _lib.ptr(al), _lib.ptr(ar), _lib.ptr(a_l),
            _lib.ptr(a_r), _lib.ptr(stat), st))
        y = torch.empty(B, C, device=dev)
        den = torch.empty(B, device=dev)
        need_info = plan.training
        info = torch.zeros((), device=dev)
        from .models import _rows_kernel_ok, _rows_operand
        xr = _rows_operand(x, tail_feat) if tail_feat is not None else x
        if (tail_feat is not None and (tail_grad is not None or not need_info) and _rows_kernel_ok(xr, tail_feat, C)
                and MP_CHUNK <= 256):
            # materialised rows: the lean row-gather kernel with per-entry GAT weights (csrc/mp_rows.cuh)
            nck = (plan.nnz + MP_CHUNK - 1) // MP_CHUNK
            wsb = 1024 + 8 * (nck * ((C + 127) // 128) // 4 + 2)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev) if need_info else None
            _lib.check(lib.vqgnn_gat_fwd_rows(
                _lib.ptr(plan.fwd_rowptr), _lib.ptr(plan.fwd_col), _lib.ptr(plan.fwd_val),
                _lib.ptr(plan.chunk_rows('fwd')), MP_CHUNK, plan.nnz, R, B, _lib.ptr(xr), xr.stride(0),
                _lib.ptr(tail_feat), plan.T, _lib.ptr(tail_grad), C, C, _lib.ptr(a_l), _lib.ptr(a_r), _lib.ptr(stat),
                float(slope), float(wu), _lib.ptr(y), y.stride(0), _lib.ptr(den),
                _lib.ptr(info) if need_info else None, _lib.ptr(ws), wsb if need_info else 0, st))
            ctx.layer, ctx.plan, ctx.wu, ctx.fire_hook, ctx.slope = layer, plan, float(wu), fire_hook, float(slope)
            ctx.att_shape = att_l.shape
            ctx.save_for_backward(x, al, ar, a_l, a_r, stat, y, den)
            return y, info
        ws = torch.empty(8, dtype=torch.float64, device=dev) if need_info else None
        _lib.check(lib.vqgnn_gat_fwd(
            _lib.ptr(plan.fwd_rowptr), _lib.ptr(plan.fwd_col), _lib.ptr(plan.fwd_val),
            _lib.ptr(plan.chunk_rows('fwd')), MP_CHUNK, plan.nnz, R, B, _lib.ptr(x), x.stride(0),
            _lib.ptr(plan.tail_node), _lib.ptr(bank.codes), _lib.ptr(bank.O), bank.nb, bank.M, bank.D, bank.Wp,
            _lib.ptr(tail_feat), C, _lib.ptr(a_l), _lib.ptr(a_r), _lib.ptr(stat), float(slope), float(wu),
            _lib.ptr(y), y.stride(0), _lib.ptr(den), _lib.ptr(info) if need_info else None, _lib.ptr(ws), st))
        ctx.layer, ctx.plan, ctx.wu, ctx.fire_hook, ctx.slope = layer, plan, float(wu), fire_hook, float(slope)
        ctx.att_shape = att_l.shape
        ctx.save_for_backward(x, al, ar, a_l, a_r, stat, y, den)
        return y, info

    @staticmethod
    def backward(ctx, dy: Tensor, dinfo: Tensor):
        x, al, ar, a_l, a_r, stat, y, den = ctx.saved_tensors
        layer, plan, wu = ctx.layer, ctx.plan, ctx.wu
        lib, st = _lib.load(), _lib.stream()
        bank = layer.bank
        B, C = x.shape
        R, dev = plan.R, x.device
        dy = dy.contiguous().float()
        dinfo = dinfo.contiguous().float()
        dyn = getattr(ctx.tail[1], 'head_rows', None) if ctx.tail[1] is not None else None
        if dyn is None or dyn.shape != x.shape:
            dyn = torch.empty(B, C, device=dev)
        dden = torch.empty(B, device=dev)
        ds_l, ds_r = torch.empty(R, device=dev), torch.empty(R, device=dev)
        datt_l, datt_r = torch.empty(C + 1, device=dev), torch.empty(C + 1, device=dev)
        dx = torch.empty(B, C, device=dev) if ctx.needs_input_grad[0] else None
        xo = x
        if ctx.tail[0] is not None:      # batch rows next to the codeword rows: one 32-bit-offset window (models._tail_table)
            from .models import _rows_operand
            xo = _rows_operand(x, ctx.tail[0])
        x_keep, x = x, xo
        _lib.check(lib.vqgnn_gat_bwd(
            _lib.ptr(plan.fwd_rowptr), _lib.ptr(plan.fwd_col), _lib.ptr(plan.fwd_val),
            _lib.ptr(plan.chunk_rows('fwd')), plan.nnz, R,
            _lib.ptr(plan.bwd_rowptr), _lib.ptr(plan.bwd_col), _lib.ptr(plan.bwd_val),
            _lib.ptr(plan.chunk_rows('bwd')), int(plan.bwd_col.numel()), plan.small_chunk, B, _lib.ptr(x), x.stride(0),
            _lib.ptr(plan.tail_node), _lib.ptr(bank.codes), _lib.ptr(bank.O), bank.nb, bank.M, bank.D, bank.Wp,
            _lib.ptr(ctx.tail[0]), _lib.ptr(ctx.tail[1]), C,
            _lib.ptr(al), _lib.ptr(ar), _lib.ptr(a_l), _lib.ptr(a_r), _lib.ptr(stat), ctx.slope,
            _lib.ptr(y), y.stride(0), _lib.ptr(den), _lib.ptr(dy), dy.stride(0), wu, _lib.ptr(dinfo),
            _lib.ptr(dyn), dyn.stride(0), _lib.ptr(dden), _lib.ptr(ds_l), _lib.ptr(ds_r),
            _lib.ptr(dx), dx.stride(0) if dx is not None else 0, _lib.ptr(datt_l), _lib.ptr(datt_r), st))
        if ctx.fire_hook:
            # the hook sees d loss / d (un-normalised conv output)[:B, :C]  (vq_gnn_v2/models.py:181-185)
            bank.update(x_keep, dyn, plan.batch_idx)
        return dx, datt_l.view(ctx.att_shape), datt_r.view(ctx.att_shape), None, None, None, None, None


class VQGAT1Function(torch.autograd.Function):
    """v1 per-branch GAT: `LowRankGNNBlock.forward` for all branches at once (vq_gnn_v1/models.py:143-233 with
    `mapper`, vq_gnn_v1/utils/dataloader.py:144-192).  att_l / att_r: [nb, 5] (the blocks' parameters stacked)."""

    @staticmethod
    def forward(ctx, x: Tensor, att_l: Tensor, att_r: Tensor, layer, plan: BatchPlan, wu: float,
                fire_hook: bool, slope: float):
        _lib.require_device(x)
        lib, st = _lib.load(), _lib.stream()
        bank = layer.bank
        B, C = x.shape
        nb, M, dev = bank.nb, bank.M, x.device
        al, ar = att_l.detach().contiguous(), att_r.detach().contiguous()
        a_l, a_r = torch.empty(B, nb, device=dev), torch.empty(B, nb, device=dev)
        cs, stat = torch.empty(nb, M, 2, device=dev), torch.empty(nb, 2, device=dev)
        _lib.check(lib.vqgnn_gat1_scores(B, _lib.ptr(x), x.stride(0), _lib.ptr(bank.O), nb, M, bank.D, bank.Wp,
                                         float(wu), _lib.ptr(al), _lib.ptr(ar), _lib.ptr(a_l), _lib.ptr(a_r),
                                         _lib.ptr(cs), _lib.ptr(stat), st))
        y, den = torch.empty(B, C, device=dev), torch.empty(B, nb, device=dev)
        need_info = plan.training and plan.has_rval
        info = torch.zeros((), device=dev)
        ws = torch.empty(8, dtype=torch.float64, device=dev) if need_info else None
        _lib.check(lib.vqgnn_gat1_fwd(
            _lib.ptr(plan.fwd_rowptr), _lib.ptr(plan.fwd_col), _lib.ptr(plan.fwd_val), _lib.ptr(plan.fwd_rval),
            _lib.ptr(plan.chunk_rows('fwd')), MP_CHUNK, plan.nnz, B, _lib.ptr(x), x.stride(0), _lib.ptr(bank.codes),
            _lib.ptr(bank.O), nb, M, bank.D, bank.Wp, float(wu), _lib.ptr(a_l), _lib.ptr(a_r), _lib.ptr(cs),
            _lib.ptr(stat), float(slope), _lib.ptr(y), y.stride(0), _lib.ptr(den),
            _lib.ptr(info) if need_info else None, _lib.ptr(ws), st))
        ctx.layer, ctx.plan, ctx.wu, ctx.fire_hook, ctx.slope = layer, plan, float(wu), fire_hook, float(slope)
        ctx.save_for_backward(x, al, ar, a_l, a_r, cs, stat, y, den)
        return y, info

    @staticmethod
    def backward(ctx, dy: Tensor, dinfo: Tensor):
        x, al, ar, a_l, a_r, cs, stat, y, den = ctx.saved_tensors
        layer, plan, wu = ctx.layer, ctx.plan, ctx.wu
        lib, st = _lib.load(), _lib.stream()
        bank = layer.bank
        B, C = x.shape
        nb, M, dev = bank.nb, bank.M, x.device
        if not plan.has_rval:
            raise RuntimeError("the v1 GAT backward needs a training plan (reverse values A_NB_v)")
        dy = dy.contiguous().float()
        dinfo = dinfo.contiguous().float()
        gy = torch.empty(B, nb * 5, device=dev)
        ds, dcs = torch.empty(B, nb, 2, device=dev), torch.empty(nb, M, 2, device=dev)
        datt_l, datt_r = torch.empty(nb, 5, device=dev), torch.empty(nb, 5, device=dev)
        dx = torch.empty(B, C, device=dev) if ctx.needs_input_grad[0] else None
        _lib.check(lib.vqgnn_gat1_bwd(
            _lib.ptr(plan.fwd_rowptr), _lib.ptr(plan.fwd_col), _lib.ptr(plan.fwd_val), _lib.ptr(plan.fwd_rval),
            _lib.ptr(plan.chunk_rows('fwd')), MP_CHUNK, plan.nnz, B, _lib.ptr(x), x.stride(0), _lib.ptr(bank.codes),
            _lib.ptr(bank.O), nb, M, bank.D, bank.Wp, wu, _lib.ptr(al), _lib.ptr(ar), _lib.ptr(a_l), _lib.ptr(a_r),
            _lib.ptr(cs), _lib.ptr(stat), ctx.slope, _lib.ptr(y), y.stride(0), _lib.ptr(den), _lib.ptr(dy),
            dy.stride(0), _lib.ptr(dinfo), _lib.ptr(gy), _lib.ptr(ds), _lib.ptr(dcs), _lib.ptr(dx),
            dx.stride(0) if dx is not None else 0, _lib.ptr(datt_l), _lib.ptr(datt_r), st))
        if ctx.fire_hook:
            # hook(grad) on X_output_B [B, D+1] of every branch (vq_gnn_v1/models.py:199-203): add_flag width
            bank.update(x, gy, plan.batch_idx)
        return dx, datt_l, datt_r, None, None, None, None, None


def gat_conv(layer, x: Tensor, plan: BatchPlan, wu: float, fire_hook: bool):
    """-> (y [B, C] normalised batch rows, info_backward scalar)."""
    if plan.version == 'v1':
        convs = [b.conv for b in layer.gnn_block]
        att_l = torch.stack([c.att_l.reshape(-1) for c in convs])       # [nb, D+1]
        att_r = torch.stack([c.att_r.reshape(-1) for c in convs])
        if layer.num_D != 4:
            raise NotImplementedError("the v1 GAT kernels are written for num_D = 4 (every reference config)")
        return VQGAT1Function.apply(x, att_l, att_r, layer, plan, float(wu), fire_hook,
                                    float(convs[0].negative_slope))
    if plan.version != 'v2':
        raise NotImplementedError("GAT is implemented for the v2 (B+B') formulation; the v1 per-branch "
                                  "(B+M, D+1 columns, add_flag quantiser) GAT is the next row (DESIGN.md)")
    conv = layer.conv
    return VQGATFunction.apply(x, conv.att_l, conv.att_r, layer, plan, float(wu), fire_hook,
                               float(conv.negative_slope))
