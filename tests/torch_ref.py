"""Plain PyTorch fp32 restatements of the layer-level ops that run ON THE GPU at full config sizes (the CPU oracle
uses dense (B+B')^2 / (B+M)^2 adjacencies and is limited to small cases).  Same formulas as oracle/restate.py
(vq_gnn_v2/models.py:161-198, vq_gnn_v2/convs.py:65-101,165-266, vq_gnn_v1/models.py:170-223), expressed with
torch.sparse / index_add so autograd provides the reference gradients.  Test infrastructure only."""
import torch
import torch.nn.functional as F


def xin_v2(x, plan, bank):
    """[x ; codeword features] and the gradient codewords of the tail nodes."""
    D = bank.D
    codes = bank.codes[plan.tail_node.long()].long()                    # [B', nb]
    ar = torch.arange(bank.nb, device=x.device)
    cw = bank.O[ar.view(1, -1), codes]                                  # [B', nb, Wp]
    xf = cw[:, :, :D].reshape(codes.shape[0], -1)
    gf = cw[:, :, D:2 * D].reshape(codes.shape[0], -1)
    return torch.cat([x, xf], 0), gf


def _coo(plan):
    deg = (plan.fwd_rowptr[1:] - plan.fwd_rowptr[:-1]).long()
    row = torch.repeat_interleave(torch.arange(plan.R, device=deg.device), deg)
    return row, plan.fwd_col.long(), plan.fwd_val


def gcn_v2(x, plan, bank, wu):
    """-> (Y[:B], info)   Y = A [x ; xf], info = wu * sum(Y[B:] * gf)."""
    B = plan.B
    xin, gf = xin_v2(x, plan, bank)
    row, col, val = _coo(plan)
    n = max(plan.R, xin.shape[0])
    A = torch.sparse_coo_tensor(torch.stack([row, col]), val, (plan.R, xin.shape[0])).coalesce()
    y = torch.sparse.mm(A, xin)
    info = (y[B:] * gf[:plan.R - B]).sum() * wu if plan.R > B else y.sum() * 0
    return y[:B], info


def gat_v2(x, plan, bank, wu, att_l, att_r, slope=0.2):
    B = plan.B
    xin, gf = xin_v2(x, plan, bank)
    xin = torch.cat([xin, torch.ones(xin.shape[0], 1, device=x.device)], 1)
    a_l = (xin * att_l.view(1, -1)).sum(-1)
    a_r = (xin * att_r.view(1, -1)).sum(-1)
    scale = torch.sqrt(a_l.max() ** 2 + 1) * torch.sqrt(a_r.max() ** 2 + 1)
    a_l, a_r = a_l / scale, a_r / scale
    row, col, val = _coo(plan)
    w = val * F.leaky_relu(a_l[col] + a_r[row], slope).exp()
    y = torch.zeros(plan.R, xin.shape[1], device=x.device).index_add_(0, row, w.unsqueeze(1) * xin[col])
    yB = y[:B, :-1] / (y[:B, -1:] + 1e-16)
    info = (y[B:, :-1] * gf[:plan.R - B]).sum() * wu if plan.R > B else y.sum() * 0
    return yB, info


def sage_gcn_v1(x, plan, bank, wu):
    """v1 (B+M) formulation evaluated per edge: y = A_in x + wu * sum_e val O[code, :D];
    info = wu * sum_r <x[r], sum_e rval O[code, D:2D]>."""
    B, D, nb = plan.B, bank.D, bank.nb
    row, col, val = _coo(plan)
    tail = col >= B
    A_in = torch.sparse_coo_tensor(torch.stack([row[~tail], col[~tail]]), val[~tail], (B, B)).coalesce()
    y = torch.sparse.mm(A_in, x)
    tr, tn, tv, trv = row[tail], col[tail] - B, val[tail], plan.fwd_rval[tail]
    ar = torch.arange(nb, device=x.device)
    yq = torch.zeros(B, nb * D, device=x.device)
    gq = torch.zeros(B, nb * D, device=x.device)
    step = 1 << 20                                                       # bound the per-edge materialisation
    for s in range(0, tr.numel(), step):
        sl = slice(s, s + step)
        cw = bank.O[ar.view(1, -1), bank.codes[tn[sl]].long()]           # [e, nb, Wp]
        yq.index_add_(0, tr[sl], (tv[sl].view(-1, 1, 1) * cw[:, :, :D]).reshape(-1, nb * D))
        gq.index_add_(0, tr[sl], (trv[sl].view(-1, 1, 1) * cw[:, :, D:2 * D]).reshape(-1, nb * D))
    y = y + wu * yq
    info = wu * (x * gq).sum()
    return y, info
