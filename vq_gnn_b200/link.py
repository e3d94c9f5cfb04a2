"""Link-prediction head of the collab configuration (BASELINE.json configs[3]): `LinkPredictor`, the in-batch
positive edges and the trivially-sampled negatives of vq_gnn_v2/main_link.py:18-41, 54-71 and
`prepare_batch_input_link` (vq_gnn_v2/utils/misc.py:76-90).  Thin torch glue around the node embeddings that the
VQ layers produce: the MLP is three library GEMMs, the edge gathers are index_selects."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from .graph import BatchPlan

Tensor = torch.Tensor


class LinkPredictor(nn.Module):
    """main_link.py:18-41: sigmoid(MLP(x_i * x_j)), `num_layers` Linear layers, ReLU + dropout in between."""

    def __init__(self, in_channels, hidden_channels, out_channels, num_layers, dropout):
        super().__init__()
        self.lins = nn.ModuleList()
        self.lins.append(nn.Linear(in_channels, hidden_channels))
        for _ in range(num_layers - 2):
            self.lins.append(nn.Linear(hidden_channels, hidden_channels))
        self.lins.append(nn.Linear(hidden_channels, out_channels))
        self.dropout = dropout

    def reset_parameters(self):
        for lin in self.lins:
            lin.reset_parameters()

    def forward(self, x_i, x_j):
        from .models import _linear      # tall inputs: split-K weight gradient (models._TallLinear)
        x = x_i * x_j
        for lin in self.lins[:-1]:
            x = F.relu(_linear(lin, x))
            x = F.dropout(x, p=self.dropout, training=self.training)
        return torch.sigmoid(_linear(self.lins[-1], x))


def positive_edges(batch_A) -> Tuple[Tensor, Tensor]:
    """(src, dst) = the batch graph's edges with both ends among the B batch nodes, in the adjacency's COO order
    (misc.py:87-88: `edge_mask = (edge_index[0] < num_B) & (edge_index[1] < num_B)`).  Accepts the v2 batch tuple
    or a BatchPlan built from it; cached on the plan (one host sync, at batch-preparation time)."""
    if isinstance(batch_A, BatchPlan):
        plan = batch_A
        got = plan.extras.get('pos_edges')
        if got is None:
            B = plan.B
            rowptr, col = plan.fwd_rowptr.long(), plan.fwd_col.long()
            nb_rows = rowptr[:B + 1]
            deg = nb_rows[1:] - nb_rows[:-1]
            row = torch.repeat_interleave(torch.arange(B, device=col.device), deg)
            c = col[:int(nb_rows[-1])]
            keep = c < B
            got = plan.extras['pos_edges'] = (row[keep].contiguous(), c[keep].contiguous())
        return got
    batch_idx, _subset, adj = batch_A
    B = int(batch_idx.shape[0])
    row, col, _ = adj.coo()
    keep = (row < B) & (col < B)
    return row[keep], col[keep]


def link_loss(predictor: LinkPredictor, out: Tensor, batch_A, dst_neg: Optional[Tensor] = None) -> Tensor:
    """main_link.py:57-66: -log(p(src, dst) + 1e-15).mean() - log(1 - p(src, dst_neg) + 1e-15).mean(), one random
    negative destination per positive edge drawn uniformly from the batch."""
    src, dst = positive_edges(batch_A)
    pos_out = predictor(out[src], out[dst])
    pos_loss = -torch.log(pos_out + 1e-15).mean()
    if dst_neg is None:
        dst_neg = torch.randint(0, out.shape[0], src.shape, dtype=torch.long, device=out.device)
    neg_out = predictor(out[src], out[dst_neg])
    neg_loss = -torch.log(1 - neg_out + 1e-15).mean()
    return pos_loss + neg_loss
