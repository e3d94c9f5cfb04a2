"""CPU restatement (plain torch, fp32, dense adjacency) of the VQ-GNN hot path.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this module; the product path
(`vq_gnn_b200/`) never does and fails loudly without its CUDA library.

PARITY PINNING: the reference ships no tests or golden vectors.  This
restatement is pinned against outputs of the reference ITSELF, executed unmodified in the
builder container through `oracle/ref_loader.py` (pure-torch shims for the absent
torch_geometric / torch_sparse / torch_scatter leaf symbols): `tests/test_oracle_vs_reference.py`
runs both live, and `oracle/gen_golden.py` freezes the reference's outputs into
`tests/golden/*.npz`, which `tests/test_oracle_golden.py` re-checks everywhere (GPU box included).

Every function cites the reference lines it follows (paths relative to /root/reference).
State is kept per branch under the reference's own `state_dict()` key names so the same
state dict drives the reference, this oracle, and the CUDA product.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# VectorQuantizerEMA  (vq_gnn_v2/vq.py:60-279; v1 identical minus get())
# --------------------------------------------------------------------------------------
class OracleVQ:
    """One branch's quantiser state + the two update rules.

    Buffers mirror vq.py:73-88: `_embedding`, `_embedding_output`, `_ema_cluster_size`,
    `_ema_w`, and the two BatchNorm1d(affine=False) running statistics.
    """

    def __init__(self, num_embeddings: int, embedding_dim: int, decay: float = 0.99,
                 epsilon: float = 1e-24, grad_normalize_scale: Sequence[float] = (1, 1),
                 warm_up_flag: bool = False, momentum: float = 0.1, add_flag: bool = False,
                 init_random: bool = True):
        self.M, self.D = num_embeddings, embedding_dim
        self.add = 1 if add_flag else 0
        self.W = 2 * embedding_dim + self.add
        self.decay, self.eps = decay, epsilon
        self.scale = list(grad_normalize_scale)
        self.warm_up_flag, self.momentum = warm_up_flag, momentum
        D, W = self.D, self.W
        # vq.py:73-80 (same RNG consumption order as the reference constructor)
        self._embedding = torch.randn(self.M, W) if init_random else torch.zeros(self.M, W)
        self._embedding_output = torch.zeros(self.M, W)
        self._ema_cluster_size = torch.zeros(self.M)
        self._ema_w = torch.zeros(self.M, W)
        if warm_up_flag and init_random:
            self._ema_w.normal_()
        # vq.py:86-88
        self.feat_mean, self.feat_var = torch.zeros(D), torch.ones(D)
        self.grad_mean, self.grad_var = torch.zeros(D + self.add), torch.ones(D + self.add)
        # vq.py:93-98
        self._embedding[:, D:2 * D] *= self.scale[0]
        self._ema_w[:, D:2 * D] *= self.scale[0]
        if self.add:
            self._embedding[:, 2 * D] *= self.scale[1]
            self._ema_w[:, 2 * D] *= self.scale[1]
        self.bn_inited = False
        self.training = True
        # test harness: assignments of the implementation under test for the NEXT call (one-shot).  Rows where they
        # differ from this oracle's own argmin must be near-ties (relative gap < tie_gap, BASELINE.json north_star);
        # the forced codes are then used for the EMA so that the state comparison sees identical assignments.
        self.force_idx: Optional[Tensor] = None
        self.tie_gap = 1e-5
        self.forced_mismatches, self.forced_total = 0, 0

    def _choose(self, dist: Tensor, z: Tensor, e: Tensor) -> Tensor:
        idx = torch.argmin(dist, dim=1)
        self.last_dist, self.last_z, self.last_own_idx = dist, z, idx
        f, self.force_idx = self.force_idx, None
        if f is None:
            return idx
        f = f.view(-1).long().cpu()
        assert f.numel() == idx.numel()
        bad = (f != idx).nonzero().flatten()
        if bad.numel():
            # gap relative to the magnitude of the terms of ||z||^2 + ||e||^2 - 2 z.e (what fp32 rounding scales with)
            z2, e2 = (z[bad] ** 2).sum(1), (e ** 2).sum(1)
            d0, d1 = dist[bad, idx[bad]], dist[bad, f[bad]]
            gap = (d1 - d0).abs() / (z2 + torch.maximum(e2[idx[bad]], e2[f[bad]]) + 1e-30)
            if not float(gap.max()) < self.tie_gap:
                raise AssertionError(f"assignment mismatch that is not a near-tie: rel gap {float(gap.max()):.3e} "
                                     f"(allowed < {self.tie_gap:g}) at {int(bad.numel())} of {int(idx.numel())} rows")
        self.forced_mismatches += int(bad.numel())
        self.forced_total += int(idx.numel())
        return f

    # -- state-dict plumbing (reference key names) ---------------------------------------
    def load(self, sd: Dict[str, Tensor], prefix: str = "") -> "OracleVQ":
        g = lambda k: sd[prefix + k].detach().clone().float().cpu()
        self._embedding, self._embedding_output = g("_embedding"), g("_embedding_output")
        self._ema_cluster_size, self._ema_w = g("_ema_cluster_size"), g("_ema_w")
        self.feat_mean, self.feat_var = g("batch_norm_feat.running_mean"), g("batch_norm_feat.running_var")
        self.grad_mean, self.grad_var = g("batch_norm_grad.running_mean"), g("batch_norm_grad.running_var")
        return self

    def dump(self, prefix: str = "") -> Dict[str, Tensor]:
        return {prefix + "_embedding": self._embedding.clone(),
                prefix + "_embedding_output": self._embedding_output.clone(),
                prefix + "_ema_cluster_size": self._ema_cluster_size.clone(),
                prefix + "_ema_w": self._ema_w.clone(),
                prefix + "batch_norm_feat.running_mean": self.feat_mean.clone(),
                prefix + "batch_norm_feat.running_var": self.feat_var.clone(),
                prefix + "batch_norm_grad.running_mean": self.grad_mean.clone(),
                prefix + "batch_norm_grad.running_var": self.grad_var.clone()}

    # -- accessors (vq.py:120-127) ---------------------------------------------------------
    def get(self):
        return self._embedding_output

    def get_codebook(self):
        return self._embedding_output[:, :self.D]

    def get_grad(self):
        return self._embedding_output[:, self.D:]

    # -- helpers ----------------------------------------------------------------------------
    def _bn_feat(self, x):  # BatchNorm1d(D, affine=False): eps 1e-5, momentum 0.1 (vq.py:86)
        return F.batch_norm(x, self.feat_mean, self.feat_var, None, None, self.training, 0.1, 1e-5)

    def _bn_grad(self, g):  # BatchNorm1d(D+add, eps=epsilon, momentum=momentum) (vq.py:87-88)
        return F.batch_norm(g, self.grad_mean, self.grad_var, None, None, self.training,
                            self.momentum, self.eps)

    def _ema_size(self, counts):
        # vq.py:177-189 / :242-254
        self._ema_cluster_size = self._ema_cluster_size * self.decay + (1 - self.decay) * counts
        if self.warm_up_flag:
            n = torch.sum(self._ema_cluster_size)
            self._ema_cluster_size = (self._ema_cluster_size + 1e-5) / (n + self.M * 1e-5) * n
        if torch.count_nonzero(self._ema_cluster_size) != self.M:
            raise ValueError('Bad Init!')

    @staticmethod
    def distances(z: Tensor, e: Tensor) -> Tensor:
        # vq.py:166-168 / :230-232 : ||z||^2 + ||e||^2 - 2 z e^T, that association order, fp32
        return (torch.sum(z ** 2, dim=1, keepdim=True) + torch.sum(e ** 2, dim=1)
                - 2 * torch.matmul(z, e.t()))

    # -- vq.py:160-202 ------------------------------------------------------------------------
    def feature_update(self, X_B: Tensor) -> Tensor:
        D = self.D
        xn = self._bn_feat(X_B)
        dist = self.distances(xn, self._embedding[:, :D])
        idx = self._choose(dist, xn, self._embedding[:, :D])
        if self.training:
            counts = torch.bincount(idx, minlength=self.M).float()  # == sum(one-hot, 0)
            self._ema_size(counts)
            dw = torch.zeros(self.M, D).index_add_(0, idx, xn)       # == one-hot^T @ xn
            self._ema_w[:, :D] = self._ema_w[:, :D] * self.decay + (1 - self.decay) * dw
            self._embedding[:, :D] = self._ema_w[:, :D] / self._ema_cluster_size.unsqueeze(1)
            std = torch.sqrt(self.feat_var + 1e-5).unsqueeze(0)
            self._embedding_output[:, :D] = self._embedding[:, :D] * std + self.feat_mean.unsqueeze(0)
        return idx.unsqueeze(1)

    # -- vq.py:204-279 ------------------------------------------------------------------------
    def update(self, X_B: Tensor, grad: Tensor) -> Tuple[Tensor, None]:
        D = self.D
        if not self.bn_inited:  # vq.py:216-221 (unbiased var)
            self.feat_mean, self.feat_var = X_B.mean(0).clone(), X_B.var(0).clone()
            self.grad_mean, self.grad_var = grad.mean(0).clone(), grad.var(0).clone()
            self.bn_inited = True
        z = torch.cat([self._bn_feat(X_B), self._bn_grad(grad)], dim=1)
        z[:, D:2 * D] *= self.scale[0]
        if self.add:
            z[:, 2 * D] *= self.scale[1]
        dist = self.distances(z, self._embedding)
        idx = self._choose(dist, z, self._embedding)
        if self.training:
            counts = torch.bincount(idx, minlength=self.M).float()
            self._ema_size(counts)
            dw = torch.zeros(self.M, self.W).index_add_(0, idx, z)
            self._ema_w = self._ema_w * self.decay + (1 - self.decay) * dw
            self._embedding = self._ema_w / self._ema_cluster_size.unsqueeze(1)
            out = self._embedding.clone()
            out[:, D:2 * D] /= self.scale[0] + self.eps
            if self.add:
                out[:, 2 * D] /= self.scale[1] + self.eps
            var = torch.cat([self.feat_var + 1e-5, self.grad_var + self.eps])
            mean = torch.cat([self.feat_mean, self.grad_mean])
            self._embedding_output = out * torch.sqrt(var).unsqueeze(0) + mean.unsqueeze(0)
            if self.scale[0] == 0:
                self._embedding_output[:, D:] *= 0
        return idx.unsqueeze(1), None


# --------------------------------------------------------------------------------------
# message passing primitives  (convs.py)
# --------------------------------------------------------------------------------------
def gcn_propagate(adj: Tensor, x: Tensor) -> Tensor:
    """OurGCNConv.forward == adj @ x, nothing else (v2/convs.py:65-101)."""
    return adj @ x


def gat_propagate(adj: Tensor, x: Tensor, att_l: Tensor, att_r: Tensor,
                  negative_slope: float = 0.2) -> Tensor:
    """OurGATConv.forward/message with vq_softmax == un-normalised exp
    (v2/convs.py:165-266, v2/utils/vq_softmax.py:41-57).  `adj[i, j]`: i = target row, j = source.
    att_* are flat [C'] vectors (reference shape [1, 1, C'])."""
    a_l = (x * att_l.view(1, -1)).sum(-1)            # convs.py:189
    a_r = (x * att_r.view(1, -1)).sum(-1)            # convs.py:190
    scale = torch.sqrt(torch.max(a_l) ** 2 + 1) * torch.sqrt(torch.max(a_r) ** 2 + 1)  # :209
    a_l, a_r = a_l / scale, a_r / scale
    e = F.leaky_relu(a_l.view(1, -1) + a_r.view(-1, 1), negative_slope).exp()   # :259-261
    w = torch.where(adj != 0, e * adj, torch.zeros_like(adj))                    # Trick 2 (:264)
    return w @ x


# --------------------------------------------------------------------------------------
# v1 mapper  (vq_gnn_v1/utils/dataloader.py:144-192), dense restatement
# --------------------------------------------------------------------------------------
def mapper_dense(batch_A, c: Tensor, num_M: int, gnn_type: str) -> Tensor:
    deg_inv, A_BN, A_BB, A_NB_v, batch_idx = batch_A
    B = batch_idx.shape[0]
    dim = B + num_M
    c = c.to(torch.long)
    adj = torch.zeros(dim, dim)
    r, col, v = A_BN
    cm = c[col] + B
    adj.index_put_((r, cm), v, accumulate=True)                       # :149-151
    if A_NB_v is not None:
        adj.index_put_((cm, r), A_NB_v, accumulate=True)              # :153-154
    if A_BB is not None:
        br, bc, bv = A_BB
        adj.index_put_((br, bc), bv, accumulate=True)                 # :156-158
        adj.index_put_((br, c[batch_idx[bc]] + B), -bv, accumulate=True)   # :160-163
        if A_NB_v is not None:
            adj.index_put_((c[batch_idx[br]] + B, bc), -bv, accumulate=True)  # :165-168
    adj = torch.where(adj > 0, adj, torch.zeros_like(adj))            # :177-180
    if gnn_type != 'SAGE':                                            # :182-185
        i = torch.arange(B)
        adj[i, i] += deg_inv
    if gnn_type == 'GCN':                                             # :189-190 (A + A^T, sums)
        adj = adj + adj.t()
    return adj


def mapper_sparse(batch_A, c: Tensor, num_M: int, gnn_type: str) -> Tensor:
    """Same matrix as `mapper_dense`, built the way the reference builds it: concatenate COO triples,
    `coalesce` (sort + sum duplicates), drop values <= 0, add self loops, symmetrise
    (vq_gnn_v1/utils/dataloader.py:144-192).  Returns a coalesced torch sparse COO tensor; this is the
    form the CPU baseline times (the dense form is O((B+M)^2) memory)."""
    deg_inv, A_BN, A_BB, A_NB_v, batch_idx = batch_A
    B = batch_idx.shape[0]
    dim = B + num_M
    c = c.to(torch.long)
    r, col, v = A_BN
    cm = c[col] + B
    rows, cols, vals = [r], [cm], [v]
    if A_NB_v is not None:
        rows.append(cm), cols.append(r), vals.append(A_NB_v)
    if A_BB is not None:
        br, bc, bv = A_BB
        rows.append(br), cols.append(bc), vals.append(bv)
        rows.append(br), cols.append(c[batch_idx[bc]] + B), vals.append(-bv)
        if A_NB_v is not None:
            rows.append(c[batch_idx[br]] + B), cols.append(bc), vals.append(-bv)
    idx = torch.stack([torch.cat(rows), torch.cat(cols)])
    adj = torch.sparse_coo_tensor(idx, torch.cat(vals), (dim, dim)).coalesce()
    keep = adj.values() > 0
    idx, vals = adj.indices()[:, keep], adj.values()[keep]
    if gnn_type != 'SAGE':
        i = torch.arange(B)
        idx = torch.cat([idx, torch.stack([i, i])], 1)
        vals = torch.cat([vals, deg_inv])
    if gnn_type == 'GCN':
        idx = torch.cat([idx, idx.flip(0)], 1)
        vals = torch.cat([vals, vals])
    return torch.sparse_coo_tensor(idx, vals, (dim, dim)).coalesce()


# --------------------------------------------------------------------------------------
# layer forward, v2 "B+B'" formulation  (vq_gnn_v2/models.py:144-231)
# --------------------------------------------------------------------------------------
class OracleLayer:
    """LowRankGNNLayer restated.  `version` selects v1 (per-branch B+M graphs through `mapper`,
    vq_gnn_v1/models.py:143-233,307-367) or v2 (one B+B' graph, vq_gnn_v2/models.py:144-231).

    hook_mode: 'fire'  -> the VQ hook runs on d loss / d conv-output (v1 behaviour, and the v2
                          authors' evident intent)
               'literal_v2' -> reproduce v2's dangling-slice bug: the hook never runs.
    """

    def __init__(self, in_channels: int, out_channels: int, num_M: int, num_D: int, num_N: int,
                 conv_type: str = 'GCN', version: str = 'v2', skip: bool = False,
                 grad_scale: Sequence[float] = (1, 1), warm_up_flag: bool = False,
                 momentum: float = 0.1, hook_mode: str = 'fire', sparse: bool = False,
                 branches: Optional[Sequence[int]] = None):
        assert in_channels % num_D == 0, 'Cannot fully split'
        self.C, self.C_out, self.M, self.D, self.N = in_channels, out_channels, num_M, num_D, num_N
        self.nb = in_channels // num_D
        self.conv_type, self.version, self.skip, self.hook_mode = conv_type, version, skip, hook_mode
        self.sparse = sparse          # v1 GCN/SAGE only: torch.sparse mapper + spmm (CPU-baseline form)
        self.branches = branches      # bench sampling: run only these branches (outputs of the others are 0)
        add_flag = (version == 'v1' and conv_type == 'GAT')          # v1/models.py:53 ; v2/models.py:30
        self.vq: List[OracleVQ] = []
        self.c_indices: List[Tensor] = []
        for _ in range(self.nb):
            self.c_indices.append(torch.randint(0, num_M, (num_N,), dtype=torch.short))
            self.vq.append(OracleVQ(num_M, num_D, grad_normalize_scale=grad_scale,
                                    warm_up_flag=warm_up_flag, momentum=momentum, add_flag=add_flag))
        self.inited = False
        self.training = True
        self.params: Dict[str, Tensor] = {}
        # test harness: [B, nb] codes chosen by the implementation under test in the same step (see OracleVQ.force_idx)
        self.forced_codes: Optional[Tensor] = None

    def _force(self, i: int):
        if self.forced_codes is not None:
            self.vq[i].force_idx = self.forced_codes[:, i]

    def forced_mismatch_rate(self) -> float:
        tot = sum(q.forced_total for q in self.vq)
        return sum(q.forced_mismatches for q in self.vq) / max(tot, 1)

    # ---- load parameters / buffers from a reference-keyed state dict ------------------------
    def load_state_dict(self, sd: Dict[str, Tensor]) -> "OracleLayer":
        for i in range(self.nb):
            self.vq[i].load(sd, f"gnn_block.{i}.vq.")
            self.c_indices[i] = sd[f"gnn_block.{i}.c_indices"].detach().clone().cpu()
        self.params = {}
        for k, v in sd.items():
            if not k.startswith("gnn_block.") or ".conv." in k:
                if v.is_floating_point():
                    self.params[k] = v.detach().clone().float().cpu().requires_grad_(True)
        return self

    def state_dict(self) -> Dict[str, Tensor]:
        out = {}
        for i in range(self.nb):
            out.update(self.vq[i].dump(f"gnn_block.{i}.vq."))
            out[f"gnn_block.{i}.c_indices"] = self.c_indices[i].clone()
        return out

    def train(self, mode: bool = True):
        self.training = mode
        for q in self.vq:
            q.training = mode
        return self

    def set_inited(self, flag: bool = True):
        self.inited = flag

    def _att(self, i: Optional[int]):
        if self.version == 'v1':
            return (self.params[f"gnn_block.{i}.conv.att_l"].view(-1),
                    self.params[f"gnn_block.{i}.conv.att_r"].view(-1))
        return self.params["conv.att_l"].view(-1), self.params["conv.att_r"].view(-1)

    def _linear(self, name, x):
        return F.linear(x, self.params[name + ".weight"], self.params[name + ".bias"])

    # ---- hooks -----------------------------------------------------------------------------
    def _fire(self, i: int, X_B: Tensor, batch_idx: Tensor, grad: Tensor):
        self._force(i)
        idx, _ = self.vq[i].update(X_B, grad)                         # models.py:39-46
        self.c_indices[i][batch_idx] = idx.squeeze(1).to(torch.short)

    # ---- forward ---------------------------------------------------------------------------
    def forward(self, x: Tensor, batch_A, warm_up_rate: float = 1.0, unlabeled: bool = False):
        if self.version == 'v1':
            y, info = self._conv_v1(x, batch_A, warm_up_rate, unlabeled)
        else:
            y, info = self._conv_v2(x, batch_A, warm_up_rate, unlabeled)
        self.last_conv_out = y
        out = self._linear("gnn_transform", y)                         # v2/models.py:202
        if self.conv_type == 'SAGE':
            out = out + self._linear("fc_sage", x)                     # :203-204
        if self.skip:
            out = out + self._linear("linear_skip", x)                 # :228-229
        info_backwards = info if self.training else 0                  # :199-200
        return out, info_backwards

    __call__ = forward

    def _conv_v2(self, x, batch_A, wu, unlabeled):
        batch_idx, subset, adj = batch_A
        adj = adj.to_dense() if not isinstance(adj, Tensor) else adj
        B, D = x.shape[0], self.D
        first_order_idx = subset[B:]
        xf, gf = [], []
        for i in range(self.nb):
            xs = x[:, D * i:D * (i + 1)]
            if not self.inited or unlabeled:                           # :165-166, :61-63
                self._force(i)
                idx = self.vq[i].feature_update(xs.detach())
                self.c_indices[i][batch_idx] = idx.squeeze(1).to(torch.short)
            codes = self.c_indices[i][first_order_idx].to(torch.long)  # :168
            cw = self.vq[i].get()[codes]                               # :169
            xf.append(cw[:, :D].clone()), gf.append(cw[:, D:].clone())
        xf, gf = torch.cat(xf, 1), torch.cat(gf, 1)
        xin = torch.cat([x, xf], 0)                                    # :173
        if self.conv_type == 'GAT':
            xin = torch.cat([xin, torch.ones(xin.shape[0], 1)], 1)     # :176-177
            y = gat_propagate(adj, xin, *self._att(None))
        else:
            y = gcn_propagate(adj, xin)
        y_B = y[:B]
        if self.inited and self.training and not unlabeled and self.hook_mode == 'fire' \
                and torch.is_grad_enabled():
            X_det = x.detach()
            if not y_B.requires_grad:      # x_output_B.requires_grad_() on a no-grad tensor: a new leaf
                y_B = y_B.detach().requires_grad_()

            def hook(grad, X_det=X_det, batch_idx=batch_idx):
                for i in range(self.nb):
                    self._fire(i, X_det[:, D * i:D * (i + 1)], batch_idx, grad[:, D * i:D * (i + 1)])
                return grad
            y_B.register_hook(hook)
        y_rest = y[B:]
        if self.conv_type == 'GAT':                                    # :187-189
            y_B = y_B[:, :-1] / (y_B[:, -1].unsqueeze(1) + 1e-16)
            y_rest = y_rest[:, :-1]
        info = torch.sum(y_rest * gf * wu)                             # :198
        return y_B, info

    def _conv_v1(self, x, batch_A, wu, unlabeled):
        D, B = self.D, x.shape[0]
        batch_idx = batch_A[-1]
        outs, info_total = [], 0
        for i in range(self.nb):
            X_B = x[:, D * i:D * (i + 1)]
            if self.branches is not None and i not in self.branches:
                outs.append(torch.zeros(B, D))
                continue
            if self.training and (not self.inited or unlabeled):       # v1/models.py:149-165
                self._force(i)
                idx = self.vq[i].feature_update(X_B.detach())
                self.c_indices[i][batch_idx] = idx.squeeze(1).to(torch.short)
            if self.sparse and self.conv_type != 'GAT':
                adj = mapper_sparse(batch_A, self.c_indices[i], self.M, self.conv_type)
            else:
                adj = mapper_dense(batch_A, self.c_indices[i], self.M, self.conv_type)   # :170
            X_bar = self.vq[i].get_codebook().clone()                  # :173
            X_in = torch.cat([X_B, X_bar * wu], 0)                     # :181
            if self.conv_type == 'GAT':
                X_in = torch.cat([X_in, torch.ones(X_in.shape[0], 1)], 1)    # :188-189
                X_out = gat_propagate(adj, X_in, *self._att(i))
            elif adj.is_sparse:
                X_out = torch.sparse.mm(adj, X_in)
            else:
                X_out = gcn_propagate(adj, X_in)
            X_out_B, X_out_M = X_out[:B], X_out[B:]                    # :197
            if self.inited and self.training and not unlabeled and torch.is_grad_enabled():  # :199-203
                if not X_out_B.requires_grad:   # X_output_B.requires_grad_() (:202) makes it a leaf
                    X_out_B = X_out_B.detach().requires_grad_()

                def hook(grad, i=i, X_det=X_B.detach(), batch_idx=batch_idx):
                    self._fire(i, X_det, batch_idx, grad)
                    return grad
                X_out_B.register_hook(hook)
            if self.conv_type == 'GAT':                                # :209-210
                X_out_B = X_out_B[:, :D] / (X_out_B[:, D].unsqueeze(1) + 1e-16)
            info = torch.sum(X_out_M * self.vq[i].get_grad().clone() * wu)   # :223
            if self.training:
                info_total = info_total + info
            outs.append(X_out_B)
        return torch.cat(outs, 1), info_total                          # :337


# --------------------------------------------------------------------------------------
# adjacency normalisation (vq_gnn_v2/utils/misc.py:14-34 ; vq_gnn_v1/main_node.py:323-349)
# --------------------------------------------------------------------------------------
def norm_adj_dense_v2(A: Tensor, conv_type: str) -> Tensor:
    A = A.clone()
    n = A.shape[0]
    if conv_type in ('GCN', 'GAT'):
        A[torch.arange(n), torch.arange(n)] = 1.0
    deg = A.sum(1)
    if conv_type == 'GCN':
        dis = deg.pow(-0.5)
        dis[dis == float('inf')] = 0
        return dis.view(-1, 1) * A * dis.view(1, -1)
    di = deg.pow(-1)
    di[di == float('inf')] = 0
    return di.view(-1, 1) * A


def act_leaky_gelu(x):  # v2/models.py:296
    return 0.1 * x + 0.9 * F.gelu(x)
