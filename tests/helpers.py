"""Shared builders for the parity tests: small seeded graphs, reference-format batches, layers."""
import torch

from vq_gnn_b200 import sampling, synth
from vq_gnn_b200.graph import CSRAdj

LAYER_KW = dict(dropout=0., num_branch=0, cluster='vq', ln_para=True, no_second_fc=True, kmeans_iter=100,
                EMA_flag=True, split=True, kmeans_init=False, dropbranch=0, use_gcn=False, commitment_cost=0.,
                hook=True, weight_ahead=False, transformer_flag=False)


def layer_args(C, C_out, M, D, N, conv_type, skip=False, grad_scale=(1, 1), warm_up_flag=True, momentum=0.1):
    """Positional argument list of LowRankGNNLayer.__init__ (vq_gnn_v2/models.py:67-70)."""
    return [C, C_out, 0., M, D, N, 0, 'vq', True, True, 100, True, True, False, 0, skip, False, 0.,
            list(grad_scale), True, False, warm_up_flag, momentum, conv_type, False]


def make_graph(N, E, conv_type, version, seed=0, power_law=0.0):
    return synth.make_graph(N, E, conv_type, version, seed=seed, power_law=power_law)


def make_batch(g, B, version, seed=0, train=True, recovery=True):
    gen = torch.Generator().manual_seed(seed + 1000)
    node_idx = torch.randperm(g.N, generator=gen)[:B]
    if version == 'v2':
        return sampling.k_hop_batch_v2(g, node_idx, train_flag=train)
    return sampling.collate_batch_v1(g, node_idx, train_flag=train, recovery_flag=recovery)


def batch_to(batch_A, device):
    out = []
    for t in batch_A:
        if t is None:
            out.append(None)
        elif isinstance(t, tuple):
            out.append(tuple(u.to(device) for u in t))
        else:
            out.append(t.to(device))
    return tuple(out)


def to_shim_batch(batch_A, ts):
    """Convert a v2 batch (CSRAdj) into one holding the oracle-shim SparseTensor (for the live reference)."""
    if len(batch_A) != 3:
        return batch_A
    batch_idx, subset, adj = batch_A
    row, col, val = adj.coo()
    return batch_idx, subset, ts.SparseTensor(row=row, col=col, value=val, sparse_sizes=adj.sparse_sizes())


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def att_grad_err(got: dict, want: dict):
    """Error of the GAT attention-vector gradients, relative to their JOINT scale.  "Trick 1" (convs.py:209-211)
    makes the scores invariant to a rescaling of att_l (att_r) when max|a_l| >> 1, so the gradient along
    att_l is a difference of large terms: in fp32 the reference itself moves by ~1e-3 of |d att_l| between
    CPU and GPU (scripts/debug_gat_att_grad.py), while it is stable to ~1e-6 of the joint (att_l, att_r) gradient."""
    keys = [k for k in want if "att_" in k and k in got]
    if not keys:
        return 0.0
    scale = max(float(want[k].abs().max()) for k in keys) + 1e-30
    return max(float((got[k].detach().double().cpu() - want[k].detach().double().cpu()).abs().max()) for k in keys) / scale


def loss_weights(shape):
    """Per-element random loss weights: the gradient must VARY across batch rows, otherwise the
    gradient BatchNorm (eps = 1e-24, vq.py:87) divides rounding noise by 1e-12."""
    return torch.randn(*shape, generator=torch.Generator().manual_seed(77))


def state_mismatches(sd_cuda, sd_oracle, tol):
    """Compare reference-keyed state dicts.  Running means are compared on the scale of the matching
    running std (a gradient that went through a BatchNorm has an exactly-zero batch mean, so its
    running mean is pure rounding noise).  Returns (list of failing keys, number of differing codes)."""
    bad, n_codes = [], 0
    for k, v in sd_oracle.items():
        c = sd_cuda[k].detach().cpu()
        if not v.is_floating_point():
            n_codes += int((c != v).sum())
            continue
        if k.endswith("running_mean"):
            std = sd_oracle[k[:-len("running_mean")] + "running_var"].double().sqrt()
            err = float(((c.double() - v.double()).abs() / (std + v.double().abs() + 1e-30)).max())
        else:
            err = rel_err(c, v)
        if not err < tol:
            bad.append((k, err))
    return bad, n_codes


# ---- golden fixtures (tests/golden/*.npz, written by oracle/gen_golden.py) -----------------------
def pack_batch(batch_A):
    """Flatten a v1 / v2 `batch_A` tuple into named tensors."""
    rec = {}
    if len(batch_A) == 3:
        batch_idx, subset, adj = batch_A
        row, col, val = adj.coo()
        rec.update({"bA.batch_idx": batch_idx, "bA.subset": subset, "bA.row": row, "bA.col": col,
                    "bA.val": val, "bA.dim": torch.tensor(adj.sparse_sizes()[0])})
    else:
        deg_inv, A_BN, A_BB, A_NB_v, batch_idx = batch_A
        rec.update({"bA.deg_inv": deg_inv, "bA.batch_idx": batch_idx})
        for n, t in zip("rcv", A_BN):
            rec["bA.A_BN." + n] = t
        if A_BB is not None:
            for n, t in zip("rcv", A_BB):
                rec["bA.A_BB." + n] = t
        if A_NB_v is not None:
            rec["bA.A_NB_v"] = A_NB_v
    return rec


def unpack_batch(z):
    t = lambda k: torch.from_numpy(z[k])
    if "bA.subset" in z:
        dim = int(z["bA.dim"])
        return t("bA.batch_idx"), t("bA.subset"), CSRAdj.from_coo(t("bA.row"), t("bA.col"), t("bA.val"), (dim, dim))
    A_BN = tuple(t("bA.A_BN." + n) for n in "rcv")
    A_BB = tuple(t("bA.A_BB." + n) for n in "rcv") if "bA.A_BB.r" in z else None
    A_NB_v = t("bA.A_NB_v") if "bA.A_NB_v" in z else None
    return t("bA.deg_inv"), A_BN, A_BB, A_NB_v, t("bA.batch_idx")


def load_golden(name):
    import os

    import numpy as np
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz")
    return np.load(path, allow_pickle=False)


def golden_sd(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}
