"""CPU tests of the host-side logic: C-ABI export list, state_dict parity, plan construction."""
import ctypes
import os
import re

import pytest
import torch

import vq_gnn_b200 as V
from oracle import restate
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "vqgnn.h")).read()
    names = set(re.findall(r"\b(vqgnn_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 10
    lib = ctypes.CDLL(V._lib.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, f"declared in include/vqgnn.h but not exported: {missing}"
    assert lib.vqgnn_abi_version() == 3
    for n in names:   # the binding table must cover the header too
        assert n in V._lib._SIGNATURES, n


def test_tail_table_keeps_the_batch_rows_next_to_the_codeword_rows():
    """models._tail_table / _rows_operand: the materialised codeword table [T, C] is a view into ONE [B + T, C] buffer
    whose first B rows receive the batch-row operand (the reference's x_input = cat(x, codewords),
    vq_gnn_v2/models.py:161-179), so that vqgnn_mp_fwd_rows can address every operand row by a 32-bit offset."""
    from types import SimpleNamespace
    from vq_gnn_b200 import models as Mo
    plan = SimpleNamespace(B=5, T=7, device=torch.device("cpu"))
    t = Mo._tail_table(plan, 8)
    assert t.shape == (7, 8) and t.head_rows.shape == (5, 8)
    assert t.data_ptr() == t.head_rows.data_ptr() + 5 * 8 * 4 and t.is_contiguous()
    x = torch.randn(5, 8)
    xr = Mo._rows_operand(x, t)
    assert xr is t.head_rows and torch.equal(xr, x)
    assert Mo._rows_operand(torch.randn(4, 8), t).shape == (4, 8)          # shape mismatch: the tensor itself
    assert Mo._rows_operand(x, torch.empty(7, 8)) is x                     # a plain table has no spare rows
    # the guard of the kernel's preconditions (width, alignment, one 64 GB window)
    assert Mo._rows_kernel_ok(xr, t, 8) is False and Mo._rows_kernel_ok(torch.empty(5, 16), None, 16) in (True, False)


def test_no_cpu_fallback():
    vq = V.VectorQuantizerEMA(16, 4, grad_normalize_scale=[1, 1])
    with pytest.raises(V._lib.VQGNNLibraryError):
        vq.feature_update(torch.randn(8, 4))
    with pytest.raises(ValueError):
        V.VectorQuantizerEMA(16, 4, grad_normalize_scale=(1, 1))      # vq.py:91-92


@pytest.mark.parametrize("version,conv", [("v2", "GCN"), ("v2", "SAGE"), ("v2", "GAT"), ("v1", "GCN"),
                                          ("v1", "SAGE"), ("v1", "GAT")])
def test_state_dict_keys_and_views(version, conv):
    torch.manual_seed(0)
    layer = V.LowRankGNNLayer(*H.layer_args(8, 6, 16, 4, 50, conv, skip=True), version=version)
    sd = layer.state_dict()
    assert f"gnn_block.1.vq._embedding" in sd and "gnn_block.0.c_indices" in sd
    W = 9 if (version == "v1" and conv == "GAT") else 8
    assert sd["gnn_block.0.vq._embedding"].shape == (16, W)
    assert sd["gnn_block.0.c_indices"].dtype == torch.int16
    # per-branch buffers alias the stacked bank
    layer.bank.O[1, 2, 3] = 5.0
    assert float(layer.gnn_block[1].vq._embedding_output[2, 3]) == 5.0
    # load_state_dict writes through the views
    sd2 = {k: (v.clone() + 1 if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    layer.load_state_dict(sd2)
    assert torch.equal(layer.bank.E[0, :, :W], sd2["gnn_block.0.vq._embedding"])
    assert torch.equal(layer.bank.codes[:, 1], sd2["gnn_block.1.c_indices"])
    # the oracle accepts the same state dict
    o = restate.OracleLayer(8, 6, 16, 4, 50, conv, version, skip=True, warm_up_flag=True).load_state_dict(sd2)
    assert torch.equal(o.vq[1]._ema_w, sd2["gnn_block.1.vq._ema_w"])


@pytest.mark.parametrize("conv", ["GCN", "SAGE", "GAT"])
@pytest.mark.parametrize("train,recovery", [(True, True), (True, False), (False, True)])
def test_plan_v1_equals_mapper(conv, train, recovery):
    """The kernel plan must encode exactly the matrix `mapper` builds (v1/utils/dataloader.py:144-192)."""
    N, B, M = 120, 30, 8
    g = H.make_graph(N, 400, conv, "v1", seed=3)
    batch_A = H.make_batch(g, B, "v1", seed=3, train=train, recovery=recovery)
    c = torch.randint(0, M, (N,), dtype=torch.short, generator=torch.Generator().manual_seed(5))
    ref = restate.mapper_dense(batch_A, c, M, conv)
    plan = V.graph.plan_from_v1(batch_A, conv, N, train, "cpu")
    dense = torch.zeros(B + M, B + M)
    deg = plan.fwd_rowptr[1:] - plan.fwd_rowptr[:-1]
    rows = torch.repeat_interleave(torch.arange(B), deg.long())
    cols, vals = plan.fwd_col.long(), plan.fwd_val
    inb = cols < B
    dense.index_put_((rows[inb], cols[inb]), vals[inb], accumulate=True)
    cm = c.long()[cols[~inb] - B] + B
    dense.index_put_((rows[~inb], cm), vals[~inb], accumulate=True)
    if plan.fwd_rval is not None:
        dense.index_put_((cm, rows[~inb]), plan.fwd_rval[~inb], accumulate=True)
    assert torch.allclose(dense, ref, atol=1e-6), (dense - ref).abs().max()
    # transposed structure == in-batch block transposed
    bt = torch.zeros(B, B)
    bdeg = plan.bwd_rowptr[1:] - plan.bwd_rowptr[:-1]
    bj = torch.repeat_interleave(torch.arange(B), bdeg.long())
    bt.index_put_((plan.bwd_col.long(), bj), plan.bwd_val, accumulate=True)
    assert torch.allclose(bt, ref[:B, :B], atol=1e-6)


@pytest.mark.parametrize("train", [True, False])
def test_plan_v2_roundtrip(train):
    N, B = 150, 40
    g = H.make_graph(N, 500, "GCN", "v2", seed=1)
    batch_idx, subset, adj = H.make_batch(g, B, "v2", seed=1, train=train)
    plan = V.graph.plan_from_v2((batch_idx, subset, adj), "GCN", N, train, "cpu")
    assert plan.B == B and plan.R == (subset.numel() if train else B)
    dense = adj.to_dense()
    rec = torch.zeros_like(dense)
    deg = (plan.fwd_rowptr[1:] - plan.fwd_rowptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(plan.R), deg)
    rec.index_put_((rows, plan.fwd_col.long()), plan.fwd_val, accumulate=True)
    assert torch.equal(rec[:plan.R], dense[:plan.R])
    bt = torch.zeros(plan.R, B)
    bdeg = (plan.bwd_rowptr[1:] - plan.bwd_rowptr[:-1]).long()
    bj = torch.repeat_interleave(torch.arange(B), bdeg)
    bt.index_put_((plan.bwd_col.long(), bj), plan.bwd_val, accumulate=True)
    assert torch.equal(bt, dense[:plan.R, :B])
    assert torch.equal(plan.tail_node.long(), subset[B:])
    # batch nodes lead the subset (vq_gnn_v2/dataloader.py:128)
    assert torch.equal(subset[:B], batch_idx)


@pytest.mark.parametrize("train", [True, False])
def test_k_hop_batch_matches_the_reference_loader(train):
    """sampling.k_hop_batch_v2 (the device-agnostic restatement that csrc/khop.cu is checked against on the GPU) vs the
    UNMODIFIED reference's `OurDataLoader._k_hop_subgraph` (vq_gnn_v2/dataloader.py:98-148) run through
    oracle/ref_loader: same subset (batch nodes first), same edges and values in global node ids."""
    import types
    from oracle import ref_loader
    from vq_gnn_b200 import sampling
    if not ref_loader.available():
        pytest.skip("reference sources not on this box")
    ref = ref_loader.load_reference("v2")
    N, B = 500, 90
    g = H.make_graph(N, 4000, "GCN", "v2", seed=41, power_law=1.4)
    nodes = torch.randperm(N, generator=torch.Generator().manual_seed(3))[:B]
    fake = types.SimpleNamespace(N=N, edge_index=torch.stack([g.row, g.col]), edge_w=g.val, train_flag=train)
    subset_r, ei_r, w_r = ref.dataloader.OurDataLoader._k_hop_subgraph(fake, nodes)
    batch_idx, subset, adj = sampling.k_hop_batch_v2(g, nodes, train_flag=train)
    assert torch.equal(subset[:B], nodes) and torch.equal(subset_r[:B], nodes)
    assert torch.equal(torch.sort(subset)[0], torch.sort(subset_r)[0])

    def triples(sub, row, col, val):
        key = sub[row] * N + sub[col]
        order = torch.argsort(key)
        return key[order], val[order]
    row, col, val = adj.coo()
    k0, v0 = triples(subset, row, col, val)
    k1, v1 = triples(subset_r, ei_r[0], ei_r[1], w_r)
    assert torch.equal(k0, k1) and torch.equal(v0, v1)


def test_link_head_matches_main_link():
    """vq_gnn_b200.link: positive edges = the batch graph's edges with both ends among the batch nodes
    (vq_gnn_v2/utils/misc.py:87-88); loss = -log(p_pos + 1e-15).mean() - log(1 - p_neg + 1e-15).mean() with
    p = sigmoid(MLP(x_i * x_j)) (main_link.py:18-41, 57-66), restated literally here."""
    import torch.nn.functional as F
    from vq_gnn_b200 import link, sampling
    N, B = 400, 70
    g = H.make_graph(N, 3000, "GCN", "v2", seed=12)
    nodes = torch.randperm(N, generator=torch.Generator().manual_seed(2))[:B]
    bA = sampling.k_hop_batch_v2(g, nodes, True)
    src, dst = link.positive_edges(bA)
    row, col, _ = bA[2].coo()
    mask = (row < B) & (col < B)
    assert torch.equal(src, row[mask]) and torch.equal(dst, col[mask]) and src.numel() > 0
    torch.manual_seed(0)
    pred = link.LinkPredictor(16, 12, 1, 3, 0.0)
    assert [tuple(l.weight.shape) for l in pred.lins] == [(12, 16), (12, 12), (1, 12)]
    out = torch.randn(B, 16)
    dst_neg = torch.randint(0, B, src.shape, generator=torch.Generator().manual_seed(4))

    def p(xi, xj):
        h = xi * xj
        for lin in pred.lins[:-1]:
            h = F.relu(lin(h))
        return torch.sigmoid(pred.lins[-1](h))
    want = -torch.log(p(out[src], out[dst]) + 1e-15).mean() - torch.log(1 - p(out[src], out[dst_neg]) + 1e-15).mean()
    got = link.link_loss(pred, out, bA, dst_neg=dst_neg)
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)


def test_rw_and_edge_samplers():
    """rw / edge samplers (vq_gnn_v2/dataloader.py:71-75): walks follow edges, isolated nodes stay put, outputs unique."""
    from vq_gnn_b200 import sampling
    N = 300
    g = H.make_graph(N, 900, "GCN", "v1", seed=8)          # v1 graphs store no diagonal
    gen = torch.Generator().manual_seed(1)
    seeds = torch.randperm(N, generator=gen)[:40]
    w = sampling.random_walk(g, seeds, 4, generator=gen)
    assert w.shape == (40, 5) and torch.equal(w[:, 0], seeds)
    dense = torch.zeros(N, N, dtype=torch.bool)
    dense[g.row, g.col] = True
    deg = g.rowptr[1:] - g.rowptr[:-1]
    for t in range(4):
        a, b = w[:, t], w[:, t + 1]
        ok = dense[a, b] | ((deg[a] == 0) & (a == b))
        assert bool(ok.all())
    nodes = sampling.rw_sampler(g, seeds, 4, generator=gen)
    assert torch.equal(nodes, torch.unique(nodes)) and nodes.numel() <= 40 * 5
    e = sampling.edge_sampler(g, seeds, generator=gen)
    assert torch.equal(e, torch.unique(e)) and e.numel() <= 80 and bool(torch.isin(seeds, e).all())
