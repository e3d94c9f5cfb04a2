"""DevicePrefetcher (vq_gnn_b200/loader.py): batches uploaded / prepared on the side stream must train exactly
like batches moved synchronously, in both modes (enqueue-from-caller and worker thread)."""
import pytest
import torch
import torch.nn.functional as F

import vq_gnn_b200 as V
from tests import helpers as H
from vq_gnn_b200.loader import DevicePrefetcher

pytestmark = pytest.mark.gpu


def _pin(t):
    if t is None:
        return None
    if isinstance(t, tuple):
        return tuple(_pin(u) for u in t)
    return t.contiguous().pin_memory()


@pytest.mark.parametrize("version,conv,threaded", [("v1", "SAGE", False), ("v1", "SAGE", True), ("v1", "GCN", True),
                                                   ("v2", "GCN", False), ("v2", "GCN", True), ("v2", "SAGE", True),
                                                   ("v2", "GAT", False), ("v2", "GAT", True)])
def test_prefetched_batches_train_like_direct_ones(version, conv, threaded):
    """GCN / SAGE: the whole path (plan builders, message passing, VQ update) is free of order-dependent float
    accumulation, so a run fed through the prefetcher (side stream, optionally a worker thread) must reproduce the
    direct run BIT FOR BIT -- any race or atomic-order dependence shows up as a differing bit.  GAT (fp32 REDs in
    its kernels): same losses / state within 1e-4 on the scale of each tensor."""
    dev = torch.device("cuda:0")
    N, B, M, C = 600, 120, 16, 8
    g = H.make_graph(N, 6000, conv, version, seed=3)
    host = []
    for s in range(3):
        bA = H.make_batch(g, B, version, seed=s)
        x = torch.randn(B, C, generator=torch.Generator().manual_seed(s))
        y = torch.randint(0, 5, (B,), generator=torch.Generator().manual_seed(10 + s))
        host.append((_pin(x), _pin(bA) if version == "v1" else bA, _pin(y)))

    def build():
        torch.manual_seed(0)
        m = V.LowRankGNN(C, 8, 5, 2, 0., M, 4, N, no_second_fc=True, skip=False, commitment_cost=0.,
                         grad_scale=[1, 1], act='relu', bn_flag=True, warm_up_flag=True, conv_type=conv,
                         version=version).to(dev).train()
        return m, torch.optim.SGD(m.parameters(), lr=1e-3)

    def step(m, opt, x, bA, y, i):
        if i == 1:
            m.set_inited(True)
        opt.zero_grad()
        out, _, info = m((x, bA), 1)
        loss = F.cross_entropy(out, y) + info
        loss.backward()
        opt.step()
        return float(loss)

    m0, o0 = build()
    ref = [step(m0, o0, x.to(dev), H.batch_to(bA, dev), y.to(dev), i) for i, (x, bA, y) in enumerate(host * 2)]
    m1, o1 = build()
    pf = DevicePrefetcher(host, dev, prepare=lambda b: (b[0], m1.prepare(b[1]), b[2]), count=6, threaded=threaded)
    got = []
    for i in range(6):
        x, plan, y = pf.next()
        got.append(step(m1, o1, x, plan, y, i))
    pf.drain()
    sd0, sd1 = m0.state_dict(), m1.state_dict()
    if conv != "GAT":
        assert ref == got, (ref, got)
        for k in sd0:
            assert torch.equal(sd0[k], sd1[k]), k
    else:
        assert ref == pytest.approx(got, rel=1e-4, abs=1e-5)
        for k in sd0:
            a, b = sd0[k].float(), sd1[k].float()
            assert float((a - b).abs().max()) <= 1e-4 * max(float(a.abs().max()), 1e-6), k


@pytest.mark.parametrize("conv,train,power_law", [("GCN", True, 0.0), ("GAT", True, 1.3), ("SAGE", False, 1.3),
                                                  ("GCN", False, 0.0)])
def test_device_khop_v2_matches_torch_restatement(conv, train, power_law):
    """csrc/khop.cu (vqgnn_khop_mark / count / fill + transposed CSR) vs sampling.k_hop_batch_v2 -> plan_from_v2
    (itself checked against the reference's `_k_hop_subgraph` in tests/test_host_logic.py): same subset, same
    relabelled adjacency (entry for entry: both keep the graph's stored order inside a row up to the relabelling),
    same backward structure."""
    from vq_gnn_b200 import graph as G, sampling
    dev = torch.device("cuda:0")
    N, B = 4000, 600
    g = H.make_graph(N, 50_000, conv, "v2", seed=23, power_law=power_law).to(dev)
    nodes = torch.randperm(N, generator=torch.Generator().manual_seed(5))[:B].to(dev)
    p_t = G.plan_from_v2(sampling.k_hop_batch_v2(g, nodes, train_flag=train), conv, N, train, dev)
    p_d = G.plan_from_graph_v2(g, nodes, conv, train)
    assert (p_d.B, p_d.R, p_d.T) == (p_t.B, p_t.R, p_t.T)
    assert torch.equal(p_d.batch_idx, p_t.batch_idx) and torch.equal(p_d.tail_node, p_t.tail_node)
    assert torch.equal(p_d.fwd_rowptr, p_t.fwd_rowptr)

    def dense(plan):
        deg = (plan.fwd_rowptr[1:] - plan.fwd_rowptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(plan.R, device=dev), deg)
        a = torch.zeros(plan.R, plan.B + plan.T, device=dev, dtype=torch.float64)
        return a.index_put_((rows, plan.fwd_col.long()), plan.fwd_val.double(), accumulate=True)

    def dense_bwd(plan):
        bdeg = (plan.bwd_rowptr[1:] - plan.bwd_rowptr[:-1]).long()
        bj = torch.repeat_interleave(torch.arange(plan.B, device=dev), bdeg)
        t = torch.zeros(plan.R, plan.B, device=dev, dtype=torch.float64)
        return t.index_put_((plan.bwd_col.long(), bj), plan.bwd_val.double(), accumulate=True)
    assert torch.equal(dense(p_d), dense(p_t))
    if train:
        assert torch.equal(p_d.bwd_rowptr, p_t.bwd_rowptr) and torch.equal(dense_bwd(p_d), dense_bwd(p_t))
        # every column of the transposed CSR is sorted by source row (order-independent of the cursor scatter)
        bj = torch.repeat_interleave(torch.arange(p_d.B, device=dev), (p_d.bwd_rowptr[1:] - p_d.bwd_rowptr[:-1]).long())
        key = bj.long() * p_d.R + p_d.bwd_col.long()
        assert bool((key[1:] > key[:-1]).all())


@pytest.mark.parametrize("conv,train,recovery", [("SAGE", True, True), ("GCN", True, True), ("GAT", True, False),
                                                 ("SAGE", False, True)])
def test_device_collate_v1_matches_torch_restatement(conv, train, recovery):
    """csrc/khop.cu (vqgnn_collate_v1_count / fill) produces the reference's v1 batch tuple bit for bit like
    sampling.collate_batch_v1 (vq_gnn_v1/utils/dataloader.py:64-86)."""
    from vq_gnn_b200 import graph as G, sampling
    dev = torch.device("cuda:0")
    N, B = 3000, 400
    g = H.make_graph(N, 60_000, conv, "v1", seed=29, power_law=1.3).to(dev)
    nodes = torch.randperm(N, generator=torch.Generator().manual_seed(6))[:B].to(dev)
    want = sampling.collate_batch_v1(g, nodes, train_flag=train, recovery_flag=recovery)
    got = G.batch_from_graph_v1(g, nodes, train_flag=train, recovery_flag=recovery)
    assert torch.equal(got[0], want[0]) and torch.equal(got[4], want[4])
    for a, b in zip(got[1], want[1]):
        assert torch.equal(a, b)
    assert (got[2] is None) == (want[2] is None) and (got[3] is None) == (want[3] is None)
    if want[2] is not None:
        for a, b in zip(got[2], want[2]):
            assert torch.equal(a, b)
    if want[3] is not None:
        assert torch.equal(got[3], want[3])


def test_prepare_from_graph_trains_like_the_host_batch():
    """model.prepare_from_graph (device-built batch) vs model.prepare of the host-format batch: identical training
    (bit for bit: GCN path) over a few steps."""
    from vq_gnn_b200 import sampling
    dev = torch.device("cuda:0")
    N, B, M, C = 800, 150, 16, 8
    for version, conv in (("v2", "GCN"), ("v1", "SAGE")):
        g = H.make_graph(N, 8000, conv, version, seed=3).to(dev)
        X = torch.randn(N, C, generator=torch.Generator().manual_seed(1)).to(dev)
        Y = torch.randint(0, 5, (N,), generator=torch.Generator().manual_seed(2)).to(dev)
        ids = [torch.randperm(N, generator=torch.Generator().manual_seed(10 + s))[:B].to(dev) for s in range(3)]

        def run(from_graph):
            torch.manual_seed(0)
            m = V.LowRankGNN(C, 8, 5, 2, 0., M, 4, N, no_second_fc=True, skip=False, commitment_cost=0.,
                             grad_scale=[1, 1], act='relu', bn_flag=True, warm_up_flag=True, conv_type=conv,
                             version=version).to(dev).train()
            opt = torch.optim.SGD(m.parameters(), lr=1e-3)
            losses = []
            for i, nd in enumerate(ids * 2):
                if i == 1:
                    m.set_inited(True)
                if from_graph:
                    plan = m.prepare_from_graph(g, nd)
                else:
                    bA = (sampling.k_hop_batch_v2(g, nd, True) if version == "v2"
                          else sampling.collate_batch_v1(g, nd, True, True))
                    plan = m.prepare(bA)
                opt.zero_grad()
                out, _, info = m((X[nd], plan), 1)
                loss = F.cross_entropy(out, Y[nd]) + info
                loss.backward()
                opt.step()
                losses.append(float(loss))
            return losses, m.state_dict()
        l0, s0 = run(False)
        l1, s1 = run(True)
        if version == "v1":          # same entry order in both builders: bit-identical
            assert l0 == l1
            for k in s0:
                assert torch.equal(s0[k], s1[k]), k
        else:                        # v2: the torch builder sorts a row by relabelled column, the device one keeps the
            assert l0 == pytest.approx(l1, rel=1e-5)        # graph's stored order -- same sums in a different order


@pytest.mark.parametrize("version,conv", [("v2", "GCN"), ("v1", "SAGE"), ("v2", "GAT")])
def test_streaming_warm_start_matches_per_batch_init(version, conv):
    """LowRankGNN.warm_start (the reference's init(): L(L+1)/2 streaming layer passes over the whole graph with the
    test loader's batch-rows-only batches, main_node.py:17-37) vs the same loop written with host-built batches."""
    from vq_gnn_b200 import sampling
    dev = torch.device("cuda:0")
    N, M, C, bs = 900, 16, 8, 250
    g = H.make_graph(N, 7000, conv, version, seed=5).to(dev)
    X = torch.randn(N, C, generator=torch.Generator().manual_seed(1)).to(dev)

    def build():
        torch.manual_seed(0)
        return V.LowRankGNN(C, 8, 5, 3, 0., M, 4, N, no_second_fc=True, skip=False, commitment_cost=0.,
                            grad_scale=[1, 1], act='leaky_gelu', bn_flag=True, warm_up_flag=True, conv_type=conv,
                            version=version).to(dev).train()
    m0, m1 = build(), build()
    n_b = m0.warm_start(g, X, bs)
    assert n_b == (N + bs - 1) // bs and all(l.inited for l in m0.convs)
    with torch.no_grad():
        for layer_idx in range(1, 4):
            for lo in range(0, N, bs):
                ids = torch.arange(lo, min(lo + bs, N), device=dev)
                bA = (sampling.k_hop_batch_v2(g, ids, train_flag=False) if version == "v2"
                      else sampling.collate_batch_v1(g, ids, train_flag=False, recovery_flag=False))
                plan = V.build_plan(bA, conv, N, False, dev)
                plan.training = True          # the reference's init(): eval-structured batches, model in train mode
                m1.init((X[ids], plan), layer_idx)
    m1.set_inited(True)
    sd0, sd1 = m0.state_dict(), m1.state_dict()
    for k in sd0:
        a, b = sd0[k].float(), sd1[k].float()
        assert float((a - b).abs().max()) <= 1e-5 * max(float(a.abs().max()), 1e-6), k
    m0.check_status()
