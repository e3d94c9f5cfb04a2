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
