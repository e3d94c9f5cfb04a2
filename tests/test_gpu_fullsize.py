"""Parity at BASELINE.json's full config shapes on the GPU.  The CPU oracle cannot run these sizes (dense
(B+B')^2 adjacency), so each case checks the CUDA layer op (through the C-ABI) against
  (1) a plain PyTorch fp32 restatement of the same op on the GPU (tests/torch_ref.py: torch.sparse / index_add +
      autograd) -- outputs, info_backward, d x (and d att for GAT) within 1e-4 relative, and
  (2) size-independent properties: count histograms sum to B exactly and codes stay in range, re-assignment
      without an update is idempotent, the two forward kernels (generic L2-gather vs shared-memory codebook)
      agree, and forward/backward are adjoint: <y(x1) - y(x2), w> == <x1 - x2, dx(w)>.
Graphs are the seeded synthetic shapes of vq_gnn_b200/synth.py (no dataset is available offline)."""
import pytest
import torch

import vq_gnn_b200 as V
from tests import helpers as H
from tests import torch_ref as R
from vq_gnn_b200 import sampling, synth
from vq_gnn_b200.models import VQConvFunction, _trigger
from vq_gnn_b200.gat import VQGATFunction

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _layer(C, C_out, M, N, conv, version, dev, seed=0, skip=False):
    torch.manual_seed(seed)
    layer = V.LowRankGNNLayer(*H.layer_args(C, C_out, M, 4, N, conv, skip=skip), version=version).to(dev).train()
    return layer


def _warm(layer, x, plan, steps=2):
    """feature warm start + `steps` real updates so gradient codewords and codes are non-trivial."""
    dev = x.device
    out = layer(x, plan, 1.0, False)
    layer.set_inited(True)
    for s in range(steps):
        xx = x.clone().requires_grad_(True)
        out = layer(xx, plan, 1.0, False)
        w = torch.randn(out[0].shape, device=dev, generator=torch.Generator(device=dev).manual_seed(50 + s))
        ((out[0] * w).sum() + out[5]).backward()
    layer.check_status()


def _vq_properties(layer, x, plan):
    bank = layer.bank
    B = x.shape[0]
    counts = bank.last_stats[:, :, bank.Wp]
    assert torch.equal(counts.sum(1), torch.full((bank.nb,), float(B), device=x.device))     # histogram, bit-exact
    assert int(bank.codes.min()) >= 0 and int(bank.codes.max()) < bank.M
    layer.eval()
    i1 = bank.run(x.detach(), None, plan.batch_idx, False, write_codes=False).clone()
    i2 = bank.run(x.detach(), None, plan.batch_idx, False, write_codes=False)
    layer.train()
    assert torch.equal(i1, i2)                                                                # idempotent
    assert int(i1.min()) >= 0 and int(i1.max()) < bank.M


def _check_conv(layer, x, plan, wu, ref_fn):
    dev = x.device
    gen = torch.Generator(device=dev).manual_seed(7)
    w = torch.randn(x.shape, device=dev, generator=gen)
    xc = x.clone().requires_grad_(True)
    y, info = VQConvFunction.apply(xc, None, layer, plan, wu, False)
    ((y * w).sum() + info).backward()
    xr = x.clone().requires_grad_(True)
    yr, infor = ref_fn(xr, plan, layer.bank, wu)
    ((yr * w).sum() + infor).backward()
    assert H.rel_err(y, yr) < TOL, ("y", H.rel_err(y, yr))
    assert abs(float(info) - float(infor)) <= TOL * max(1e-3, abs(float(infor))), ("info", float(info), float(infor))
    assert H.rel_err(xc.grad, xr.grad) < TOL, ("dx", H.rel_err(xc.grad, xr.grad))
    # adjointness of the fwd / bwd kernel pair (dinfo = 0 isolates A^T)
    x2 = x + torch.randn(x.shape, device=dev, generator=gen)
    x2c = x2.clone().requires_grad_(True)
    y2, _ = VQConvFunction.apply(x2c, None, layer, plan, wu, False)
    xa = x.clone().requires_grad_(True)
    ya, _ = VQConvFunction.apply(xa, None, layer, plan, wu, False)
    (ya * w).sum().backward()
    lhs = float(((y2 - ya).double() * w.double()).sum())
    rhs = float(((x2 - x).double() * xa.grad.double()).sum())
    # both sides are sums of ~B*C signed terms: compare on the scale of the terms (sum |a_i b_i|), not of the
    # (cancelling) totals -- y2 - ya subtracts fp32 numbers that share the large out-of-batch contribution
    scale = float(((y2 - ya).double().abs() * w.double().abs()).sum())
    assert abs(lhs - rhs) <= 1e-4 * scale, (lhs, rhs, scale)
    return y, info


def test_c1_arxiv_shape_v2_gcn():
    dev = torch.device("cuda:0")
    s = synth.CONFIG_SHAPES["c1_arxiv"]
    g = synth.make_graph(s["N"], s["E"], "GCN", "v2", seed=0, num_blocks=80, device=dev)
    parts = torch.randperm(80, generator=torch.Generator().manual_seed(1))[:40]
    nodes = sampling.cluster_batch(s["N"], 80, parts).to(dev)
    bA = sampling.k_hop_batch_v2(g, nodes, True)
    layer = _layer(128, 128, 256, s["N"], "GCN", "v2", dev)
    plan = V.build_plan(bA, "GCN", s["N"], True, dev)
    x = torch.randn(nodes.numel(), 128, device=dev, generator=torch.Generator(device=dev).manual_seed(2))
    _warm(layer, x, plan)
    _vq_properties(layer, x, plan)
    _check_conv(layer, x, plan, 0.9, R.gcn_v2)


def test_c2_reddit_shape_v1_sage_both_kernels():
    dev = torch.device("cuda:0")
    s = synth.CONFIG_SHAPES["c2_reddit"]
    g = synth.make_graph(s["N"], s["E"], "SAGE", "v1", seed=0, power_law=2.2, device=dev)
    gen = torch.Generator(device=dev).manual_seed(3)
    seeds = torch.randperm(s["N"], generator=gen, device=dev)[:6000]
    nodes = sampling.cont_sampler(g, seeds, 3, 6000, generator=gen)[2]
    bA = sampling.collate_batch_v1(g, nodes, True, True)
    layer = _layer(128, 128, 1024, s["N"], "SAGE", "v1", dev)
    plan = V.build_plan(bA, "SAGE", s["N"], True, dev)
    assert plan.nnz > 2_000_000
    x = torch.randn(nodes.numel(), 128, device=dev, generator=torch.Generator(device=dev).manual_seed(2))
    _warm(layer, x, plan)
    _vq_properties(layer, x, plan)
    layer.use_tail_kernel = False
    y0, i0 = _check_conv(layer, x, plan, 1.0, R.sage_gcn_v1)           # generic L2-gather kernel
    layer.use_tail_kernel = 'force'
    y1, i1 = _check_conv(layer, x, plan, 1.0, R.sage_gcn_v1)           # shared-memory codebook kernel
    assert H.rel_err(y1, y0) < 1e-5 and abs(float(i1) - float(i0)) <= 1e-4 * max(1e-3, abs(float(i0)))


def test_c3_ppi_shape_v2_gat():
    dev = torch.device("cuda:0")
    s = synth.CONFIG_SHAPES["c3_ppi"]
    g = synth.make_graph(s["N"], s["E"], "GAT", "v2", seed=0, num_blocks=20, device=dev)
    nodes = torch.randperm(s["N"], generator=torch.Generator().manual_seed(4))[:10000].to(dev)
    bA = sampling.k_hop_batch_v2(g, nodes, True)
    C = 256
    layer = _layer(C, C, 4096, s["N"], "GAT", "v2", dev, skip=True)
    plan = V.build_plan(bA, "GAT", s["N"], True, dev)
    x = torch.randn(nodes.numel(), C, device=dev, generator=torch.Generator(device=dev).manual_seed(2))
    _warm(layer, x, plan)
    _vq_properties(layer, x, plan)
    gen = torch.Generator(device=dev).manual_seed(7)
    w = torch.randn(x.shape, device=dev, generator=gen)
    conv = layer.conv
    xc = x.clone().requires_grad_(True)
    conv.att_l.grad = conv.att_r.grad = None
    y, info = VQGATFunction.apply(xc, conv.att_l, conv.att_r, layer, plan, 0.9, False, 0.2)
    ((y * w).sum() + info).backward()
    xr = x.clone().requires_grad_(True)
    al = conv.att_l.detach().clone().requires_grad_(True)
    ar = conv.att_r.detach().clone().requires_grad_(True)
    yr, infor = R.gat_v2(xr, plan, layer.bank, 0.9, al.view(-1), ar.view(-1))
    ((yr * w).sum() + infor).backward()
    assert H.rel_err(y, yr) < TOL, ("y", H.rel_err(y, yr))
    assert abs(float(info) - float(infor)) <= TOL * max(1e-3, abs(float(infor)))
    assert H.rel_err(xc.grad, xr.grad) < TOL, ("dx", H.rel_err(xc.grad, xr.grad))
    err = H.att_grad_err({"att_l": conv.att_l.grad, "att_r": conv.att_r.grad}, {"att_l": al.grad, "att_r": ar.grad})
    assert err < TOL, ("att", err)


def test_c5_products_shape_v2_gcn():
    dev = torch.device("cuda:0")
    s = synth.CONFIG_SHAPES["c5_products"]
    g = synth.make_graph(s["N"], s["E"], "GCN", "v2", seed=0, num_blocks=64, device=dev)
    # rank 0 of 8: node sampler inside its own contiguous partition
    lo, hi = V.dist.partition_range(s["N"], 0, 8)
    nodes = (lo + torch.randperm(hi - lo, generator=torch.Generator().manual_seed(5))[:20000]).to(dev)
    bA = sampling.k_hop_batch_v2(g, nodes, True)
    layer = _layer(128, 128, 4096, s["N"], "GCN", "v2", dev)
    plan = V.build_plan(bA, "GCN", s["N"], True, dev)
    x = torch.randn(nodes.numel(), 128, device=dev, generator=torch.Generator(device=dev).manual_seed(2))
    _warm(layer, x, plan, steps=1)
    _vq_properties(layer, x, plan)
    _check_conv(layer, x, plan, 1.0, R.gcn_v2)


def test_c4_collab_shape_v2_gcn_link():
    """configs[3]: the layer op at the collab shape (integer edge weights kept as values, power-law degrees, cont
    sampler batch of 50000) + one link-prediction train step through LinkPredictor (positive edges = in-batch edges)."""
    from vq_gnn_b200 import link
    dev = torch.device("cuda:0")
    s = synth.CONFIG_SHAPES["c4_collab"]
    N = s["N"]
    rowptr, row, col = synth.random_edges(N, s["E"], seed=0, power_law=s["power_law"], device=dev)
    lo_, hi_ = torch.minimum(row, col), torch.maximum(row, col)
    w = ((lo_ * 2654435761 + hi_ * 40503) % 3 + 1).float()
    g = synth.normalized_graph(N, rowptr, row, col, "GCN", "v2", edge_weight=w)
    gen = torch.Generator(device=dev).manual_seed(6)
    seeds = torch.randperm(N, generator=gen, device=dev)[:50000]
    nodes = sampling.cont_sampler(g, seeds, 15, 50000, generator=gen)[4]
    layer = _layer(128, 128, 1024, N, "GCN", "v2", dev, skip=True)
    plan = V.graph.plan_from_graph_v2(g, nodes, "GCN", True).warm()
    x = torch.randn(nodes.numel(), 128, device=dev, generator=torch.Generator(device=dev).manual_seed(2))
    _warm(layer, x, plan)
    _vq_properties(layer, x, plan)
    _check_conv(layer, x, plan, 1.0, R.gcn_v2)
    # one link-prediction step: finite loss, gradients reach the layer and the predictor
    torch.manual_seed(1)
    pred = link.LinkPredictor(128, 128, 1, 3, 0.0).to(dev)
    src, dst = link.positive_edges(plan)
    assert src.numel() > 0 and int(src.max()) < plan.B and int(dst.max()) < plan.B
    xx = x.clone().requires_grad_(True)
    out = layer(xx, plan, 1.0, False)
    loss = link.link_loss(pred, out[0], plan) + out[5]
    loss.backward()
    assert torch.isfinite(loss) and float(xx.grad.abs().sum()) > 0
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in pred.parameters())
