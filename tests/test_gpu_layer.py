"""GPU parity: LowRankGNNLayer / LowRankGNN fwd + bwd + VQ hook (CUDA through the C-ABI) vs the CPU
oracle (oracle/restate.py, itself pinned to the unmodified reference) on identical state and inputs."""
import pytest
import torch

import vq_gnn_b200 as V
from oracle import restate
from tests import helpers as H

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


def _run_cuda(layer, batch_A, x, steps, dev, wu=1.0):
    layer.train()
    outs = []
    bA = H.batch_to(batch_A, dev)
    for s in range(steps):
        if s == 1:
            layer.set_inited(True)
        xx = x.clone().to(dev).requires_grad_(True)
        for p in layer.parameters():
            p.grad = None
        out = layer(xx, bA, wu, False)
        w = H.loss_weights(out[0].shape).to(dev)
        loss = (out[0] * w).sum() + out[5]
        loss.backward()
        outs.append((out[0].detach().cpu(), torch.as_tensor(out[5]).detach().cpu(), xx.grad.cpu(),
                     {k: p.grad.cpu() for k, p in layer.named_parameters() if p.grad is not None},
                     layer.bank.last_idx.cpu().long()))
    return outs


def _run_oracle(o, batch_A, x, steps, wu=1.0, cuda_outs=None):
    """cuda_outs: the CUDA run of the same steps -- its codes are handed to the oracle (OracleLayer.forced_codes), which
    asserts that every row where they differ from its own argmin is a near-tie (relative gap < 1e-5) and then
    continues from identical assignments."""
    o.train()
    outs = []
    for s in range(steps):
        if s == 1:
            o.set_inited(True)
        o.forced_codes = None if cuda_outs is None else cuda_outs[s][4]
        xx = x.clone().requires_grad_(True)
        for p in o.params.values():
            p.grad = None
        out, info = o(xx, batch_A, wu, False)
        loss = (out * H.loss_weights(out.shape)).sum() + info
        loss.backward()
        outs.append((out.detach(), torch.as_tensor(info).detach(), xx.grad.clone(),
                     {k: p.grad.clone() for k, p in o.params.items() if p.grad is not None}))
    return outs


def _compare(c_outs, o_outs, layer, o):
    for s, ((co, ci, cg, cp, _), (oo, oi, og, op)) in enumerate(zip(c_outs, o_outs)):
        assert H.rel_err(co, oo) < REL_TOL, (s, "out", H.rel_err(co, oo))
        assert abs(float(ci) - float(oi)) <= REL_TOL * max(1e-3, abs(float(oi))), (s, "info", float(ci), float(oi))
        assert H.rel_err(cg, og) < REL_TOL, (s, "dx", H.rel_err(cg, og))
        for k in op:
            if k in cp and "att_" not in k:
                assert H.rel_err(cp[k], op[k]) < REL_TOL, (s, k)
        assert H.att_grad_err(cp, op) < REL_TOL, (s, "att grads", H.att_grad_err(cp, op))
    bad, n_code_mismatch = H.state_mismatches(layer.state_dict(), o.state_dict(), REL_TOL)
    assert not bad, bad
    assert n_code_mismatch == 0, n_code_mismatch
    assert o.forced_mismatch_rate() < 1e-3, o.forced_mismatch_rate()     # near-ties only (checked by the oracle)


CASES = [("v2", "GCN", 8, 4, False), ("v2", "SAGE", 8, 4, False), ("v1", "GCN", 8, 4, False),
         ("v1", "SAGE", 8, 4, False), ("v2", "GCN", 128, 4, True), ("v1", "SAGE", 128, 4, False),
         ("v2", "SAGE", 12, 2, False), ("v1", "GCN", 24, 8, True), ("v2", "GCN", 52, 4, False),
         ("v2", "GAT", 8, 4, True), ("v2", "GAT", 128, 4, True), ("v2", "GAT", 52, 4, False),
         ("v2", "GAT", 12, 2, True), ("v2", "GAT", 260, 4, True),
         ("v1", "GAT", 8, 4, True), ("v1", "GAT", 128, 4, False), ("v1", "GAT", 52, 4, True)]


@pytest.mark.parametrize("version,conv,C,D,skip", CASES)
def test_layer_matches_oracle(version, conv, C, D, skip):
    dev = torch.device("cuda:0")
    N, B, M, C_out = 400, 120, 16, 10
    g = H.make_graph(N, 2000, conv, version, seed=7)
    batch_A = H.make_batch(g, B, version, seed=7)
    torch.manual_seed(11)
    layer = V.LowRankGNNLayer(*H.layer_args(C, C_out, M, D, N, conv, skip=skip), version=version)
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    o = restate.OracleLayer(C, C_out, M, D, N, conv, version, skip=skip, warm_up_flag=True).load_state_dict(sd)
    layer = layer.to(dev)
    x = torch.randn(B, C, generator=torch.Generator().manual_seed(3))
    c_outs = _run_cuda(layer, batch_A, x, 4, dev, wu=0.7)
    o_outs = _run_oracle(o, batch_A, x, 4, wu=0.7, cuda_outs=c_outs)
    _compare(c_outs, o_outs, layer, o)
    layer.check_status()
    # the hook really fired: gradient codewords became non-zero, info_backward is non-zero
    assert float(layer.bank.O[:, :, D:2 * D].abs().sum()) > 0
    assert abs(float(c_outs[-1][1])) > 0


def test_literal_v2_hooks_never_fire():
    dev = torch.device("cuda:0")
    N, B, M, C = 300, 80, 16, 8
    g = H.make_graph(N, 1200, "GCN", "v2", seed=2)
    batch_A = H.make_batch(g, B, "v2", seed=2)
    torch.manual_seed(1)
    layer = V.LowRankGNNLayer(*H.layer_args(C, 6, M, 4, N, "GCN"), version="v2", literal_v2_hooks=True)
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    o = restate.OracleLayer(C, 6, M, 4, N, "GCN", "v2", warm_up_flag=True, hook_mode="literal_v2").load_state_dict(sd)
    layer = layer.to(dev)
    layer.materialize_tail = 'force'      # v2: dense rows of gathered codewords (no-op for v1 / GAT)
    x = torch.randn(B, C, generator=torch.Generator().manual_seed(3))
    c_outs = _run_cuda(layer, batch_A, x, 3, dev)
    _compare(c_outs, _run_oracle(o, batch_A, x, 3, cuda_outs=c_outs), layer, o)
    assert float(layer.bank.O[:, :, 4:].abs().sum()) == 0


def test_eval_mode_and_unlabeled():
    dev = torch.device("cuda:0")
    N, B, M, C = 300, 80, 16, 8
    for version, conv in (("v2", "GCN"), ("v1", "SAGE"), ("v2", "GAT")):   # v2 GAT eval: scores span all B + B' nodes
        g = H.make_graph(N, 1200, conv, version, seed=4)
        torch.manual_seed(2)
        layer = V.LowRankGNNLayer(*H.layer_args(C, 6, M, 4, N, conv), version=version)
        sd = {k: v.clone() for k, v in layer.state_dict().items()}
        o = restate.OracleLayer(C, 6, M, 4, N, conv, version, warm_up_flag=True).load_state_dict(sd)
        layer = layer.to(dev)
        x = torch.randn(B, C, generator=torch.Generator().manual_seed(3))
        tr = H.make_batch(g, B, version, seed=4, train=True)
        c_tr = _run_cuda(layer, tr, x, 2, dev)
        _compare(c_tr, _run_oracle(o, tr, x, 2, cuda_outs=c_tr), layer, o)
        ev = H.make_batch(g, B, version, seed=5, train=False)
        layer.eval(), o.train(False)
        with torch.no_grad():
            out_c = layer(x.to(dev), H.batch_to(ev, dev), 1, False)
            out_o, info_o = o(x, ev, 1.0, False)
        assert H.rel_err(out_c[0], out_o) < REL_TOL
        assert out_c[5] == 0 and info_o == 0


def test_full_model_train_step_matches_oracle_stack():
    """3-layer LowRankGNN (bn + leaky_gelu) vs the same stack built from OracleLayers."""
    import torch.nn.functional as F
    dev = torch.device("cuda:0")
    N, B, M, D = 500, 150, 16, 4
    for version, conv in (("v2", "GCN"), ("v1", "SAGE")):
        g = H.make_graph(N, 2500, conv, version, seed=9)
        batch_A = H.make_batch(g, B, version, seed=9)
        torch.manual_seed(5)
        model = V.LowRankGNN(12, 16, 7, 3, 0., M, D, N, no_second_fc=True, skip=False, commitment_cost=0.,
                             grad_scale=[1, 1], act='leaky_gelu', bn_flag=True, warm_up_flag=True,
                             conv_type=conv, version=version)
        dims = [(12, 16), (16, 16), (16, 7)]
        oracles = []
        for li, (ci, co) in enumerate(dims):
            lsd = {k[len(f"convs.{li}."):]: v.clone() for k, v in model.state_dict().items()
                   if k.startswith(f"convs.{li}.")}
            oracles.append(restate.OracleLayer(ci, co, M, D, N, conv, version, warm_up_flag=True).load_state_dict(lsd))
        model = model.to(dev).train()
        x = torch.randn(B, 12, generator=torch.Generator().manual_seed(3))
        y = torch.randint(0, 7, (B,), generator=torch.Generator().manual_seed(4))
        bA = H.batch_to(batch_A, dev)
        for step in range(4):
            if step == 1:
                model.set_inited(True)
                for o in oracles:
                    o.set_inited(True)
            # --- CUDA
            model.zero_grad()
            out, _, info = model((x.to(dev), bA), 1)
            loss = F.cross_entropy(out, y.to(dev)) + info
            loss.backward()
            # --- oracle stack (vq_gnn_v2/models.py:308-348), continuing from the CUDA path's (near-tie-checked) codes
            h = x.clone()
            info_o = 0
            for li, o in enumerate(oracles):
                o.forced_codes = model.convs[li].bank.last_idx.cpu().long()
                for p in o.params.values():
                    p.grad = None
                h, inf = o(h, batch_A, 1.0, False)
                info_o = info_o + inf
                if li < 2:
                    h = F.batch_norm(h, None, None, training=True)
                    h = restate.act_leaky_gelu(h)
            loss_o = F.cross_entropy(h, y) + info_o
            loss_o.backward()
            assert H.rel_err(out, h) < 5 * REL_TOL, (version, step, H.rel_err(out, h))
            assert abs(float(loss) - float(loss_o)) < 5 * REL_TOL * max(1.0, abs(float(loss_o)))
            for li, o in enumerate(oracles):
                gw = model.convs[li].gnn_transform.weight.grad.cpu()
                assert H.rel_err(gw, o.params["gnn_transform.weight"].grad) < 5 * REL_TOL, (version, step, li)
        for li, o in enumerate(oracles):
            bad, n_codes = H.state_mismatches(model.convs[li].state_dict(), o.state_dict(), 5 * REL_TOL)
            assert not bad and n_codes == 0, (version, li, bad, n_codes)


@pytest.mark.parametrize("version,conv", [("v1", "SAGE"), ("v1", "GCN"), ("v2", "GCN"), ("v2", "SAGE"),
                                          ("v2", "GAT"), ("v1", "GAT")])
def test_hub_rows_cut_by_chunk_boundaries(version, conv):
    """Power-law graph whose hub rows hold thousands of entries (>> the 256-entry warp chunk of the
    message-passing kernels) next to empty rows: exercises the RED-accumulated partial rows."""
    dev = torch.device("cuda:0")
    N, B, M, C, D = 3000, 400, 32, 8, 4
    g = H.make_graph(N, 150_000, conv, version, seed=21, power_law=1.2)
    deg = g.rowptr[1:] - g.rowptr[:-1]
    assert int(deg.max()) > 1024
    hubs = torch.argsort(deg, descending=True)[:B // 2]
    gen = torch.Generator().manual_seed(5)
    rest = torch.randperm(N, generator=gen)
    rest = rest[~torch.isin(rest, hubs)][:B - hubs.numel()]
    node_idx = torch.cat([hubs, rest])[torch.randperm(B, generator=gen)]
    from vq_gnn_b200 import sampling
    batch_A = (sampling.k_hop_batch_v2(g, node_idx, True) if version == "v2"
               else sampling.collate_batch_v1(g, node_idx, True, True))
    torch.manual_seed(13)
    layer = V.LowRankGNNLayer(*H.layer_args(C, 6, M, D, N, conv), version=version)
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    o = restate.OracleLayer(C, 6, M, D, N, conv, version, warm_up_flag=True).load_state_dict(sd)
    layer = layer.to(dev)
    x = torch.randn(B, C, generator=torch.Generator().manual_seed(3))
    c_outs = _run_cuda(layer, batch_A, x, 3, dev)
    _compare(c_outs, _run_oracle(o, batch_A, x, 3, cuda_outs=c_outs), layer, o)


@pytest.mark.parametrize("conv,C,M,B,E", [("SAGE", 8, 16, 120, 2000), ("GCN", 8, 16, 120, 2000),
                                          ("SAGE", 28, 1024, 200, 6000), ("SAGE", 132, 64, 150, 40000),
                                          ("GCN", 24, 512, 64, 30000)])
def test_v1_shared_memory_tail_kernel(conv, C, M, B, E):
    """The shared-memory codebook kernel (csrc/mp_tail.cu) forced on small graphs: short and empty rows,
    branch groups of 8 (M <= 768) and 6 (M = 1024) with a ragged last group, rows cut by chunk boundaries."""
    dev = torch.device("cuda:0")
    N, D, C_out = 600, 4, 10
    g = H.make_graph(N, E, conv, "v1", seed=17, power_law=1.5 if E > 10000 else 0.0)
    batch_A = H.make_batch(g, B, "v1", seed=17)
    torch.manual_seed(19)
    layer = V.LowRankGNNLayer(*H.layer_args(C, C_out, M, D, N, conv), version="v1")
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    o = restate.OracleLayer(C, C_out, M, D, N, conv, "v1", warm_up_flag=True).load_state_dict(sd)
    layer = layer.to(dev)
    layer.use_tail_kernel = 'force'
    x = torch.randn(B, C, generator=torch.Generator().manual_seed(3))
    l0 = V._lib.launch_count()
    c_outs = _run_cuda(layer, batch_A, x, 3, dev, wu=0.8)
    assert layer.bank.codes_g is not None and V._lib.launch_count() > l0
    _compare(c_outs, _run_oracle(o, batch_A, x, 3, wu=0.8, cuda_outs=c_outs), layer, o)
    # the group-major mirror tracks the code table
    G = layer.bank.G
    cg = layer.bank.grouped_codes()
    for k in range(layer.bank.nb):
        assert torch.equal(cg[k // G, :, k % G], layer.bank.codes[:, k])


@pytest.mark.parametrize("conv,train,recovery", [("SAGE", True, True), ("GCN", True, True), ("GAT", True, True),
                                                 ("SAGE", False, True), ("GCN", True, False)])
def test_device_plan_builder_matches_torch_builder(conv, train, recovery):
    """csrc/plan.cu (vqgnn_plan_v1_build) against the torch builder: same tail / in-batch / transposed structure
    (as dense matrices: the order inside a row is free), same counts; power-law graph with hub rows."""
    from vq_gnn_b200 import graph as G
    dev = torch.device("cuda:0")
    N, B = 2000, 300
    g = H.make_graph(N, 60_000, conv, "v1", seed=31, power_law=1.3)
    bA = H.batch_to(H.make_batch(g, B, "v1", seed=31, train=train, recovery=recovery), dev)
    p_t = G.plan_from_v1(bA, conv, N, train, dev)
    p_d = G.plan_from_v1_device(bA, conv, N, train, dev)

    def dense(plan):
        deg = (plan.fwd_rowptr[1:] - plan.fwd_rowptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(B, device=dev), deg)
        a = torch.zeros(B, B + N, device=dev, dtype=torch.float64)
        a.index_put_((rows, plan.fwd_col.long()), plan.fwd_val.double(), accumulate=True)
        b = torch.zeros(B, B + N, device=dev, dtype=torch.float64)
        b.index_put_((rows, plan.fwd_col.long()), plan.fwd_rval.double(), accumulate=True)
        bdeg = (plan.bwd_rowptr[1:] - plan.bwd_rowptr[:-1]).long()
        bj = torch.repeat_interleave(torch.arange(B, device=dev), bdeg)
        t = torch.zeros(B, B, device=dev, dtype=torch.float64)
        t.index_put_((plan.bwd_col.long(), bj), plan.bwd_val.double(), accumulate=True)
        return a, b, t
    for x, y in zip(dense(p_t), dense(p_d)):
        assert torch.equal(x, y)
    sp = p_d.split_v1()
    n_tail = int(sp['tail'][6].item())
    assert n_tail == int((p_t.fwd_col >= B).sum())
    assert int(sp['tail'][0][-1]) == n_tail and torch.equal(p_t.bwd_rowptr, p_d.bwd_rowptr)


@pytest.mark.parametrize("C", [8, 5, 132])
def test_plain_conv_forward_matches_oracle(C):
    """OurGCNConv / OurGATConv.forward on an explicit adjacency (no codeword rows): convs.py:65-101, 165-266."""
    dev = torch.device("cuda:0")
    n = 300
    g = H.make_graph(n, 3000, "GAT", "v2", seed=5)
    from vq_gnn_b200.graph import CSRAdj
    adj = CSRAdj(g.rowptr, g.col, g.val, (n, n))
    x = torch.randn(n, C, generator=torch.Generator().manual_seed(1))
    dense = adj.to_dense()
    gcn = V.OurGCNConv(C, C, normalize=False)
    assert H.rel_err(gcn(x.to(dev), adj.to(dev)), restate.gcn_propagate(dense, x)) < REL_TOL
    torch.manual_seed(2)
    gat = V.OurGATConv(C, C, bias=False, add_self_loops=False).to(dev)
    want = restate.gat_propagate(dense, x, gat.att_l.detach().cpu().view(-1), gat.att_r.detach().cpu().view(-1))
    assert H.rel_err(gat(x.to(dev), adj.to(dev)), want) < REL_TOL


def test_codes_apply_updates_last_entry_wins():
    """vqgnn_codes_apply_updates (the multi-GPU code-table update) vs its torch restatement, with repeated nodes."""
    from vq_gnn_b200 import _lib, dist as vdist
    dev = torch.device("cuda:0")
    lib, st = _lib.load(), _lib.stream()
    N, nb, n, k0, nbc, G = 5000, 14, 6000, 0, 14, 6
    gen = torch.Generator().manual_seed(3)
    nodes = torch.randint(0, N, (n,), generator=gen, dtype=torch.int32)          # many repeats
    new = torch.randint(0, 1024, (n, nbc), generator=gen, dtype=torch.int16)
    codes0 = torch.randint(0, 1024, (N, nb), generator=gen, dtype=torch.int16)
    want = codes0.clone()
    vdist.apply_code_updates_(want, nodes, new, k0)
    codes = codes0.clone().to(dev)
    ng = (nb + G - 1) // G
    codes_g = torch.zeros(ng, N, 8, dtype=torch.int16, device=dev)
    _lib.check(lib.vqgnn_codes_group(_lib.ptr(codes), nb, None, N, N, G, _lib.ptr(codes_g), st))
    owner = torch.empty(N, dtype=torch.int32, device=dev)
    nd, nw = nodes.to(dev), new.to(dev)
    _lib.check(lib.vqgnn_codes_apply_updates(_lib.ptr(nd), _lib.ptr(nw), n, nbc, k0, _lib.ptr(codes), nb, N,
                                             _lib.ptr(codes_g), G, _lib.ptr(owner), st))
    assert torch.equal(codes.cpu(), want)
    for k in range(nb):
        assert torch.equal(codes_g[k // G, :, k % G].cpu(), want[:, k])


@pytest.mark.parametrize("conv", ["GCN", "GAT"])
def test_v2_device_plan_matches_torch_builder(conv):
    """vqgnn_csr_transpose_lt (deferred count read) against the torch builder of the v2 plan."""
    from vq_gnn_b200 import graph as G
    dev = torch.device("cuda:0")
    N, B = 3000, 500
    g = H.make_graph(N, 40_000, conv, "v2", seed=13, power_law=1.5)
    bA = H.batch_to(H.make_batch(g, B, "v2", seed=13, train=True), dev)
    p_t = G.plan_from_v2(bA, conv, N, True, dev)
    p_d = G.plan_from_v2_device(bA, conv, N, dev)
    assert p_d._bwd is None                                   # nothing read back yet
    assert torch.equal(p_t.fwd_rowptr, p_d.fwd_rowptr) and torch.equal(p_t.fwd_col, p_d.fwd_col)
    assert torch.equal(p_t.tail_node, p_d.tail_node) and p_t.R == p_d.R

    def dense_bwd(plan):
        bdeg = (plan.bwd_rowptr[1:] - plan.bwd_rowptr[:-1]).long()
        bj = torch.repeat_interleave(torch.arange(B, device=dev), bdeg)
        t = torch.zeros(plan.R, B, device=dev, dtype=torch.float64)
        return t.index_put_((plan.bwd_col.long(), bj), plan.bwd_val.double(), accumulate=True)
    assert torch.equal(dense_bwd(p_t), dense_bwd(p_d))
    assert torch.equal(p_t.bwd_rowptr, p_d.bwd_rowptr) and p_t.bwd_col.numel() == p_d.bwd_col.numel()


@pytest.mark.parametrize("version,conv", [("v2", "GCN"), ("v1", "SAGE")])
def test_side_stream_vq_updates_are_bit_identical(version, conv):
    """VQBank.async_update: the hook's update runs on a side stream and re-joins before the layer's next forward;
    training must be bit-identical to the in-line update (the path has no order-dependent float accumulation)."""
    import torch.nn.functional as F
    dev = torch.device("cuda:0")
    N, B, M, C = 700, 160, 16, 12
    g = H.make_graph(N, 6000, conv, version, seed=9)
    bAs = [H.batch_to(H.make_batch(g, B, version, seed=s), dev) for s in range(3)]
    xs = [torch.randn(B, C, generator=torch.Generator().manual_seed(s)).to(dev) for s in range(3)]
    ys = [torch.randint(0, 7, (B,), generator=torch.Generator().manual_seed(20 + s)).to(dev) for s in range(3)]

    def run(async_flag):
        torch.manual_seed(0)
        m = V.LowRankGNN(C, 16, 7, 3, 0., M, 4, N, no_second_fc=True, skip=False, commitment_cost=0.,
                         grad_scale=[1, 1], act='leaky_gelu', bn_flag=True, warm_up_flag=True, conv_type=conv,
                         version=version).to(dev).train()
        m.set_async_vq_updates(async_flag)
        opt = torch.optim.RMSprop(m.parameters(), lr=1e-3)
        losses = []
        for i in range(6):
            if i == 1:
                m.set_inited(True)
            opt.zero_grad()
            out, _, info = m((xs[i % 3], bAs[i % 3]), 1)
            loss = F.cross_entropy(out, ys[i % 3]) + info
            loss.backward()
            opt.step()
            losses.append(float(loss))
        m.join_vq_updates()
        torch.cuda.synchronize()
        return losses, {k: v.clone() for k, v in m.state_dict().items()}
    l0, s0 = run(False)
    l1, s1 = run(True)
    assert l0 == l1
    for k in s0:
        assert torch.equal(s0[k], s1[k]), k


@pytest.mark.parametrize("conv,C,slab", [("GCN", 8, 16), ("SAGE", 52, 16), ("GCN", 128, 32), ("GCN", 100, 64)])
def test_v2_split_info_kernel(conv, C, slab, monkeypatch):
    """v2 training with the out-of-batch rows routed through csrc/mp_info.cu (slab-major tables, SDDMM-shaped
    info_backward) and the batch rows through the generic kernel: same outputs / info / gradients / state as the
    oracle; power-law graph (hub rows, empty rows), C not a multiple of the slab."""
    from vq_gnn_b200 import models as Mo
    monkeypatch.setattr(Mo, "INFO_SLAB", slab)
    dev = torch.device("cuda:0")
    N, B, M, D = 1500, 200, 32, 4
    g = H.make_graph(N, 30_000, conv, "v2", seed=33, power_law=1.4)
    batch_A = H.make_batch(g, B, "v2", seed=33)
    torch.manual_seed(17)
    layer = V.LowRankGNNLayer(*H.layer_args(C, 6, M, D, N, conv), version="v2")
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    o = restate.OracleLayer(C, 6, M, D, N, conv, "v2", warm_up_flag=True).load_state_dict(sd)
    layer = layer.to(dev)
    layer.split_info = 'force'
    x = torch.randn(B, C, generator=torch.Generator().manual_seed(3))
    c_outs = _run_cuda(layer, batch_A, x, 3, dev, wu=0.8)
    _compare(c_outs, _run_oracle(o, batch_A, x, 3, wu=0.8, cuda_outs=c_outs), layer, o)
    assert abs(float(c_outs[-1][1])) > 0


@pytest.mark.parametrize("conv,C,power_law", [("GCN", 128, 1.4), ("SAGE", 100, 1.2), ("GCN", 64, 0.0),
                                             ("GCN", 260, 1.4), ("SAGE", 52, 1.4), ("GCN", 16, 0.0)])
def test_v2_row_gather_forward(conv, C, power_law, monkeypatch):
    """v2 layers with materialised out-of-batch rows through the lean asynchronous row-gather kernel (csrc/mp_rows.cuh,
    vqgnn_mp_fwd_rows): same outputs / info / gradients / state as the oracle over three train steps; power-law graphs
    (hub rows cut by several chunk boundaries, empty rows), C below / above one 128-column slab and not a multiple of
    it; and bit-identical y to the generic kernel it replaces (same per-row order of additions)."""
    from vq_gnn_b200 import models as Mo
    dev = torch.device("cuda:0")
    N, B, M, D = 3000, 300, 32, 4
    g = H.make_graph(N, 60_000, conv, "v2", seed=35, power_law=power_law)
    batch_A = H.make_batch(g, B, "v2", seed=35)
    torch.manual_seed(19)
    layer = V.LowRankGNNLayer(*H.layer_args(C, 6, M, D, N, conv), version="v2")
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    o = restate.OracleLayer(C, 6, M, D, N, conv, "v2", warm_up_flag=True).load_state_dict(sd)
    layer = layer.to(dev)
    layer.materialize_tail = 'force'
    x = torch.randn(B, C, generator=torch.Generator().manual_seed(3))
    lc0 = V._lib.launch_count()
    monkeypatch.setattr(Mo, "USE_ROWS_KERNEL", True)
    c_outs = _run_cuda(layer, batch_A, x, 3, dev, wu=0.8)
    assert V._lib.launch_count() > lc0
    _compare(c_outs, _run_oracle(o, batch_A, x, 3, wu=0.8, cuda_outs=c_outs), layer, o)
    assert abs(float(c_outs[-1][1])) > 0
    # against the generic kernel on the same state: y bit-identical, info within fp32 summation noise
    plan = V.build_plan(H.batch_to(batch_A, dev), conv, N, True, dev)
    xd = x.to(dev)
    from vq_gnn_b200.models import VQConvFunction
    y1, i1 = VQConvFunction.apply(xd, None, layer, plan, 0.8, False)
    monkeypatch.setattr(Mo, "USE_ROWS_KERNEL", False)
    y0, i0 = VQConvFunction.apply(xd, None, layer, plan, 0.8, False)
    assert torch.equal(y0, y1)
    assert abs(float(i0) - float(i1)) <= 1e-5 * max(abs(float(i0)), 1e-6), (float(i0), float(i1))
    # eval mode (R == B, no info)
    layer.eval()
    plan_e = V.build_plan(H.batch_to(H.make_batch(g, B, "v2", seed=35, train=False), dev), conv, N, False, dev)
    monkeypatch.setattr(Mo, "USE_ROWS_KERNEL", True)
    ye1, _ = VQConvFunction.apply(xd, None, layer, plan_e, 1.0, False)
    monkeypatch.setattr(Mo, "USE_ROWS_KERNEL", False)
    ye0, _ = VQConvFunction.apply(xd, None, layer, plan_e, 1.0, False)
    assert torch.equal(ye0, ye1)


@pytest.mark.parametrize("C,power_law", [(128, 1.4), (100, 0.0), (260, 1.2), (52, 1.4)])
def test_v2_gat_row_gather_forward(C, power_law, monkeypatch):
    """v2 GAT with materialised codeword rows through the row-gather kernel with per-entry GAT weights
    (vqgnn_gat_fwd_rows): outputs / info / gradients / state against the oracle over three train steps on power-law
    graphs (rows cut by chunk boundaries), and against the generic GAT kernel on the same state."""
    from vq_gnn_b200 import models as Mo
    from vq_gnn_b200.gat import VQGATFunction
    dev = torch.device("cuda:0")
    N, B, M, D = 3000, 300, 32, 4
    g = H.make_graph(N, 60_000, "GAT", "v2", seed=37, power_law=power_law)
    batch_A = H.make_batch(g, B, "v2", seed=37)
    torch.manual_seed(23)
    layer = V.LowRankGNNLayer(*H.layer_args(C, 6, M, D, N, "GAT"), version="v2")
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    o = restate.OracleLayer(C, 6, M, D, N, "GAT", "v2", warm_up_flag=True).load_state_dict(sd)
    layer = layer.to(dev)
    layer.materialize_tail = 'force'
    monkeypatch.setattr(Mo, "USE_ROWS_KERNEL", True)
    x = torch.randn(B, C, generator=torch.Generator().manual_seed(3))
    c_outs = _run_cuda(layer, batch_A, x, 3, dev, wu=0.8)
    _compare(c_outs, _run_oracle(o, batch_A, x, 3, wu=0.8, cuda_outs=c_outs), layer, o)
    plan = V.build_plan(H.batch_to(batch_A, dev), "GAT", N, True, dev)
    conv = layer.conv
    slope = float(conv.negative_slope)
    xd = x.to(dev)
    with torch.no_grad():
        y1, i1 = VQGATFunction.apply(xd, conv.att_l, conv.att_r, layer, plan, 0.8, False, slope)
        monkeypatch.setattr(Mo, "USE_ROWS_KERNEL", False)
        y0, i0 = VQGATFunction.apply(xd, conv.att_l, conv.att_r, layer, plan, 0.8, False, slope)
    assert H.rel_err(y1, y0) < 1e-5
    assert abs(float(i0) - float(i1)) <= 1e-4 * max(abs(float(i0)), 1e-6), (float(i0), float(i1))


def test_v1_tail_kernel_without_in_batch_block():
    """v1 SAGE batch without A_BB (recovery_flag=False: every neighbour goes through its codeword, no self loops):
    the in-batch CSR is EMPTY and the shared-memory tail kernel carries the whole forward (the reference's init()
    feeds such batches, main_node.py:17-37)."""
    dev = torch.device("cuda:0")
    N, B, M, C = 500, 100, 16, 8
    g = H.make_graph(N, 9000, "SAGE", "v1", seed=14)
    batch_A = H.make_batch(g, B, "v1", seed=14, train=True, recovery=False)
    assert batch_A[2] is None
    torch.manual_seed(19)
    layer = V.LowRankGNNLayer(*H.layer_args(C, 6, M, 4, N, "SAGE"), version="v1")
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    o = restate.OracleLayer(C, 6, M, 4, N, "SAGE", "v1", warm_up_flag=True).load_state_dict(sd)
    layer = layer.to(dev)
    layer.use_tail_kernel = 'force'
    x = torch.randn(B, C, generator=torch.Generator().manual_seed(3))
    c_outs = _run_cuda(layer, batch_A, x, 3, dev)
    _compare(c_outs, _run_oracle(o, batch_A, x, 3, cuda_outs=c_outs), layer, o)


def test_prefetched_tail_rows_are_bit_identical():
    """LowRankGNN.prefetch_tails: layers 2..L materialise their out-of-batch codeword rows on a side stream at the start
    of the forward; training must be bit-identical to materialising in line."""
    import torch.nn.functional as F
    dev = torch.device("cuda:0")
    N, B, M, C = 900, 160, 16, 128
    g = H.make_graph(N, 9000, "GCN", "v2", seed=9)
    bAs = [H.batch_to(H.make_batch(g, B, "v2", seed=s), dev) for s in range(3)]
    xs = [torch.randn(B, C, generator=torch.Generator().manual_seed(s)).to(dev) for s in range(3)]
    ys = [torch.randint(0, 7, (B,), generator=torch.Generator().manual_seed(20 + s)).to(dev) for s in range(3)]

    def run(prefetch, async_vq):
        torch.manual_seed(0)
        m = V.LowRankGNN(C, 64, 7, 3, 0., M, 4, N, no_second_fc=True, skip=False, commitment_cost=0.,
                         grad_scale=[1, 1], act='leaky_gelu', bn_flag=True, warm_up_flag=True, conv_type="GCN",
                         version="v2").to(dev).train()
        m.prefetch_tails = prefetch
        m.set_async_vq_updates(async_vq)
        for layer in m.convs:
            layer.materialize_tail = 'force'
        opt = torch.optim.RMSprop(m.parameters(), lr=1e-3)
        losses = []
        for i in range(6):
            if i == 1:
                m.set_inited(True)
            opt.zero_grad()
            out, _, info = m((xs[i % 3], bAs[i % 3]), 1)
            loss = F.cross_entropy(out, ys[i % 3]) + info
            loss.backward()
            opt.step()
            losses.append(float(loss))
        m.join_vq_updates()
        torch.cuda.synchronize()
        return losses, {k: v.clone() for k, v in m.state_dict().items()}
    l0, s0 = run(False, False)
    for cfg in ((True, False), (True, True)):
        l1, s1 = run(*cfg)
        assert l0 == l1, cfg
        for k in s0:
            assert torch.equal(s0[k], s1[k]), (cfg, k)
