"""Pins the CPU restatement (oracle/restate.py) against the UNMODIFIED reference, executed live through
oracle/ref_loader.py.  Runs only where /root/reference exists (the builder container); on the GPU box
the same comparisons are frozen in tests/golden/ (see test_oracle_golden.py)."""
import pytest
import torch

from oracle import ref_loader, restate
from tests import helpers as H

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.mark.parametrize("M,D,B,add_flag,warm", [(16, 4, 200, False, True), (64, 4, 500, False, False),
                                                (32, 4, 300, True, True), (8, 2, 100, False, True)])
def test_vq_matches_reference(M, D, B, add_flag, warm):
    ref = ref_loader.load_reference("v2")
    torch.manual_seed(1)
    r = ref.vq.VectorQuantizerEMA(M, D, grad_normalize_scale=[1, 0.5], warm_up_flag=warm, momentum=0.1,
                                  add_flag=add_flag)
    torch.manual_seed(1)
    o = restate.OracleVQ(M, D, grad_normalize_scale=[1, 0.5], warm_up_flag=warm, momentum=0.1, add_flag=add_flag)
    assert torch.equal(r._embedding, o._embedding) and torch.equal(r._ema_w, o._ema_w)
    g = torch.Generator().manual_seed(2)
    if not warm:   # without Laplace smoothing an empty cluster is 'Bad Init!' in both (vq.py:188)
        X = torch.randn(B, D, generator=g)
        with pytest.raises(ValueError):
            r.feature_update(X)
        with pytest.raises(ValueError):
            o.feature_update(X)
        return
    for step in range(3):
        X = torch.randn(B, D, generator=g) * 2 + 0.3
        i_r, i_o = r.feature_update(X), o.feature_update(X)
        assert torch.equal(i_r, i_o)
    for step in range(3):
        X = torch.randn(B, D, generator=g) * 2 + 0.3
        G = torch.randn(B, D + int(add_flag), generator=g) * 1e-3
        (i_r, _), (i_o, _) = r.update(X, G), o.update(X, G)
        assert torch.equal(i_r, i_o)
    sd = r.state_dict()
    for k, v in o.dump().items():
        assert torch.allclose(sd[k], v, rtol=1e-6, atol=1e-7), k


def _run_ref_layer(ref, version, conv, sd, args, batch_A, x, steps, ts, x_grad=True):
    layer = ref.models.LowRankGNNLayer(*args)
    layer.load_state_dict(sd)
    layer.train()
    outs = []
    bA = H.to_shim_batch(batch_A, ts)
    for s in range(steps):
        if s == 1:
            for b in layer.gnn_block:
                b.inited = True
        xx = x.clone().requires_grad_(x_grad)
        for p in layer.parameters():
            p.grad = None
        out = layer(xx, bA, 1, False)
        loss = (out[0] * H.loss_weights(out[0].shape)).sum() + out[5]
        loss.backward()
        outs.append((out[0].detach(), torch.as_tensor(out[5]).detach(),
                     xx.grad.clone() if x_grad else torch.zeros(1),
                     {k: p.grad.clone() for k, p in layer.named_parameters() if p.grad is not None}))
    return outs, layer.state_dict()


def _run_oracle_layer(version, conv, sd, cfg, batch_A, x, steps, hook_mode, x_grad=True):
    o = restate.OracleLayer(cfg["C"], cfg["C_out"], cfg["M"], cfg["D"], cfg["N"], conv, version, skip=cfg["skip"],
                            warm_up_flag=True, hook_mode=hook_mode).load_state_dict(sd)
    o.train()
    outs = []
    for s in range(steps):
        if s == 1:
            o.set_inited(True)
        xx = x.clone().requires_grad_(x_grad)
        for p in o.params.values():
            p.grad = None
        out, info = o(xx, batch_A, 1.0, False)
        loss = (out * H.loss_weights(out.shape)).sum() + info
        loss.backward()
        outs.append((out.detach(), torch.as_tensor(info).detach(),
                     xx.grad.clone() if x_grad else torch.zeros(1),
                     {k: p.grad.clone() for k, p in o.params.items() if p.grad is not None}))
    return outs, o.state_dict()


@pytest.mark.parametrize("x_grad", [True, False])
@pytest.mark.parametrize("version,conv", [("v2", "GCN"), ("v2", "SAGE"), ("v2", "GAT"), ("v1", "GCN"),
                                          ("v1", "SAGE"), ("v1", "GAT")])
def test_layer_matches_reference(version, conv, x_grad):
    ref = ref_loader.load_reference(version)
    ts = ref_loader.shim_sparse()
    cfg = dict(N=300, B=80, C=8, C_out=6, M=16, D=4, skip=(conv == "GAT"))
    g = H.make_graph(cfg["N"], 1200, conv, version, seed=7)
    batch_A = H.make_batch(g, cfg["B"], version, seed=7)
    torch.manual_seed(11)
    args = H.layer_args(cfg["C"], cfg["C_out"], cfg["M"], cfg["D"], cfg["N"], conv, skip=cfg["skip"])
    sd = ref.models.LowRankGNNLayer(*args).state_dict()
    x = torch.randn(cfg["B"], cfg["C"], generator=torch.Generator().manual_seed(3))
    steps = 4
    r_outs, r_sd = _run_ref_layer(ref, version, conv, sd, args, batch_A, x, steps, ts, x_grad)
    # the live v2 reference never fires its hook (dangling slice)
    o_outs, o_sd = _run_oracle_layer(version, conv, sd, cfg, batch_A, x, steps,
                                     "literal_v2" if version == "v2" else "fire", x_grad)
    for s, ((ro, ri, rg, rp), (oo, oi, og, op)) in enumerate(zip(r_outs, o_outs)):
        assert H.rel_err(oo, ro) < 2e-5, (s, "out")
        assert abs(float(oi) - float(ri)) <= 2e-5 * max(1.0, abs(float(ri))), (s, "info", float(oi), float(ri))
        assert H.rel_err(og, rg) < 2e-5, (s, "dx")
        for k in rp:
            assert H.rel_err(op[k], rp[k]) < 2e-5, (s, k)
    for k, v in o_sd.items():
        if v.is_floating_point():
            assert torch.allclose(r_sd[k], v, rtol=1e-4, atol=1e-6), k
        else:
            assert torch.equal(r_sd[k], v), k
    if version == "v1":   # the hook must really have fired: gradient codewords are non-zero after step 2
        assert float(r_sd["gnn_block.0.vq._embedding_output"][:, cfg["D"]:].abs().sum()) > 0
        assert float(torch.as_tensor(r_outs[-1][1]).abs()) > 0
