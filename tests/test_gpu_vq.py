"""GPU parity: VectorQuantizerEMA (CUDA, through the C-ABI) vs the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): codeword-count histograms bit-exact given identical assignments;
assignments equal except at near-ties (relative distance gap < 1e-5), mismatch rate reported;
floating-point state within 1e-4 relative."""
import pytest
import torch

import vq_gnn_b200 as V
from oracle import restate
from tests import helpers as H

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4        # north_star: "within 1e-4 relative in fp32"
TIE_GAP = 1e-5        # north_star: near-tie allowance


def _mismatch_rate(o):
    return o.forced_mismatches / max(o.forced_total, 1)


@pytest.mark.parametrize("impl", [0, "auto"])
@pytest.mark.parametrize("M,D,B,add_flag", [(16, 4, 300, False), (256, 4, 5000, False), (64, 4, 1000, True),
                                            (1024, 4, 6000, False), (4096, 4, 10000, False), (32, 2, 500, False),
                                            (48, 8, 700, False)])
def test_vq_update_matches_oracle(M, D, B, add_flag, impl):
    """The CUDA path runs first; the oracle then receives its codes (OracleVQ.force_idx): every row where they differ
    from the oracle's own argmin must be a near-tie (relative gap < 1e-5, checked inside the oracle, rate printed),
    and the state comparison that follows sees identical assignments -- no skipped case."""
    if impl == "auto" and D != 4:
        pytest.skip("the tcgen05 assignment is written for num_D = 4")
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    vq = V.VectorQuantizerEMA(M, D, grad_normalize_scale=[1, 0.5], warm_up_flag=True, momentum=0.1,
                              add_flag=add_flag)
    o = restate.OracleVQ(M, D, grad_normalize_scale=[1, 0.5], warm_up_flag=True, add_flag=add_flag,
                         init_random=False).load(vq.state_dict())
    o.tie_gap = TIE_GAP
    vq = vq.to(dev)
    vq.bank.assign_impl = impl
    vq.train()
    g = torch.Generator().manual_seed(2)
    for step in range(3):
        X = torch.randn(B, D, generator=g) * 2 + 0.3
        i_g = vq.feature_update(X.to(dev))
        assert i_g.shape == (B, 1) and i_g.dtype == torch.long
        o.force_idx = i_g.cpu()
        i_o = o.feature_update(X)
        assert torch.equal(i_o, i_g.cpu())
    for step in range(3):
        X = torch.randn(B, D, generator=g) * 2 + 0.3
        G = torch.randn(B, D + int(add_flag), generator=g) * 1e-3
        i_g, enc = vq.update(X.to(dev), G.to(dev))
        o.force_idx = i_g.cpu()
        i_o, _ = o.update(X, G)
        # histogram bit-exact given identical assignments
        assert torch.equal(enc.sum(0).cpu(), torch.bincount(i_o.view(-1), minlength=M).float())
        assert torch.equal(vq.bank.last_stats[0, :, vq.bank.Wp].cpu(),
                           torch.bincount(i_o.view(-1), minlength=M).float())
    print(f"assignment mismatch rate (near-ties only): {_mismatch_rate(o):.2e}")
    assert _mismatch_rate(o) < 1e-3
    sd = {k: v.cpu() for k, v in vq.state_dict().items()}
    for k, v in o.dump().items():
        assert H.rel_err(sd[k], v) < REL_TOL, (k, H.rel_err(sd[k], v))
    assert int(sd["batch_norm_feat.num_batches_tracked"]) == 6
    assert int(sd["batch_norm_grad.num_batches_tracked"]) == 3


def test_vq_update_is_bit_reproducible():
    """Two runs of the same update sequence (fresh modules, different allocation history) give identical bits: the
    statistics are ordered sums (vqgnn_vq_segsum, two-level moments), not float atomics."""
    dev = torch.device("cuda:0")

    def run(pad):
        junk = torch.empty(pad, device=dev)       # shift the allocator so addresses / launch timing differ
        torch.manual_seed(7)
        vq = V.VectorQuantizerEMA(256, 4, grad_normalize_scale=[1, 1], warm_up_flag=True).to(dev).train()
        g = torch.Generator().manual_seed(8)
        for _ in range(4):
            X = torch.randn(20000, 4, generator=g).to(dev)
            G = (torch.randn(20000, 4, generator=g) * 1e-3).to(dev)
            vq.update(X, G)
        del junk
        return {k: v.clone() for k, v in vq.state_dict().items()}
    a, b = run(1), run(1 << 20)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_vq_eval_mode_only_assigns():
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    vq = V.VectorQuantizerEMA(32, 4, grad_normalize_scale=[1, 1], warm_up_flag=True).to(dev)
    X = torch.randn(400, 4, device=dev)
    vq.train()
    vq.feature_update(X)
    before = {k: v.clone() for k, v in vq.state_dict().items()}
    vq.eval()
    o = restate.OracleVQ(32, 4, warm_up_flag=True, init_random=False).load({k: v.cpu() for k, v in before.items()})
    o.training = False
    i_g = vq.feature_update(X)
    o.force_idx = i_g.cpu()
    i_o = o.feature_update(X.cpu())
    assert o.forced_mismatches == 0
    for k, v in vq.state_dict().items():
        assert torch.equal(v, before[k]), k


def test_bad_init_raises():
    dev = torch.device("cuda:0")
    torch.manual_seed(4)
    vq = V.VectorQuantizerEMA(64, 4, grad_normalize_scale=[1, 1], warm_up_flag=False).to(dev)
    vq.train()
    before = {k: v.clone() for k, v in vq.state_dict().items()}
    with pytest.raises(ValueError, match="Bad Init"):      # vq.py:188-189: an empty cluster without smoothing
        vq.feature_update(torch.randn(20, 4, device=dev))
    # the reference raises before it touches _ema_w / _embedding / _embedding_output: no NaN / inf may appear
    after = vq.state_dict()
    for k in ("_ema_w", "_embedding", "_embedding_output"):
        assert torch.equal(after[k], before[k]), k
    assert all(torch.isfinite(v.float()).all() for v in after.values())


def test_bad_init_surfaces_on_the_layer_path_without_a_sync():
    """warm_up_flag=False leaves unused codewords at size 0 on the first update; the layer path polls the status word
    asynchronously (VQBank.poll_status) and raises within a few forwards instead of training on NaN codewords."""
    dev = torch.device("cuda:0")
    N, B, C = 300, 40, 8
    g = H.make_graph(N, 1200, "GCN", "v2", seed=2)
    bA = H.batch_to(H.make_batch(g, B, "v2", seed=2), dev)
    torch.manual_seed(1)
    layer = V.LowRankGNNLayer(*H.layer_args(C, 6, 64, 4, N, "GCN", warm_up_flag=False), version="v2").to(dev).train()
    x = torch.randn(B, C, device=dev)
    with pytest.raises(ValueError, match="Bad Init"):
        for _ in range(4):
            layer(x, bA, 1.0, False)
            torch.cuda.synchronize()
    assert all(torch.isfinite(v.float()).all() for v in layer.state_dict().values())


def test_edge_cases_single_row_and_large_M():
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    vq = V.VectorQuantizerEMA(4096, 4, grad_normalize_scale=[1, 1], warm_up_flag=True)
    o = restate.OracleVQ(4096, 4, warm_up_flag=True, init_random=False).load(vq.state_dict())
    vq = vq.to(dev).train()
    X = torch.randn(2, 4)
    i_g = vq.feature_update(X.to(dev))
    o.force_idx = i_g.cpu()
    i_o = o.feature_update(X)
    assert o.forced_mismatches == 0
    sd = {k: v.cpu() for k, v in vq.state_dict().items()}
    for k, v in o.dump().items():
        assert H.rel_err(sd[k], v) < REL_TOL, k
