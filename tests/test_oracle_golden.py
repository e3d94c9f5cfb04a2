"""The CPU restatement (oracle/restate.py) against the frozen outputs of the UNMODIFIED reference
(tests/golden/*.npz, written by oracle/gen_golden.py in the builder container).  Runs everywhere: this
is what pins the oracle on the GPU box, where /root/reference does not exist."""
import pytest
import torch

from oracle import restate
from tests import helpers as H

LAYER_FILES = ["layer_v2_gcn", "layer_v2_sage", "layer_v2_gat", "layer_v1_gcn", "layer_v1_sage", "layer_v1_gat",
               "layer_v1_sage_wide", "layer_v2_gat_wide"]
VQ_FILES = ["vq_m16", "vq_m32_add", "vq_m64"]


@pytest.mark.parametrize("name", VQ_FILES)
def test_oracle_vq_matches_golden(name):
    z = H.load_golden(name)
    M, D, add = int(z["cfg.M"]), int(z["cfg.D"]), bool(int(z["cfg.add_flag"]))
    o = restate.OracleVQ(M, D, grad_normalize_scale=[1, 0.5], warm_up_flag=True, momentum=0.1, add_flag=add,
                         init_random=False).load(H.golden_sd(z, "sd0."))
    for s in range(3):
        idx = o.feature_update(torch.from_numpy(z[f"f{s}.x"]))
        assert torch.equal(idx, torch.from_numpy(z[f"f{s}.idx"]))       # bit-exact codes
    for s in range(3):
        idx, _ = o.update(torch.from_numpy(z[f"u{s}.x"]), torch.from_numpy(z[f"u{s}.g"]))
        assert torch.equal(idx, torch.from_numpy(z[f"u{s}.idx"]))
        ref_sd = H.golden_sd(z, f"u{s}.sd.")
        for k, v in o.dump().items():
            assert torch.allclose(ref_sd[k], v, rtol=2e-5, atol=2e-6), (s, k)   # summation order (GEMM vs index_add)


@pytest.mark.parametrize("name", LAYER_FILES)
def test_oracle_layer_matches_golden(name):
    z = H.load_golden(name)
    version, conv = str(z["meta.version"]), str(z["meta.conv"])
    cfg = {k[4:]: int(z[k]) for k in z.files if k.startswith("cfg.")}
    batch_A = H.unpack_batch(z)
    x = torch.from_numpy(z["x"])
    # the live v2 reference never fires its hook (dangling slice)
    o = restate.OracleLayer(cfg["C"], cfg["C_out"], cfg["M"], cfg["D"], cfg["N"], conv, version,
                            skip=bool(cfg["skip"]), warm_up_flag=True,
                            hook_mode="literal_v2" if version == "v2" else "fire")
    o.load_state_dict(H.golden_sd(z, "sd0.")).train()
    for s in range(int(z["meta.steps"])):
        if s == 1:
            o.set_inited(True)
        xx = x.clone().requires_grad_(True)
        for p in o.params.values():
            p.grad = None
        out, info = o(xx, batch_A, 1.0, False)
        ((out * H.loss_weights(out.shape)).sum() + info).backward()
        assert H.rel_err(out, torch.from_numpy(z[f"step{s}.out"])) < 2e-5, (s, "out")
        ri = float(z[f"step{s}.info"][0])
        assert abs(float(info) - ri) <= 2e-5 * max(1.0, abs(ri)), (s, "info")
        assert H.rel_err(xx.grad, torch.from_numpy(z[f"step{s}.dx"])) < 2e-5, (s, "dx")
        for k, p in o.params.items():
            gk = f"step{s}.grad.{k}"
            if p.grad is not None and gk in z.files:
                assert H.rel_err(p.grad, torch.from_numpy(z[gk])) < 2e-5, (s, k)
    ref_sd = H.golden_sd(z, "sd1.")
    for k, v in o.state_dict().items():
        if v.is_floating_point():
            assert torch.allclose(ref_sd[k], v, rtol=1e-4, atol=1e-6), k
        else:
            assert torch.equal(ref_sd[k], v), k
