"""GPU parity against the frozen outputs of the UNMODIFIED reference (tests/golden/*.npz): the CUDA
layers (through the C-ABI) are driven with the reference's own state dict, batch and inputs and must
reproduce its outputs, gradients, codes and post-update quantiser state (fp32: 1e-4 relative;
codes: bit-exact)."""
import pytest
import torch

import vq_gnn_b200 as V
from tests import helpers as H

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4
LAYER_FILES = ["layer_v2_gcn", "layer_v2_sage", "layer_v2_gat", "layer_v1_gcn", "layer_v1_sage", "layer_v1_gat",
               "layer_v1_sage_wide", "layer_v2_gat_wide"]
VQ_FILES = ["vq_m16", "vq_m32_add", "vq_m64"]


@pytest.mark.parametrize("name", VQ_FILES)
def test_cuda_vq_matches_golden(name):
    dev = torch.device("cuda:0")
    z = H.load_golden(name)
    M, D, add = int(z["cfg.M"]), int(z["cfg.D"]), bool(int(z["cfg.add_flag"]))
    vq = V.VectorQuantizerEMA(M, D, grad_normalize_scale=[1, 0.5], warm_up_flag=True, momentum=0.1, add_flag=add)
    vq.load_state_dict(H.golden_sd(z, "sd0."))
    vq = vq.to(dev).train()
    for s in range(3):
        idx = vq.feature_update(torch.from_numpy(z[f"f{s}.x"]).to(dev))
        assert torch.equal(idx.cpu(), torch.from_numpy(z[f"f{s}.idx"]))
    for s in range(3):
        idx, _ = vq.update(torch.from_numpy(z[f"u{s}.x"]).to(dev), torch.from_numpy(z[f"u{s}.g"]).to(dev))
        assert torch.equal(idx.cpu(), torch.from_numpy(z[f"u{s}.idx"]))
        bad, _ = H.state_mismatches(vq.state_dict(), H.golden_sd(z, f"u{s}.sd."), REL_TOL)
        assert not bad, (s, bad)


@pytest.mark.parametrize("name", LAYER_FILES)
def test_cuda_layer_matches_golden(name):
    dev = torch.device("cuda:0")
    z = H.load_golden(name)
    version, conv = str(z["meta.version"]), str(z["meta.conv"])
    cfg = {k[4:]: int(z[k]) for k in z.files if k.startswith("cfg.")}
    layer = V.LowRankGNNLayer(*H.layer_args(cfg["C"], cfg["C_out"], cfg["M"], cfg["D"], cfg["N"], conv,
                                            skip=bool(cfg["skip"])),
                              version=version, literal_v2_hooks=(version == "v2"))
    layer.load_state_dict(H.golden_sd(z, "sd0."))
    layer = layer.to(dev).train()
    bA = H.batch_to(H.unpack_batch(z), dev)
    x = torch.from_numpy(z["x"])
    for s in range(int(z["meta.steps"])):
        if s == 1:
            layer.set_inited(True)
        xx = x.clone().to(dev).requires_grad_(True)
        for p in layer.parameters():
            p.grad = None
        out = layer(xx, bA, 1, False)
        ((out[0] * H.loss_weights(out[0].shape).to(dev)).sum() + out[5]).backward()
        assert H.rel_err(out[0], torch.from_numpy(z[f"step{s}.out"])) < REL_TOL, (s, "out")
        ri = float(z[f"step{s}.info"][0])
        assert abs(float(out[5]) - ri) <= REL_TOL * max(1e-3, abs(ri)), (s, "info", float(out[5]), ri)
        assert H.rel_err(xx.grad, torch.from_numpy(z[f"step{s}.dx"])) < REL_TOL, (s, "dx")
        got, want = {}, {}
        for k, p in layer.named_parameters():
            gk = f"step{s}.grad.{k}"
            if p.grad is not None and gk in z.files:
                got[k], want[k] = p.grad, torch.from_numpy(z[gk])
                if "att_" not in k:
                    assert H.rel_err(p.grad, want[k]) < REL_TOL, (s, k)
        assert H.att_grad_err(got, want) < REL_TOL, (s, "att grads")
    bad, n_codes = H.state_mismatches(layer.state_dict(), H.golden_sd(z, "sd1."), REL_TOL)
    assert not bad and n_codes == 0, (bad, n_codes)
