"""Freeze outputs of the UNMODIFIED reference into tests/golden/*.npz.

TEST INFRASTRUCTURE ONLY.  Run in the builder container (where /root/reference exists):
    python -m oracle.gen_golden
The reference ships no tests or golden vectors, so these fixtures are produced by
executing the reference's own `vq.py` / `convs.py` / `models.py` / `mapper` (through oracle/ref_loader.py)
on small seeded inputs.  Each file holds the inputs (reference-keyed state dict, batch_A, x), the
per-step outputs (layer output, info_backward, d loss/d x, parameter gradients) and the final state
dict.  `tests/test_oracle_golden.py` checks the CPU restatement (oracle/restate.py) against them on
every box, and the `-m gpu` tests check the CUDA path against the same files.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tests import helpers as H  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
STEPS = 4
LAYER_CASES = [  # name, version, conv, cfg
    ("layer_v2_gcn", "v2", "GCN", dict(N=300, B=80, C=8, C_out=6, M=16, D=4, skip=False, E=1200)),
    ("layer_v2_sage", "v2", "SAGE", dict(N=300, B=80, C=8, C_out=6, M=16, D=4, skip=False, E=1200)),
    ("layer_v2_gat", "v2", "GAT", dict(N=300, B=80, C=8, C_out=6, M=16, D=4, skip=True, E=1200)),
    ("layer_v1_gcn", "v1", "GCN", dict(N=300, B=80, C=8, C_out=6, M=16, D=4, skip=False, E=1200)),
    ("layer_v1_sage", "v1", "SAGE", dict(N=300, B=80, C=8, C_out=6, M=16, D=4, skip=False, E=1200)),
    ("layer_v1_gat", "v1", "GAT", dict(N=300, B=80, C=8, C_out=6, M=16, D=4, skip=True, E=1200)),
    ("layer_v1_sage_wide", "v1", "SAGE", dict(N=500, B=150, C=132, C_out=16, M=32, D=4, skip=False, E=6000)),
    ("layer_v2_gat_wide", "v2", "GAT", dict(N=400, B=120, C=52, C_out=12, M=32, D=4, skip=True, E=2500)),
]
VQ_CASES = [("vq_m16", 16, 4, 200, False), ("vq_m32_add", 32, 4, 300, True), ("vq_m64", 64, 4, 1000, False)]


def gen_layer(name, version, conv, cfg):
    ref = ref_loader.load_reference(version)
    ts = ref_loader.shim_sparse()
    g = H.make_graph(cfg["N"], cfg["E"], conv, version, seed=7)
    batch_A = H.make_batch(g, cfg["B"], version, seed=7)
    torch.manual_seed(11)
    args = H.layer_args(cfg["C"], cfg["C_out"], cfg["M"], cfg["D"], cfg["N"], conv, skip=cfg["skip"])
    layer = ref.models.LowRankGNNLayer(*args)
    sd0 = {k: v.clone() for k, v in layer.state_dict().items()}
    layer.train()
    x = torch.randn(cfg["B"], cfg["C"], generator=torch.Generator().manual_seed(3))
    bA = H.to_shim_batch(batch_A, ts)
    rec = {"meta.version": version, "meta.conv": conv, "meta.steps": STEPS, "x": x}
    rec.update({f"cfg.{k}": v for k, v in cfg.items()})
    rec.update({f"sd0.{k}": v for k, v in sd0.items()})
    rec.update(H.pack_batch(batch_A))
    for s in range(STEPS):
        if s == 1:
            for b in layer.gnn_block:
                b.inited = True
        xx = x.clone().requires_grad_(True)
        for p in layer.parameters():
            p.grad = None
        out = layer(xx, bA, 1, False)
        loss = (out[0] * H.loss_weights(out[0].shape)).sum() + out[5]
        loss.backward()
        rec[f"step{s}.out"] = out[0].detach()
        rec[f"step{s}.info"] = torch.as_tensor(out[5]).detach().float().reshape(1)
        rec[f"step{s}.dx"] = xx.grad.clone()
        for k, p in layer.named_parameters():
            if p.grad is not None:
                rec[f"step{s}.grad.{k}"] = p.grad.clone()
    rec.update({f"sd1.{k}": v for k, v in layer.state_dict().items()})
    return rec


def gen_vq(name, M, D, B, add_flag):
    ref = ref_loader.load_reference("v2")
    torch.manual_seed(1)
    r = ref.vq.VectorQuantizerEMA(M, D, grad_normalize_scale=[1, 0.5], warm_up_flag=True, momentum=0.1,
                                  add_flag=add_flag)
    rec = {"cfg.M": M, "cfg.D": D, "cfg.B": B, "cfg.add_flag": int(add_flag)}
    rec.update({f"sd0.{k}": v.clone() for k, v in r.state_dict().items()})
    g = torch.Generator().manual_seed(2)
    for step in range(3):
        X = torch.randn(B, D, generator=g) * 2 + 0.3
        rec[f"f{step}.x"], rec[f"f{step}.idx"] = X, r.feature_update(X)
    for step in range(3):
        X = torch.randn(B, D, generator=g) * 2 + 0.3
        G = torch.randn(B, D + int(add_flag), generator=g) * 1e-3
        idx, _ = r.update(X, G)
        rec[f"u{step}.x"], rec[f"u{step}.g"], rec[f"u{step}.idx"] = X, G, idx
        rec.update({f"u{step}.sd.{k}": v.clone() for k, v in r.state_dict().items()})
    return rec


def save(name, rec):
    arrs = {}
    for k, v in rec.items():
        if isinstance(v, torch.Tensor):
            arrs[k] = v.detach().cpu().numpy()
        elif isinstance(v, str):
            arrs[k] = np.array(v)
        else:
            arrs[k] = np.array(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB, {len(arrs)} arrays")


def main():
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    for name, version, conv, cfg in LAYER_CASES:
        save(name, gen_layer(name, version, conv, cfg))
    for name, M, D, B, add in VQ_CASES:
        save(name, gen_vq(name, M, D, B, add))


if __name__ == "__main__":
    main()
