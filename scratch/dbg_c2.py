import sys; sys.path.insert(0, '/root/repo')
import torch
import vq_gnn_b200 as V
from tests import helpers as H, torch_ref as R
from tests.test_gpu_fullsize import _layer, _warm
from vq_gnn_b200 import sampling, synth
from vq_gnn_b200.models import VQConvFunction
dev = torch.device("cuda:0")
s = synth.CONFIG_SHAPES["c2_reddit"]
g = synth.make_graph(s["N"], s["E"], "SAGE", "v1", seed=0, power_law=2.2, device=dev)
gen = torch.Generator(device=dev).manual_seed(3)
seeds = torch.randperm(s["N"], generator=gen, device=dev)[:6000]
nodes = sampling.cont_sampler(g, seeds, 3, 6000, generator=gen)[2]
bA = sampling.collate_batch_v1(g, nodes, True, True)
layer = _layer(128, 128, 1024, s["N"], "SAGE", "v1", dev)
plan = V.build_plan(bA, "SAGE", s["N"], True, dev)
x = torch.randn(nodes.numel(), 128, device=dev, generator=torch.Generator(device=dev).manual_seed(2))
_warm(layer, x, plan)
B = plan.B
row, col, val = R._coo(plan)
tail = col >= B
A_in = torch.sparse_coo_tensor(torch.stack([row[~tail], col[~tail]]), val[~tail], (B, B)).coalesce()
w = torch.randn(x.shape, device=dev, generator=gen)
for mode in (False, 'force'):
    layer.use_tail_kernel = mode
    xc = x.clone().requires_grad_(True)
    y, info = VQConvFunction.apply(xc, None, layer, plan, 0.0, False)     # wu = 0: only the in-batch part
    (y * w).sum().backward()
    y_ref = torch.sparse.mm(A_in, x)
    dx_ref = torch.sparse.mm(A_in.t(), w)
    print(mode, "y in-batch rel", H.rel_err(y, y_ref), "dx rel", H.rel_err(xc.grad, dx_ref),
          "max|y_ref|", float(y_ref.abs().max()), "max|dx_ref|", float(dx_ref.abs().max()))
    bad = ((y - y_ref).abs().max(1).values > 1e-6).nonzero().flatten()
    print("  rows with y error:", bad.numel(), bad[:10].tolist())
    badx = ((xc.grad - dx_ref).abs().max(1).values > 1e-6).nonzero().flatten()
    print("  rows with dx error:", badx.numel(), badx[:10].tolist())
# duplicates in A_BB?
r, c, v = bA[2]
key = r * B + c
print("A_BB entries", key.numel(), "unique", torch.unique(key).numel())
rn, cn, vn = bA[1]
keyn = rn * s["N"] + cn
print("A_BN entries", keyn.numel(), "unique", torch.unique(keyn).numel(), "batch nodes unique", torch.unique(nodes).numel())
