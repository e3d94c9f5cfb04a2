import sys; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch
import vq_gnn_b200 as V
from oracle import restate
from tests import helpers as H
dev = torch.device("cuda:0")
version, conv, C, D, skip = "v2", "GAT", 8, 4, True
N, B, M, C_out = 400, 120, 16, 10
g = H.make_graph(N, 2000, conv, version, seed=7)
batch_A = H.make_batch(g, B, version, seed=7)
torch.manual_seed(11)
layer = V.LowRankGNNLayer(*H.layer_args(C, C_out, M, D, N, conv, skip=skip), version=version)
sd = {k: v.clone() for k, v in layer.state_dict().items()}
o = restate.OracleLayer(C, C_out, M, D, N, conv, version, skip=skip, warm_up_flag=True).load_state_dict(sd)
layer = layer.to(dev).train(); o.train()
x = torch.randn(B, C, generator=torch.Generator().manual_seed(3))
bA = H.batch_to(batch_A, dev)
xx = x.clone().to(dev).requires_grad_(True)
out = layer(xx, bA, 0.7, False)
w = H.loss_weights(out[0].shape).to(dev)
((out[0] * w).sum() + out[5]).backward()
xo = x.clone().requires_grad_(True)
oo, oi = o(xo, batch_A, 0.7, False)
((oo * H.loss_weights(oo.shape)).sum() + oi).backward()
print("att_l cuda", layer.conv.att_l.grad.flatten().cpu())
print("att_l orac", o.params["conv.att_l"].grad.flatten())
print("att_r cuda", layer.conv.att_r.grad.flatten().cpu())
print("att_r orac", o.params["conv.att_r"].grad.flatten())
print("dx err", H.rel_err(xx.grad, xo.grad), "out err", H.rel_err(out[0], oo))
# where is argmax
batch_idx, subset, adj = batch_A
print("R", subset.numel(), "nnz", adj.nnz())
# torch restatement on the GPU from the bank's state
import torch.nn.functional as F
bank = layer.bank
tail = subset[B:].to(dev)
codes = bank.codes[tail].long()           # [B', nb]
xf = torch.cat([bank.O[k, codes[:, k], :D] for k in range(bank.nb)], 1)
xin = torch.cat([x.to(dev), xf], 0)
xin = torch.cat([xin, torch.ones(xin.shape[0], 1, device=dev)], 1)
attl = layer.conv.att_l.detach().clone().requires_grad_(True)
attr = layer.conv.att_r.detach().clone().requires_grad_(True)
A = adj.to_dense().to(dev)
y = restate.gat_propagate(A, xin, attl.view(-1), attr.view(-1))
yB = y[:B, :-1] / (y[:B, -1:] + 1e-16)
(yB_out := layer.gnn_transform(yB) + layer.linear_skip(x.to(dev)))
(yB_out * w).sum().backward()
print("att_l torch-on-bank", attl.grad.flatten().cpu())
# compare codes / O of the bank with the oracle's
for k in range(bank.nb):
    print(k, "codes equal", torch.equal(bank.codes[:, k].cpu(), o.c_indices[k]),
          "O err", H.rel_err(bank.O[k, :, :8], o.vq[k]._embedding_output))
