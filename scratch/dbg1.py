import torch, sys
sys.path.insert(0, '.')
import vq_gnn_b200 as V
from oracle import restate
from tests import helpers as H
dev = torch.device('cuda:0')
torch.manual_seed(0)
nb, M, D, B = 2, 16, 4, 120
bank = V.VQBank(nb, M, D, warm_up_flag=True, num_N=400)
orc = []
for i in range(nb):
    torch.manual_seed(i)
    o = restate.OracleVQ(M, D, warm_up_flag=True)
    bank.E[i,:,:8] = o._embedding; bank.Wm[i,:,:8] = o._ema_w
    orc.append(o)
bank.to(dev)
g = torch.Generator().manual_seed(5)
bidx = torch.randperm(400, generator=g)[:B].to(torch.int32).to(dev)
for step in range(4):
    X = torch.randn(B, nb*D, generator=g)*2+0.3
    G = torch.randn(B, nb*D, generator=g)*0.5
    if step == 0:
        idx = bank.run(X.to(dev), None, bidx, True)
        io = [o.feature_update(X[:, i*D:(i+1)*D]) for i, o in enumerate(orc)]
    else:
        idx = bank.run(X.to(dev), G.to(dev), bidx, True)
        io = [o.update(X[:, i*D:(i+1)*D], G[:, i*D:(i+1)*D])[0] for i, o in enumerate(orc)]
    for i, o in enumerate(orc):
        mism = (idx[:, i].cpu().long() != io[i].view(-1)).sum().item()
        print(step, i, 'mism', mism,
              'rm_f', H.rel_err(bank.rm_f[i], o.feat_mean), 'rv_f', H.rel_err(bank.rv_f[i], o.feat_var),
              'rm_g', H.rel_err(bank.rm_g[i], o.grad_mean), 'rv_g', H.rel_err(bank.rv_g[i], o.grad_var),
              'size', H.rel_err(bank.size[i], o._ema_cluster_size), 'Wm', H.rel_err(bank.Wm[i,:,:8], o._ema_w),
              'E', H.rel_err(bank.E[i,:,:8], o._embedding), 'O', H.rel_err(bank.O[i,:,:8], o._embedding_output))
