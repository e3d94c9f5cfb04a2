import torch, sys
sys.path.insert(0, '.')
import torch.nn.functional as F
import vq_gnn_b200 as V
from oracle import restate
from tests import helpers as H
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
        lsd = {k[len(f"convs.{li}."):]: v.clone() for k, v in model.state_dict().items() if k.startswith(f"convs.{li}.")}
        oracles.append(restate.OracleLayer(ci, co, M, D, N, conv, version, warm_up_flag=True).load_state_dict(lsd))
    model = model.to(dev).train()
    x = torch.randn(B, 12, generator=torch.Generator().manual_seed(3))
    y = torch.randint(0, 7, (B,), generator=torch.Generator().manual_seed(4))
    bA = H.batch_to(batch_A, dev)
    for step in range(4):
        if step == 1:
            model.set_inited(True)
            for o in oracles: o.set_inited(True)
        model.zero_grad()
        out, _, info = model((x.to(dev), bA), 1)
        loss = F.cross_entropy(out, y.to(dev)) + info
        loss.backward()
        h = x.clone(); info_o = 0
        for li, o in enumerate(oracles):
            for p in o.params.values(): p.grad = None
            h, inf = o(h, batch_A, 1.0, False)
            info_o = info_o + inf
            if li < 2:
                h = F.batch_norm(h, None, None, training=True)
                h = restate.act_leaky_gelu(h)
        loss_o = F.cross_entropy(h, y) + info_o
        loss_o.backward()
        print(version, step, 'out', H.rel_err(out, h), 'loss', float(loss), float(loss_o), 'info', float(info), float(info_o))
        for li, o in enumerate(oracles):
            gw = model.convs[li].gnn_transform.weight.grad.cpu()
            sd = {k: v.cpu() for k, v in model.convs[li].state_dict().items()}
            nm = sum(int((sd[k] != v).sum()) for k, v in o.state_dict().items() if not v.is_floating_point())
            worst = max((H.rel_err(sd[k], v), k) for k, v in o.state_dict().items() if v.is_floating_point())
            print('   layer', li, 'gw', H.rel_err(gw, o.params["gnn_transform.weight"].grad), 'code mism', nm, 'worst state', worst)
