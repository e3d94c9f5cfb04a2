"""How well do the memory-bound row-gather forward and the tensor/TMEM-bound VQ update overlap on two streams?
(c5 shapes; decides whether moving the x-independent part of info_backward off the critical path can pay.)
   python scripts/overlap_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import vq_gnn_b200 as V
from vq_gnn_b200.models import VQConvFunction

dev = torch.device("cuda:0")
c = bench.CONFIGS["c5"] if hasattr(bench, "CONFIGS") else None
if c is None:
    c = [v for k, v in vars(bench).items() if isinstance(v, dict) and v.get("name", "").startswith("c5")][0]
wl = bench.Workload(c, dev, 0, 1)
model, head = bench.build_model(c, dev, wl.N, False)
model.warm_start(wl.g, wl.X, 60000)
plan = model.prepare_from_graph(wl.g, wl.node_lists[0])
layer = model.convs[1]
B = plan.B
x = torch.randn(B, 128, device=dev)
g = torch.randn(B, 128, device=dev) * 1e-3
layer.set_inited(True)
layer.train()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def fwd():
    with torch.no_grad():
        VQConvFunction.apply(x, None, layer, plan, 1.0, False)


def vq():
    layer.bank.run(x, g, plan.batch_idx, True)


def timed(fn_a, fn_b, reps=10):
    for _ in range(3):
        if fn_a:
            with torch.cuda.stream(s1):
                fn_a()
        if fn_b:
            with torch.cuda.stream(s2):
                fn_b()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if fn_a:
            with torch.cuda.stream(s1):
                fn_a()
        if fn_b:
            with torch.cuda.stream(s2):
                fn_b()
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


t_f, t_v, t_fv = timed(fwd, None), timed(None, vq), timed(fwd, vq)
print(f"forward (materialise + row gather) alone {t_f:.3f} ms, VQ update alone {t_v:.3f} ms, both concurrently {t_fv:.3f} ms "
      f"(sum {t_f + t_v:.3f}, max {max(t_f, t_v):.3f})")
