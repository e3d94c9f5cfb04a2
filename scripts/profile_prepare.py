"""Device time of the batch-graph construction (LowRankGNN.prepare_from_graph) per C-ABI call, c5 shapes.
   python scripts/profile_prepare.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from vq_gnn_b200 import _lib

dev = torch.device("cuda:0")
c = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c5"]
wl = bench.Workload(c, dev, 0, 1)
model, head = bench.build_model(c, dev, wl.N, False)
for n in wl.node_lists:
    model.prepare_from_graph(wl.g, n)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for n in wl.node_lists:
    model.prepare_from_graph(wl.g, n)
b.record()
torch.cuda.synchronize()
print(f"prepare_from_graph: {a.elapsed_time(b) / len(wl.node_lists):.3f} ms per batch (device time incl. host gaps)")
_lib.PROFILER.enabled = True
_lib.PROFILER.reset()
for n in wl.node_lists:
    model.prepare_from_graph(wl.g, n)
torch.cuda.synchronize()
for k, (cnt, ms) in sorted(_lib.PROFILER.summary().items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:36s} {cnt / len(wl.node_lists):5.1f} x {ms / len(wl.node_lists):8.3f} ms per batch")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for n in wl.node_lists:
        model.prepare_from_graph(wl.g, n)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
