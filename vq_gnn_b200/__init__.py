"""vq_gnn_b200 — B200-native (sm_100a) implementation of VQ-GNN's hot path.

Module and class names mirror the reference (devnkong/VQ-GNN, `vq.py`, `convs.py`, `models.py`) so the
layers are a drop-in; the arithmetic runs in hand-written CUDA kernels behind the C-ABI of
`libvqgnn.so` (include/vqgnn.h).  There is no CPU fallback: entry points raise if the library is
missing or the device is not sm_100.
"""
from . import _lib, convs, dist, graph, link, models, sampling, synth, vq  # noqa: F401
from .convs import OurGATConv, OurGCNConv  # noqa: F401
from .graph import BatchPlan, CSRAdj, build_plan  # noqa: F401
from .link import LinkPredictor, link_loss  # noqa: F401
from .models import LowRankGNN, LowRankGNNBlock, LowRankGNNLayer  # noqa: F401
from .vq import VectorQuantizerEMA, VQBank  # noqa: F401

__all__ = ["VectorQuantizerEMA", "VQBank", "OurGCNConv", "OurGATConv", "LowRankGNNBlock", "LowRankGNNLayer",
           "LowRankGNN", "CSRAdj", "BatchPlan", "build_plan", "LinkPredictor", "link_loss"]
