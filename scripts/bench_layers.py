#!/usr/bin/env python
"""Per-layer fwd + bwd (+ VQ update) time of one VQ-GNN layer at every BASELINE.json config shape on ONE GPU, with the
per-kernel times of the C-ABI launches and their achieved algorithmic GB/s (the headline bench.py covers configs[1]
end to end; this is the supporting table for the other shapes).  Synthetic seeded graphs (vq_gnn_b200/synth.py).

    python scripts/bench_layers.py > profiles/rNN_layer_bench.json
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import vq_gnn_b200 as V  # noqa: E402
from tests import helpers as H  # noqa: E402
from vq_gnn_b200 import _lib, sampling, synth  # noqa: E402

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def make_case(name, dev):
    s = synth.CONFIG_SHAPES[name]
    g = synth.make_graph(s["N"], s["E"], s["conv"], s["version"], seed=0, power_law=s["power_law"],
                         num_blocks=s["num_blocks"], device=dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    if name == "c1_arxiv":
        parts = torch.randperm(80, generator=torch.Generator().manual_seed(1))[:40]
        nodes = sampling.cluster_batch(s["N"], 80, parts).to(dev)
        C = 128
    elif name == "c2_reddit":
        seeds = torch.randperm(s["N"], generator=gen, device=dev)[:6000]
        nodes = sampling.cont_sampler(g, seeds, 3, 6000, generator=gen)[2]
        C = 128
    elif name == "c3_ppi":
        nodes = torch.randperm(s["N"], generator=gen, device=dev)[:10000]
        C = 256
    elif name == "c4_collab":
        seeds = torch.randperm(s["N"], generator=gen, device=dev)[:50000]
        nodes = sampling.cont_sampler(g, seeds, 15, 50000, generator=gen)[8]
        C = 128
    else:
        lo, hi = V.dist.partition_range(s["N"], 0, 8)
        nodes = lo + torch.randperm(hi - lo, generator=gen, device=dev)[:20000]
        C = 128
    bA = (sampling.k_hop_batch_v2(g, nodes, True) if s["version"] == "v2"
          else sampling.collate_batch_v1(g, nodes, True, True))
    return s, nodes, bA, C


def algorithmic_bytes(kernel, plan, bank, B, C):
    """Per-kernel compulsory bytes, the same function bench.py uses for its roofline -- one
    formula per kernel (round 1 applied the full-forward formula to the in-batch-only launch and reported frac > 1)."""
    import bench
    work, bound = bench.algorithmic_work(kernel, plan, C, bank.nb, bank.M, B)
    return work if (bound == "hbm" and work) else None


def main():
    dev = torch.device("cuda:0")
    out = {"hbm_peak_gbs": HBM, "note": "one layer (C -> C), fwd + bwd + VQ update, eager launches, CUDA events; "
                                        "3 warm-up + 10 timed iterations with a 256 MiB L2 flush between them",
           "cases": {}}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _lib.load()
    cases = os.environ.get("VQGNN_CASES", "c1_arxiv,c2_reddit,c3_ppi,c4_collab,c5_products").split(",")
    for name in cases:
        t0 = time.time()
        s, nodes, bA, C = make_case(name, dev)
        B = int(nodes.numel())
        torch.manual_seed(0)
        layer = V.LowRankGNNLayer(*H.layer_args(C, C, s["M"], 4, s["N"], s["conv"], skip=False),
                                  version=s["version"]).to(dev).train()
        if os.environ.get("VQGNN_MATERIALIZE"):
            layer.materialize_tail = {"force": "force", "0": False}.get(os.environ["VQGNN_MATERIALIZE"], True)
        impl = os.environ.get("VQGNN_ASSIGN_IMPL", "auto")
        layer.bank.assign_impl = impl if impl == "auto" else int(impl)
        plan = V.build_plan(bA, s["conv"], s["N"], True, dev).warm()
        x = torch.randn(B, C, device=dev)
        w = torch.randn(B, C, device=dev)

        def step():
            xx = x.clone().requires_grad_(True)
            o = layer(xx, plan, 1.0, False)
            ((o[0] * w).sum() + o[5]).backward()

        step()
        layer.set_inited(True)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        _lib.PROFILER.enabled = True
        _lib.PROFILER.reset()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot, n_it = 0.0, 10
        for _ in range(n_it):
            lib.vqgnn_flush_l2(_lib.ptr(flush), flush.numel(), _lib.stream())
            a.record()
            step()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        summ = _lib.PROFILER.summary()
        _lib.PROFILER.enabled = False
        summ.pop("vqgnn_flush_l2", None)
        layer.check_status()
        kern = {}
        for k, (n, ms) in sorted(summ.items(), key=lambda kv: -kv[1][1]):
            e = {"launches": n / n_it, "ms": ms / n_it}
            ab = algorithmic_bytes(k, plan, layer.bank, B, C)
            if ab:
                e["algorithmic_GBps"] = ab / (ms / n) / 1e6
                e["frac_of_hbm_peak"] = e["algorithmic_GBps"] / HBM
            kern[k] = e
        ms = tot / n_it
        out["cases"][name] = {"conv": s["conv"], "version": s["version"], "B": B, "R": plan.R, "nnz": plan.nnz, "C": C,
                              "M": s["M"], "ms_per_layer_fwd_bwd": ms, "nodes_per_s_per_layer": B / (ms * 1e-3),
                              "kernels": kern, "setup_s": round(time.time() - t0, 1)}
        print(f"[layers] {name}: B={B} nnz={plan.nnz} {ms:.3f} ms -> {B / (ms * 1e-3) / 1e6:.2f} M nodes/s/layer",
              file=sys.stderr, flush=True)
        del layer, plan, bA, x, w
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
