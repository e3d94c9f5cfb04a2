#!/usr/bin/env python
"""bench.py -- train nodes/sec per VQ-GNN layer on B200 (BASELINE.json metric).

Default workload (--config c5): BASELINE.json configs[4], the config the metric's "1/2/4/8 B200" sweep is quoted
on: VQ-GNN GCN (v2 "B+B'" formulation) on a synthetic ogbn-products-shaped graph (2,449,029 nodes, 61.86 M
undirected edges, 100-d, 47 classes), hidden 128, num-M 4096, num-D 4, 3 layers; every rank owns a contiguous node
partition and samples B = 20,000 batch nodes per step from it (node sampler), replicas of the codebooks and code
tables are kept identical by allreducing the whitening moments + per-codeword EMA statistics and all-gathering the
code-table updates; dense weight gradients are allreduced once per step.  --config c1 | c2 | c3 | c4 select the
other BASELINE.json configs (arxiv GCN / Reddit v1 SAGE / PPI GAT / collab link prediction).

One "step" = one pass of the hot path over one mini-batch: L-layer LowRankGNN forward + loss + backward (the VQ
assignment + EMA update of every layer fire inside backward) + RMSprop step.
value = (batch nodes of all ranks) * num_layers / t_step   [batch nodes / second / layer].

Keys: see the task contract.  `value` is device-resident (plans prebuilt, CUDA events, whole-step CUDA graphs);
`e2e` goes through the public API from HOST node ids: H2D of the ids, device-side batch construction
(LowRankGNN.prepare_from_graph) + feature/label gather, train step, D2H of the loss, all inside the timed region;
`roofline` is for the dominant kernel of the step found live with CUDA events around every C-ABI launch;
`cpu_baseline` / `--impl reference` time the UNMODIFIED reference (oracle/_ref, through oracle/shims) -- or the
oracle port when that directory is absent -- on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json configs[i] -> workload; model flags from the reference README commands / parser defaults
CONFIGS = {
    "c1": dict(idx=0, name="c1_arxiv_gcn_v2", shape="c1_arxiv", feat=128, C_in=128, hidden=128, classes=40, M=256, D=4,
               layers=3, conv="GCN", version="v2", skip=False, sampler="cluster", parts=80, batch_parts=40, loss="ce",
               lr=1e-3, what="VQ-GNN GCN, synthetic ogbn-arxiv-shaped graph (169,343 nodes, 1.17M edges, 128-d, 40 "
                            "classes), num-M 256, num-D 4, cluster sampler 80 parts, batch 40 parts"),
    "c2": dict(idx=1, name="c2_reddit_sage_v1", shape="c2_reddit", feat=602, C_in=604, hidden=128, classes=41, M=1024,
               D=4, layers=3, conv="SAGE", version="v1", skip=False, sampler="cont", B=6000, walk=3, loss="ce", lr=1e-3,
               what="VQ-GNN SAGE-Mean (v1 formulation), synthetic Reddit-shaped graph (232,965 nodes, ~114.6M directed "
                    "edges, 602-d padded to 604, 41 classes), hidden 128, num-M 1024, num-D 4, batch 6000, cont sampler "
                    "walk 3"),
    "c3": dict(idx=2, name="c3_ppi_gat_v2", shape="c3_ppi", feat=50, C_in=52, hidden=256, classes=121, M=4096, D=4,
               layers=3, conv="GAT", version="v2", skip=True, sampler="node", B=10000, loss="bce", lr=3e-3,
               what="VQ-GNN GAT, synthetic PPI-shaped graph (44,906 train nodes, 50-d padded to 52, 121 multi-labels), "
                    "hidden 256, num-M 4096, num-D 4, batch 10000, node sampler, --skip"),
    "c4": dict(idx=3, name="c4_collab_gcn_link_v2", shape="c4_collab", feat=128, C_in=128, hidden=128, classes=128,
               M=1024, D=4, layers=3, conv="GCN", version="v2", skip=True, sampler="cont", B=50000, walk=15, loss="link",
               lr=3e-3,
               what="VQ-GNN GCN link prediction, synthetic ogbl-collab-shaped graph (235,868 nodes, 1.29M edges, 128-d), "
                    "num-M 1024, num-D 4, batch 50000, cont sampler walk 15, LinkPredictor head, --skip"),
    "c5": dict(idx=4, name="c5_products_gcn_v2", shape="c5_products", feat=100, C_in=100, hidden=128, classes=47, M=4096,
               D=4, layers=3, conv="GCN", version="v2", skip=False, sampler="node", B=20000, loss="ce", lr=1e-3,
               what="VQ-GNN GCN scale sweep, synthetic ogbn-products-shaped graph (2,449,029 nodes, 61.86M undirected "
                    "edges, 100-d, 47 classes), hidden 128, num-M 4096, num-D 4, node sampler inside the rank's own "
                    "node partition, batch 20000 per rank (weak) or 20000 in total (strong)"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [t.strip() for t in s.split(",")]
            try:
                sm.append(float(f[0])), (mx := float(f[1]))
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []   # median over the busier half of the samples (under load)
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# workload construction (synthetic, seeded, identical on every rank; generated with torch ops -- not timed)
# ------------------------------------------------------------------------------------------------
class Workload:
    """ONE synthetic graph (the same on every rank: same Philox seed, fp64 CPU CDF for the power-law shapes), the
    resident feature / label matrices, and this rank's mini-batches (node ids drawn from the rank's own contiguous
    node range)."""

    def __init__(self, cfg, dev, rank: int, world: int, scale: float = 1.0, n_batches: int = 4, strong: bool = False):
        from vq_gnn_b200 import dist as vdist
        from vq_gnn_b200 import sampling, synth
        self.cfg, self.dev = cfg, dev
        s = synth.CONFIG_SHAPES[cfg["shape"]]
        N, E = int(s["N"] * scale), int(s["E"] * scale)
        t0 = time.time()
        rowptr, row, col = synth.random_edges(N, E, seed=0, power_law=s["power_law"], num_blocks=s["num_blocks"],
                                              device=dev)
        w = None
        if cfg["name"].startswith("c4"):    # collab keeps integer edge weights as values (vq_gnn_v2/main_link.py:273-274);
            lo_, hi_ = torch.minimum(row, col), torch.maximum(row, col)   # symmetric: a hash of the unordered pair
            w = ((lo_ * 2654435761 + hi_ * 40503) % 3 + 1).float()
        self.g = g = synth.normalized_graph(N, rowptr, row, col, cfg["conv"], cfg["version"], edge_weight=w)
        del row
        self.N = N
        gen = torch.Generator(device=dev).manual_seed(7)
        self.X = torch.randn(N, cfg["C_in"], generator=gen, device=dev)
        if cfg["C_in"] > cfg["feat"]:
            self.X[:, cfg["feat"]:] = 0      # zero padding to a multiple of num_D (vq_gnn_v2/utils/misc.py:212-219)
        if cfg["loss"] == "bce":
            self.Y = (torch.rand(N, cfg["classes"], generator=gen, device=dev) < 0.3).float()
        else:
            self.Y = torch.randint(0, cfg["classes"], (N,), generator=gen, device=dev)
        # ---- this rank's batches
        genr = torch.Generator(device=dev).manual_seed(100 + rank)
        lo, hi = vdist.partition_range(N, rank, world)
        self.B = B = (cfg.get("B", 0) // world if strong else cfg.get("B", 0)) or None
        node_lists = []
        if cfg["sampler"] == "node":
            for _ in range(n_batches):
                node_lists.append(lo + torch.randperm(hi - lo, generator=genr, device=dev)[:B])
        elif cfg["sampler"] == "cont":
            seeds = lo + torch.randperm(hi - lo, generator=genr, device=dev)[:B]
            walks = sampling.cont_sampler(g, seeds, cfg["walk"], B, generator=genr)
            node_lists = walks[:n_batches] if cfg["walk"] < 8 else walks[1:1 + n_batches]
        else:   # cluster: contiguous id blocks stand in for METIS parts
            gcpu = torch.Generator().manual_seed(100 + rank)
            for _ in range(n_batches):
                parts = torch.randperm(cfg["parts"], generator=gcpu)[:cfg["batch_parts"]]
                node_lists.append(sampling.cluster_batch(N, cfg["parts"], parts).to(dev))
        self.node_lists = [n.contiguous() for n in node_lists]
        # (features, None, labels) per batch: the batch GRAPHS are built by the product itself
        # (LowRankGNN.prepare_from_graph, csrc/khop.cu); make_batch() restates one in the reference's host format
        # for the CPU arm only
        self.batches = [(self.X[n], None, self.Y[n]) for n in self.node_lists]
        if dev.type == "cuda":
            torch.cuda.synchronize()
        self.checksum = [int(g.nnz), int(g.col.sum()), int(torch.round(g.val.double().sum() * 1e3))]
        log(f"[bench] rank {rank}: graph N={N} nnz={g.nnz} built in {time.time() - t0:.1f}s; "
            f"batches B={[int(n.numel()) for n in self.node_lists]}")

    def make_batch(self, nodes):
        from vq_gnn_b200 import sampling
        c = self.cfg
        if c["version"] == "v2":
            bA = sampling.k_hop_batch_v2(self.g, nodes, True)
        else:
            bA = sampling.collate_batch_v1(self.g, nodes, True, True)
        return self.X[nodes], bA, self.Y[nodes]


def build_model(cfg, dev, N, distributed: bool, assign_impl='auto', capacity=None):
    import vq_gnn_b200 as V
    c = cfg
    torch.manual_seed(0)
    model = V.LowRankGNN(c["C_in"], c["hidden"], c["classes"], c["layers"], 0.0, c["M"], c["D"], N,
                         no_second_fc=True, skip=c["skip"], commitment_cost=0.0, grad_scale=[1, 1], act="leaky_gelu",
                         bn_flag=True, warm_up_flag=True, momentum=0.1, conv_type=c["conv"], version=c["version"])
    model = model.to(dev).train()
    for layer in model.convs:
        layer.bank.assign_impl = assign_impl if assign_impl == 'auto' else int(assign_impl)
        layer.bank.distributed = distributed
        layer.bank.gather_capacity = capacity
    head = None
    if c["loss"] == "link":
        head = V.LinkPredictor(c["classes"], c["hidden"], 1, 3, 0.0).to(dev).train()
    return model, head


def loss_fn(cfg, out, y, head=None, plan=None):
    if cfg["loss"] == "ce":
        return F.cross_entropy(out, y)
    if cfg["loss"] == "bce":
        return F.binary_cross_entropy_with_logits(out, y)
    from vq_gnn_b200 import link
    return link.link_loss(head, out, plan)


def train_step(cfg, model, head, opt, x, plan, y, distributed: bool):
    opt.zero_grad(set_to_none=True)
    out, _, info = model((x, plan), 1)
    loss = loss_fn(cfg, out, y, head, plan) + info
    loss.backward()
    if distributed:
        from vq_gnn_b200 import dist as vdist
        params = list(model.parameters()) + (list(head.parameters()) if head is not None else [])
        vdist.allreduce_mean_grads_(params)
    opt.step()
    model.join_vq_updates()      # side-stream VQ updates (if enabled) re-join here: the step is self-contained
    return loss


# ------------------------------------------------------------------------------------------------
# algorithmic work per launch (compulsory HBM bytes / useful FLOPs), for the roofline
# ------------------------------------------------------------------------------------------------
def algorithmic_work(kernel: str, plan, C: int, nb: int, M: int, B: int):
    """-> (work, 'hbm' | 'tensor') for ONE launch of `kernel` on this plan at layer width C (fp32 = 4 B, code = 2 B,
    index = 4 B; §8d's table).  Unknown kernels return (0, 'hbm')."""
    v1 = plan.version == 'v1'
    nnz = plan.nnz
    if kernel == "vqgnn_mp_info":   # out-of-batch rows of the v2 forward: edge list + row pointers + codes + codebooks + x
        nnz_o = nnz - plan.nnz_B
        return nnz_o * 8 + (plan.T + 1) * 4 + plan.T * nb * 2 + 2 * M * C * 4 + B * C * 4, "hbm"
    if kernel == "vqgnn_tail_materialize_slab":
        return plan.T * nb * 2 + 2 * plan.T * C * 4 + 2 * M * C * 4, "hbm"
    if kernel == "vqgnn_mp_fwd_rows":   # same compulsory bytes as the generic forward / backward it replaces
        kernel = "vqgnn_mp_fwd"
    if kernel == "vqgnn_mp_fwd_rows:bwd":
        kernel = "vqgnn_mp_bwd"
    if kernel in ("vqgnn_mp_fwd", "vqgnn_gat_fwd"):
        if not v1 and plan.extras.get('split_fwd') and kernel == "vqgnn_mp_fwd":   # batch rows only (split forward)
            nB = plan.nnz_B
            return nB * 8 + (B + 1) * 4 + 2 * B * C * 4 + nB * nb * 2 + M * C * 4, "hbm"
        if v1 and 'split' in plan.extras:      # in-batch block only (the tail goes through vqgnn_mp_fwd_tail)
            nin = int(plan.extras['split']['inb'][4])
            return nin * 8 + (B + 1) * 4 + 2 * B * C * 4, "hbm"
        T, R = plan.T, plan.R
        by = nnz * 8 + (R + 1) * 4 + B * C * 4 + (T if not v1 else nnz) * nb * 2 + 2 * M * C * 4 + B * C * 4
        if kernel == "vqgnn_gat_fwd":
            by += nnz * 0 + 2 * R * 4
        return by, "hbm"
    if kernel == "vqgnn_mp_fwd_tail":
        tn = int(plan.extras['split']['tail'][5]) if 'split' in plan.extras else nnz
        if 'split' in plan.extras and len(plan.extras['split']['tail']) > 6:
            tn = int(plan.extras['split']['tail'][6].item())
        # §8d "mp_fwd v1 form": nnz(A_BN)*(8 + nb*2) + B*C*4*2 + M*C*4 (+ 4 B reverse value per entry, gq write)
        return tn * (8 + 4 + nb * 2) + (B + 1) * 4 + 3 * B * C * 4 + 2 * M * C * 4, "hbm"
    if kernel in ("vqgnn_mp_bwd", "vqgnn_gat_bwd"):
        nnz_t = int(plan.bwd_col.numel())
        by = nnz_t * 8 + (B + 1) * 4 + 2 * B * C * 4 + M * C * 4
        if kernel == "vqgnn_gat_bwd":
            by += nnz * 8 + 4 * plan.R * 4
        return by, "hbm"
    if kernel == "vqgnn_tail_materialize":
        return plan.T * nb * 2 + 2 * plan.T * C * 4 + 2 * M * C * 4, "hbm"
    if kernel == "vqgnn_vq_assign":   # joint: FLOPs = 2*B*M*2C (useful)
        return 4.0 * B * M * C, "tensor"
    if kernel == "vqgnn_vq_segsum":
        return 2 * B * C * 4 + B * nb * 2 + nb * M * 12 * 4, "hbm"
    if kernel == "vqgnn_vq_moments":
        return 2 * B * C * 4, "hbm"
    if kernel == "vqgnn_vq_finalize":
        return nb * M * (12 * 4 + 8 * 4 * 4 + 4 * 2), "hbm"
    return 0, "hbm"


# ------------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (oracle/_ref or /root/reference, via oracle/shims) or, failing that, the port
# ------------------------------------------------------------------------------------------------
def _ref_root():
    for p in (os.path.join(ROOT, "oracle", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(p, "vq_gnn_v2", "vq.py")):
            return p
    return None


def batch_to_cpu(batch):
    x, bA, y = batch
    def mv(t):
        if t is None:
            return None
        if isinstance(t, tuple):
            return tuple(u.cpu() for u in t)
        return t.to("cpu") if hasattr(t, "to") else t
    return x.cpu(), tuple(mv(t) for t in bA), y.cpu()


def cpu_reference_run(cfg, batch_cpu, N, steps: int, warmup: int, budget_s: float = 150.0):
    """One full train step of the reference model (all layers, all branches): forward + loss + backward + RMSprop
    on the host cores.  Returns (nodes/s/layer, seconds per step (median), steps timed, kind, description)."""
    torch.set_num_threads(os.cpu_count())
    c = cfg
    x, bA, y = batch_cpu
    B = x.shape[0]
    root = _ref_root()
    if root is not None and c["loss"] != "link":
        os.environ["VQGNN_REFERENCE_ROOT"] = root
        from oracle import ref_loader
        from tests import helpers as H
        ref = ref_loader.load_reference(c["version"])
        ts = ref_loader.shim_sparse()
        torch.manual_seed(0)
        model = ref.models.LowRankGNN(c["C_in"], c["hidden"], c["classes"], c["layers"], 0.0, c["M"], c["D"], N,
                                      no_second_fc=True, skip=c["skip"], commitment_cost=0.0, grad_scale=[1, 1],
                                      act="leaky_gelu", bn_flag=True, warm_up_flag=True, momentum=0.1,
                                      conv_type=c["conv"]).train()
        bA_ref = H.to_shim_batch(bA, ts)
        opt = torch.optim.RMSprop(model.parameters(), lr=c["lr"], alpha=0.99)
        kind = "reference"
        what = (f"UNMODIFIED reference ({os.path.relpath(root, ROOT) if root.startswith(ROOT) else root}: "
                f"vq_gnn_{c['version']}/models.py LowRankGNN through oracle/shims for the absent torch_sparse / PyG "
                f"leaves)")

        def set_inited():
            for conv in model.convs:
                for blk in conv.gnn_block:
                    blk.inited = True

        def step():
            opt.zero_grad()
            out, _, info = model((x, bA_ref), 1)
            loss = loss_fn(c, out, y) + info
            loss.backward()
            opt.step()
            return float(loss)
    else:
        from oracle import restate
        torch.manual_seed(0)
        dims = [(c["C_in"], c["hidden"])] + [(c["hidden"], c["hidden"])] * (c["layers"] - 2) + [(c["hidden"], c["classes"])]
        layers = []
        for ci, co in dims:
            L = restate.OracleLayer(ci, co, c["M"], c["D"], N, c["conv"], c["version"], skip=c["skip"],
                                    warm_up_flag=True, sparse=True)
            L.params = {"gnn_transform.weight": (torch.randn(co, ci) * 0.05).requires_grad_(True),
                        "gnn_transform.bias": torch.zeros(co, requires_grad=True)}
            for nm in (["fc_sage"] if c["conv"] == "SAGE" else []) + (["linear_skip"] if c["skip"] else []):
                L.params[nm + ".weight"] = (torch.randn(co, ci) * 0.05).requires_grad_(True)
                L.params[nm + ".bias"] = torch.zeros(co, requires_grad=True)
            layers.append(L.train())
        params = [p for L in layers for p in L.params.values()]
        opt = torch.optim.RMSprop(params, lr=c["lr"], alpha=0.99)
        kind = "port"
        what = "oracle port (oracle/restate.py; the reference sources are not on this box)"

        def set_inited():
            for L in layers:
                L.set_inited(True)

        def step():
            opt.zero_grad()
            h, info = x, 0
            for li, L in enumerate(layers):
                h, inf = L(h, bA, 1.0, False)
                info = info + inf
                if li < len(layers) - 1:
                    h = restate.act_leaky_gelu(F.batch_norm(h, None, None, training=True))
            loss = loss_fn(c, h, y) + info
            loss.backward()
            opt.step()
            return float(loss)

    t_start = time.perf_counter()
    step()                       # un-inited pass: the feature-only warm start of every layer (reference init())
    set_inited()
    for _ in range(max(warmup - 1, 0)):
        if time.perf_counter() - t_start > budget_s / 2:
            break
        step()
    times = []
    for _ in range(max(steps, 1)):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    t_med = statistics.median(times)
    return B * c["layers"] / t_med, t_med, len(times), kind, what


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=str, default=os.environ.get("VQGNN_BENCH_CONFIG", "c5"), choices=sorted(CONFIGS))
    ap.add_argument("--scaling", type=str, default="weak", choices=["weak", "strong"],
                    help="weak: per-GPU batch fixed (default); strong: the config's batch split over the ranks")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the synthetic graph (debug only)")
    ap.add_argument("--assign-impl", type=str, default=os.environ.get("VQGNN_ASSIGN_IMPL", "auto"),
                    help="auto (default: tcgen05/TMEM kernel when M >= 512), 1 = always tcgen05, 0 = exact-fp32 SIMT")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true",
                    help="launch the device-resident steps eagerly instead of replaying one CUDA graph per batch")
    ap.add_argument("--sync-vq", action="store_true",
                    help="run the VQ hook updates on the compute stream (default: side stream, off the critical path)")
    ap.add_argument("--cpu-batch", type=int, default=0, help="batch nodes of the CPU arm's sample (0 = the config's B)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    c = CONFIGS[args.config]
    strong = args.scaling == "strong"
    config = {"workload": f"configs[{c['idx']}]: {c['what']}; {c['layers']} layers "
                          f"({c['C_in']}->{c['hidden']}->...->{c['classes']})",
              "step": f"{c['layers']}-layer fwd + loss + bwd (VQ assign + EMA update of every layer inside bwd) + RMSprop",
              "value_formula": "sum over ranks of batch nodes * num_layers / t_step", "num_layers": c["layers"],
              "parallelism": f"dp{world} (one graph, node-partitioned batches; whitening moments + EMA stats + weight "
                             "grads allreduced, code-table updates all-gathered)",
              "scaling_mode": args.scaling,
              "l2": "4 distinct batches rotated AND a 256 MiB L2 flush between timed steps",
              "launch": "one CUDA graph per resident batch (whole train step incl. NCCL) replayed; --no-graphs = eager"}

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
        if args.cpu_batch:
            c = dict(c, B=args.cpu_batch)
        wl = Workload(c, dev, 0, 1, args.scale, n_batches=1)
        batch_cpu = batch_to_cpu(wl.make_batch(wl.node_lists[0]))
        del wl
        if dev.type == "cuda":
            torch.cuda.empty_cache()
        v, t_step, n_timed, kind, what = cpu_reference_run(c, batch_cpu, _num_nodes(c, args.scale), args.steps,
                                                           min(args.warmup, 2))
        B = batch_cpu[0].shape[0]
        sample = (f"{what}: full {c['layers']}-layer train step (fwd + loss + bwd + RMSprop, every branch of every "
                  f"layer) on one batch of the same workload (B={B}); median of {n_timed} timed steps after the "
                  f"warm-start pass; {t_step * 1e3:.0f} ms per step")
        line = {"impl": "reference", "metric": "train nodes/sec per VQ-GNN layer", "value": v,
                "unit": "nodes/s/layer", "n_gpus": args.gpus, "steps": n_timed, "warmup": min(args.warmup, 2),
                "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "nodes/s/layer", "cores": os.cpu_count(), "kind": kind,
                                 "sample": sample},
                "e2e": {"value": v, "unit": "nodes/s/layer", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ our arm
    from vq_gnn_b200 import _lib
    from vq_gnn_b200 import dist as vdist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        torch.distributed.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    _lib.require_device(torch.zeros(1, device=dev))
    pk = peaks()

    wl = Workload(c, dev, rank, world, args.scale, strong=strong)
    g, batches, N = wl.g, wl.batches, wl.N
    graph_identical = True
    if distributed:       # ONE graph: every rank must have built the same one
        cs = torch.tensor(wl.checksum, dtype=torch.int64, device=dev)
        allcs = [torch.empty_like(cs) for _ in range(world)]
        torch.distributed.all_gather(allcs, cs)
        graph_identical = all(torch.equal(a, allcs[0]) for a in allcs)
        assert graph_identical, f"ranks built different graphs: {[a.tolist() for a in allcs]}"
    Bmax = max(int(b[0].shape[0]) for b in batches)
    capacity = None
    if distributed:
        cap = torch.tensor([Bmax], dtype=torch.int64, device=dev)
        torch.distributed.all_reduce(cap, op=torch.distributed.ReduceOp.MAX)
        capacity = int(cap.item())
    model, head = build_model(c, dev, N, distributed, args.assign_impl, capacity)
    per_layer = os.environ.get("VQGNN_VQ_STREAMS", "one") == "per_layer"   # measured at c5: 5.97 vs 6.00 ms, not worth 3 communicators
    model.set_async_vq_updates(not args.sync_vq, per_layer_streams=per_layer)
    if distributed and not args.sync_vq:
        # the side-stream VQ collectives get their OWN communicator: ProcessGroupNCCL funnels every collective of a
        # group through one in-order NCCL stream, so sharing the default group would make the weight-gradient
        # allreduce on the compute stream wait for all VQ updates issued before it
        # (one per layer when the layers' updates run on their own side streams: NCCL calls on one communicator must
        # not run concurrently)
        vq_group = None
        for layer in model.convs:
            if vq_group is None or per_layer:
                vq_group = torch.distributed.new_group()
            layer.bank.process_group = vq_group
    use_graphs = not args.no_graphs
    params = list(model.parameters()) + (list(head.parameters()) if head is not None else [])
    opt = torch.optim.RMSprop(params, lr=c["lr"], alpha=0.99, capturable=use_graphs)
    opt_eager = torch.optim.RMSprop(params, lr=c["lr"], alpha=0.99) if use_graphs else opt
    # the reference's init(): streaming feature-only warm start over the WHOLE graph (L(L+1)/2 layer passes), on the device
    torch.cuda.synchronize()
    t_ws = time.perf_counter()
    ws_bs = min(N, 60000)
    ws_batches = model.warm_start(g, wl.X, ws_bs)
    model.check_status()
    torch.cuda.synchronize()
    warm_start_info = {"seconds": round(time.perf_counter() - t_ws, 3), "nodes": N, "batch_nodes": ws_bs,
                       "batches_per_pass": ws_batches, "layer_passes": c["layers"] * (c["layers"] + 1) // 2,
                       "what": "LowRankGNN.warm_start = main_node.py init(): untimed for the metric, reported"}
    log(f"[bench] rank {rank}: warm start {warm_start_info['seconds']} s")
    plans = [model.prepare_from_graph(g, n) for n in wl.node_lists]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _lib.load()

    def flush_l2():
        lib.vqgnn_flush_l2(_lib.ptr(flush), flush.numel(), _lib.stream())

    def sync_all():
        if distributed:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step_eager(i, optimizer):
        x, _, y = batches[i % len(batches)]
        return train_step(c, model, head, optimizer, x, plans[i % len(plans)], y, distributed)

    for i in range(args.warmup):
        step_eager(i, opt)
    sync_all()

    # ---- end-to-end through the public API from HOST node ids (eager; measured BEFORE the CUDA graphs of the
    #      device-resident timing exist: their private memory pools slow later eager allocation down) ----
    from vq_gnn_b200.loader import DevicePrefetcher
    host_ids = [n.cpu().pin_memory() for n in wl.node_lists]
    h2d = sum(t.numel() * t.element_size() for t in host_ids) / len(host_ids)

    def prep(ids):
        return wl.X[ids], model.prepare_from_graph(g, ids), wl.Y[ids]

    side = torch.cuda.Stream(device=dev)
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()     # double-buffered D2H target for the step results

    def e2e_run(n):
        """Every step: H2D of the batch's node ids, device-side batch construction + plan, train step, and a D2H
        copy of the step's loss.  The loss of step i is copied asynchronously into pinned memory and READ on the
        host one step later (after its event), so the host never idles the GPU waiting for the step it has just
        launched."""
        pf = DevicePrefetcher(host_ids, dev, prepare=prep, count=n, stream=side, threaded=True)
        pending, last = None, 0.0
        for i in range(n):
            x, plan, y = pf.next()
            flush_l2()
            loss = train_step(c, model, head, opt_eager, x, plan, y, distributed)
            loss_host[i & 1].copy_(loss.detach(), non_blocking=True)           # D2H of this step's result
            ev_l = torch.cuda.Event()
            ev_l.record()
            if pending is not None:
                pending[0].synchronize()
                last = float(loss_host[pending[1]])                             # host read of the previous step's loss
            pending = (ev_l, i & 1)
        if pending is not None:
            pending[0].synchronize()
            last = float(loss_host[pending[1]])
        pf.drain()
        return last

    # untimed passes until the step time has settled: the first passes grow the caching allocators (device pools of the
    # side stream, pinned host blocks of the plans' count reads) and are several times slower than the steady state
    prev, warm_passes = None, []
    for _ in range(8):
        sync_all()
        t0 = time.perf_counter()
        e2e_run(max(args.steps, 2 * len(host_ids)))
        torch.cuda.synchronize()
        cur = (time.perf_counter() - t0) * 1e3
        warm_passes.append(round(cur, 2))
        settled = prev is not None and abs(cur - prev) <= 0.1 * prev
        prev = cur
        flag = torch.tensor([1.0 if settled else 0.0], device=dev)
        if distributed:      # every rank must run the same number of (collective-carrying) passes
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        if float(flag.item()) > 0:
            break
    sync_all()
    e2e_passes = []
    for _ in range(3):
        sync_all()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        torch.cuda.synchronize()
        e2e_passes.append((time.perf_counter() - t0) * 1e3)
    e2e_ms = statistics.median(e2e_passes)          # the MEDIAN pass is reported
    e_t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if distributed:
        torch.distributed.all_reduce(e_t, op=torch.distributed.ReduceOp.MAX)
    step_nodes = float(sum(batches[i % len(batches)][0].shape[0] for i in range(args.steps)))
    e2e_nodes = torch.tensor([step_nodes], dtype=torch.float64, device=dev)
    if distributed:
        torch.distributed.all_reduce(e2e_nodes)
    e2e_value = float(e2e_nodes.item()) * c["layers"] / (float(e_t.item()) / 1e3)

    # The step is launch-bound when run eagerly: capture one CUDA graph per resident batch (the plan's shapes differ
    # per batch) and replay it.  Every kernel of the step -- forward, backward, the VQ updates, the NCCL collectives
    # and the optimiser -- is inside the graph.
    graphs, graph_launches = None, []
    if use_graphs:
        try:
            graphs = []
            side_cap = torch.cuda.Stream(device=dev)
            for bi in range(len(batches)):
                side_cap.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side_cap):
                    step_eager(bi, opt)       # allocator warm-up on the side stream
                torch.cuda.current_stream().wait_stream(side_cap)
                gph = torch.cuda.CUDAGraph()
                lc0 = _lib.launch_count()
                with torch.cuda.graph(gph):
                    step_eager(bi, opt)
                graph_launches.append(_lib.launch_count() - lc0)
                graphs.append(gph)
            sync_all()
            for gph in graphs:      # one replay each: first-replay initialisation stays out of the timed region
                gph.replay()
            sync_all()
        except Exception as e:   # capture is an optimisation, never a requirement
            log(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly")
            graphs = None
            torch.cuda.synchronize()

    def run_step(i):
        if graphs is not None:
            graphs[i % len(graphs)].replay()
        else:
            step_eager(i, opt)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i_soak in range(120):      # ~1 s of untimed replays under the nvidia-smi poll; a FIXED count on every rank
        run_step(i_soak)
        if i_soak % 8 == 7:
            torch.cuda.synchronize()
    sync_all()
    l0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    cuprof = bool(os.environ.get("VQGNN_CUPROF"))   # ncu --profile-from-start off: capture the timed steps only
    if cuprof:
        torch.cuda.profiler.start()
    for i in range(args.steps):
        flush_l2()
        ev[i][0].record()
        run_step(i)
        ev[i][1].record()
    sync_all()
    if cuprof:
        torch.cuda.profiler.stop()
    launches = _lib.launch_count() - l0 - args.steps   # minus the flush launches
    if graphs is not None:   # launches replayed from the graphs (counted once, at capture)
        launches = sum(graph_launches[i % len(graphs)] for i in range(args.steps))
    clocks = sampler.stop()
    t_ms = sum(a.elapsed_time(b) for a, b in ev)
    t_t = torch.tensor([t_ms], dtype=torch.float64, device=dev)
    nodes = torch.tensor([step_nodes], dtype=torch.float64, device=dev)
    if distributed:
        torch.distributed.all_reduce(t_t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(nodes)
    t_ms = float(t_t.item())
    ms_per_step = t_ms / args.steps
    value = float(nodes.item()) * c["layers"] / (t_ms / 1e3)
    model.check_status()
    replica_div = 0.0
    if distributed:   # codebook / code-table replicas must stay identical across ranks
        replica_div = max(max(vdist.replicas_max_abs_diff(l.bank.E), vdist.replicas_max_abs_diff(l.bank.size),
                              vdist.replicas_max_abs_diff(l.bank.codes.float()))
                          for l in model.convs)

    # ---- attribution pass: CUDA events around every C-ABI launch (dominant kernel + roofline) ---
    roofline, kernel_table = None, {}
    _lib.PROFILER.enabled = True     # every rank runs the pass (its steps contain collectives); rank 0 reports
    _lib.PROFILER.reset()
    n_attr = min(args.steps, 8)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for i in range(n_attr):
        flush_l2()
        a0.record()
        step_eager(i, opt_eager)
        a1.record()
        torch.cuda.synchronize()
        tot += a0.elapsed_time(a1)
    summ = _lib.PROFILER.summary()
    _lib.PROFILER.enabled = False
    summ.pop("vqgnn_flush_l2", None)
    for k, (n, ms) in sorted(summ.items(), key=lambda kv: -kv[1][1]):
        kernel_table[k] = {"launches_per_step": n / n_attr, "ms_per_step": ms / n_attr,
                           "share_of_step": (ms / n_attr) / (tot / n_attr)}
    top = max(((k, v) for k, v in summ.items() if not k.startswith("nccl")), key=lambda kv: kv[1][1])[0]
    # algorithmic work of EXACTLY the launches whose time is averaged: every layer of every attribution step
    dims = [c["C_in"]] + [c["hidden"]] * (c["layers"] - 1)
    work, bound = 0.0, "hbm"
    for i in range(n_attr):
        plan_i = plans[i % len(plans)]
        for li, cin in enumerate(dims):
            w_, bound = algorithmic_work(top, plan_i, cin, cin // c["D"], c["M"], plan_i.B)
            work += w_
            if top == "vqgnn_mp_fwd_rows" and li > 0 and plan_i.version == 'v2':
                # the v2 backward of layers 2..L runs through the same entry point (transposed CSR): count its bytes
                work += algorithmic_work("vqgnn_mp_fwd_rows:bwd", plan_i, cin, cin // c["D"], c["M"], plan_i.B)[0]
    n_launch, ms_total = summ[top]
    avg_ms = ms_total / n_launch
    per_launch_work = work / n_launch
    if bound == "hbm":
        achieved, peak, unit = per_launch_work / (avg_ms * 1e-3) / 1e9, pk["hbm"], "GB/s"
    else:
        achieved, peak, unit = per_launch_work / (avg_ms * 1e-3) / 1e12, pk["tf_sust"], "TFLOP/s"
    traffic, tnote = None, None
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.isfile(tpath):      # ncu --set full capture of the same kernel at this config (per launch)
        tj = json.load(open(tpath)).get(c["name"], {}).get(top)
        if tj:
            traffic = sum(tj["per_launch_bytes"]) / len(tj["per_launch_bytes"])
            tnote = tj.get("note")
    traffic_rate = None
    if traffic is not None and bound == "hbm":      # what the kernel actually pulls through HBM (ncu), next to `achieved`
        traffic_rate = {"GBps": traffic / (avg_ms * 1e-3) / 1e9, "frac_of_peak": traffic / (avg_ms * 1e-3) / 1e9 / pk["hbm"]}
    roofline = {"kernel": top, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                "traffic_rate": traffic_rate,
                "frac": achieved / peak, "traffic": traffic, "traffic_note": tnote, "peak_source": pk["source"],
                "avg_launch_ms": avg_ms, "launches_per_step": n_launch / n_attr,
                "algorithmic_work_per_launch": per_launch_work,
                "note": "bytes and time are summed over the same launches (all layers of the attribution steps)"}

    # ---- CPU baseline (rank 0, N=1 only) ---------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nb0 = wl.node_lists[0]
        cb = wl.make_batch(nb0[:args.cpu_batch] if (args.cpu_batch and args.cpu_batch < nb0.numel()) else nb0)
        v, t_step, n_timed, kind, what = cpu_reference_run(c, batch_to_cpu(cb), N, 2, 1, budget_s=120.0)
        cpu_baseline = {"value": v, "unit": "nodes/s/layer", "cores": os.cpu_count(), "kind": kind,
                        "sample": f"{what}: full {c['layers']}-layer train step on one batch of the same workload "
                                  f"(B={cb[0].shape[0]}), median of {n_timed} timed steps after the warm-start pass; "
                                  f"{t_step * 1e3:.0f} ms per step"}

    if rank == 0:
        impl_name = {"auto": f"auto ({'tcgen05' if c['M'] >= 512 else 'simt-fp32'} at M={c['M']})", "1": "tcgen05",
                     "0": "simt-fp32"}[str(args.assign_impl)]
        line = {"metric": "train nodes/sec per VQ-GNN layer", "value": value, "unit": "nodes/s/layer",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "nodes/s/layer", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": 4, "passes_ms": [round(t, 2) for t in e2e_passes],
                        "untimed_settling_passes_ms": warm_passes,
                        "note": "K steps per pass, MEDIAN of three passes; host sends node ids only, the batch graph "
                                "is built on the device from the resident graph (prepare_from_graph), features and "
                                "labels are gathered from resident matrices; eager launches, DevicePrefetcher"},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
                "replica_max_abs_diff": replica_div, "graph_identical_on_all_ranks": graph_identical,
                "batch_nodes_per_rank": [int(b[0].shape[0]) for b in batches],
                "cuda_graphs": graphs is not None, "kernels": kernel_table, "assign_impl": impl_name,
                "warm_start": warm_start_info,
                "vq_updates": "compute stream" if args.sync_vq else "side stream (overlapped with the rest of backward)"}
        print(json.dumps(line), flush=True)
    if distributed:
        torch.distributed.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(), sys.stderr.flush()
        # captured graphs hold NCCL work: tearing the process group down under them can hang, and the JSON line is
        # already out -- leave without running destructors
        os._exit(0)


def _num_nodes(cfg, scale):
    from vq_gnn_b200 import synth
    return int(synth.CONFIG_SHAPES[cfg["shape"]]["N"] * scale)


if __name__ == "__main__":
    main()
