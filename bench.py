#!/usr/bin/env python
"""bench.py — train nodes/sec per VQ-GNN layer on B200 (BASELINE.json metric).

Workload at N=1 (configs[1]): VQ-GNN SAGE-Mean on a synthetic Reddit-shaped graph (232,965 nodes,
~114.6M directed edges, 602-d padded to 604, 41 classes), hidden 128, num-M 1024, num-D 4, batch 6000,
cont sampler walk 3, v1 formulation, --warm-up --bn-flag --recovery-flag (README.md:79-82).
One "step" = one pass of the hot path over one batch: 3-layer LowRankGNN forward + loss + backward
(the VQ assignment + EMA update of every layer fire inside backward) + RMSprop step.
value = B * num_layers / t_step   [batch nodes / second / layer].

N > 1: each rank owns a contiguous node partition and its own batch; codebook statistics
(per-codeword sums/counts + whitening moments) and dense weight gradients are allreduced with NCCL
each step (weak scaling: per-GPU batch fixed).

Keys: see the task contract.  `roofline` is for the dominant kernel of the step (found live with CUDA
events around every C-ABI launch), `cpu_baseline` times the oracle port (oracle/restate.py) on the
host cores on a bounded sample, `e2e` runs the same step from pinned HOST buffers through the public
API (H2D copies + plan construction inside the timed region, loss read back).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(name="c2_reddit_sage_v1", N=232_965, E=57_307_946, feat=602, C_in=604, hidden=128, classes=41,
           M=1024, D=4, B=6000, walk=3, layers=3, conv="SAGE", version="v1", power_law=2.2)


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
        # median over the busier half of the samples (under load)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# workload construction (synthetic, seeded; generated on the GPU with torch ops -- not timed)
# ------------------------------------------------------------------------------------------------
def build_workload(dev, rank: int, world: int, scale: float = 1.0, n_batches: int = 4):
    from vq_gnn_b200 import sampling, synth
    c = CFG
    N, E = int(c["N"] * scale), int(c["E"] * scale)
    t0 = time.time()
    rowptr, row, col = synth.random_edges(N, E, seed=0, power_law=c["power_law"], device=dev)
    g = synth.normalized_graph(N, rowptr, row, col, c["conv"], c["version"])
    del row
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    # rank r samples its seeds from its own contiguous node partition
    from vq_gnn_b200 import dist as vdist
    lo, hi = vdist.partition_range(N, rank, world)
    seeds = lo + torch.randperm(hi - lo, generator=gen, device=dev)[:c["B"]]
    node_lists = sampling.cont_sampler(g, seeds, c["walk"], c["B"], generator=gen)[:n_batches]
    batches = []
    for nodes in node_lists:
        x = torch.randn(nodes.numel(), c["C_in"], generator=gen, device=dev)
        x[:, c["feat"]:] = 0          # zero padding to a multiple of num_D (v2/utils/misc.py:212-219)
        y = torch.randint(0, c["classes"], (nodes.numel(),), generator=gen, device=dev)
        batches.append((x, sampling.collate_batch_v1(g, nodes, True, True), y))
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    nnz_bn = [int(b[1][1][0].numel()) for b in batches]
    nnz_bb = [int(b[1][2][0].numel()) for b in batches]
    log(f"[bench] graph N={N} nnz={g.nnz} built in {time.time() - t0:.1f}s; batches B={[b[0].shape[0] for b in batches]} "
        f"nnz(A_BN)={nnz_bn} nnz(A_BB)={nnz_bb}")
    return g, batches


def build_model(dev, N, distributed: bool, assign_impl='auto'):
    import vq_gnn_b200 as V
    c = CFG
    torch.manual_seed(0)
    model = V.LowRankGNN(c["C_in"], c["hidden"], c["classes"], c["layers"], 0.0, c["M"], c["D"], N,
                         no_second_fc=True, skip=False, commitment_cost=0.0, grad_scale=[1, 1], act="leaky_gelu",
                         bn_flag=True, warm_up_flag=True, momentum=0.1, conv_type=c["conv"], version=c["version"])
    model = model.to(dev).train()
    for layer in model.convs:
        layer.bank.assign_impl = assign_impl if assign_impl == 'auto' else int(assign_impl)
        layer.bank.distributed = distributed
    return model


def warm_start(model, batches):
    """The reference's init(): layer-wise feature-only codebook warm start (v1/main_node.py init())."""
    with torch.no_grad():
        for layer_idx in range(1, model.num_layers + 1):
            for x, bA, _ in batches:
                model.init((x, bA), layer_idx)
    model.set_inited(True)
    model.check_status()


def train_step(model, opt, x, batch_A, y, distributed: bool):
    opt.zero_grad(set_to_none=True)
    out, _, info = model((x, batch_A), 1)
    loss = F.cross_entropy(out, y) + info
    loss.backward()
    if distributed:
        from vq_gnn_b200 import dist as vdist
        vdist.allreduce_mean_grads_(model.parameters())
    opt.step()
    return loss


# ------------------------------------------------------------------------------------------------
# algorithmic work per launch, used for the roofline of the dominant kernel
# ------------------------------------------------------------------------------------------------
def algorithmic_work(kernel: str, plan, C: int, nb: int, M: int, B: int):
    nnz = plan.nnz
    nnz_t = int(plan.bwd_col.numel())
    tail = int((plan.fwd_col >= B).sum())
    if kernel == "vqgnn_mp_fwd":    # v1 form: nnz(A_BN)*(8 + 4 rval + nb*2 codes) + in-batch*8 + x r + y,gq w + codebook
        by = tail * (12 + nb * 2) + (nnz - tail) * 8 + (B + 1) * 4 + 3 * B * C * 4 + M * C * 2 * 4
        return by, "hbm"
    if kernel == "vqgnn_mp_fwd_tail":
        # per tail entry: node id 4 + val 4 + rval 4 + nb codes x 2 B; per batch row: x read + y, gq read-modify-write;
        # plus the codebooks (feature + gradient halves) once
        by = tail * (12 + nb * 2) + (B + 1) * 4 + 5 * B * C * 4 + M * C * 2 * 4
        return by, "hbm"
    if kernel == "vqgnn_mp_bwd":
        by = nnz_t * 8 + (B + 1) * 4 + 3 * B * C * 4
        return by, "hbm"
    if kernel == "vqgnn_vq_assign":  # joint: FLOPs = 2*B*M*2C
        return 4.0 * B * M * C, "tensor"
    if kernel == "vqgnn_vq_moments":
        return 2 * B * C * 4, "hbm"
    if kernel == "vqgnn_vq_finalize":
        return nb * M * (12 * 4 + 8 * 4 * 4 + 4 * 2), "hbm"
    return 0, "hbm"


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(batches_cpu, n_branches: int, steps: int, warmup: int, N: int):
    """Times oracle/restate.py (the CPU restatement of the reference; the Python reference itself cannot
    travel to the GPU box).  Sample: `n_branches` of the 32 branches of ONE hidden layer (C=128) of the
    config-2 batch, fwd + bwd + VQ update, scaled to the full layer."""
    from oracle import restate
    c = CFG
    torch.set_num_threads(os.cpu_count())
    x_full, bA, _ = batches_cpu[0]
    B = x_full.shape[0]
    C = c["hidden"]
    nb = C // c["D"]
    torch.manual_seed(0)
    layer = restate.OracleLayer(C, C, c["M"], c["D"], N, c["conv"], c["version"], warm_up_flag=True,
                                sparse=True, branches=list(range(n_branches)))
    layer.params = {"gnn_transform.weight": (torch.randn(C, C) * 0.05).requires_grad_(True),
                    "gnn_transform.bias": torch.zeros(C, requires_grad=True),
                    "fc_sage.weight": (torch.randn(C, C) * 0.05).requires_grad_(True),
                    "fc_sage.bias": torch.zeros(C, requires_grad=True)}
    x = torch.randn(B, C, generator=torch.Generator().manual_seed(1))
    w = torch.randn(B, C, generator=torch.Generator().manual_seed(2))
    layer.train()
    times = []
    for s in range(warmup + steps):
        if s == 1:
            layer.set_inited(True)
        xx = x.clone().requires_grad_(True)
        t0 = time.perf_counter()
        out, info = layer(xx, bA, 1.0, False)
        ((out * w).sum() + info).backward()
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    t_layer = (sum(times) / len(times)) * (nb / n_branches)
    return B / t_layer, t_layer, B


def batch_to_cpu(batch):
    x, bA, y = batch
    mv = lambda t: None if t is None else (tuple(u.cpu() for u in t) if isinstance(t, tuple) else t.cpu())
    return x.cpu(), tuple(mv(t) for t in bA), y.cpu()


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the synthetic graph (debug only)")
    ap.add_argument("--assign-impl", type=str, default=os.environ.get("VQGNN_ASSIGN_IMPL", "auto"),
                    help="auto (default: tcgen05/TMEM kernel when M >= 512), 1 = always tcgen05, 0 = exact-fp32 SIMT")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true",
                    help="launch the device-resident steps eagerly instead of replaying one CUDA graph per batch")
    ap.add_argument("--cpu-branches", type=int, default=4)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    c = CFG
    config = {"workload": "configs[1]: VQ-GNN SAGE-Mean (v1 formulation), synthetic Reddit-shaped graph "
                          f"N={c['N']}, ~{2 * c['E'] / 1e6:.1f}M directed edges, 602-d (padded 604), hidden 128, "
                          "num-M 1024, num-D 4, batch 6000, cont sampler walk 3, 3 layers (604->128->128->41)",
              "step": "3-layer fwd + CE loss + bwd (VQ assign + EMA update of every layer inside bwd) + RMSprop",
              "value_formula": "B * num_layers / t_step", "batch_nodes": c["B"], "num_layers": c["layers"],
              "parallelism": f"dp{world} (node-partitioned batches; whitening moments + EMA stats + weight grads "
                             "allreduced, code-table updates all-gathered)",
              "l2": "4 distinct batches rotated AND a 256 MiB L2 flush between timed steps",
              "launch": "one CUDA graph per resident batch (whole train step incl. NCCL) replayed; --no-graphs = eager"}

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
        g, batches = build_workload(dev, 0, 1, args.scale, n_batches=1)
        batches_cpu = [batch_to_cpu(b) for b in batches]
        steps, warm = max(1, min(args.steps, 5)), 1
        v, t_layer, B = cpu_reference_step(batches_cpu, args.cpu_branches, steps, warm, g.N)
        sample = (f"oracle port (oracle/restate.py, sparse mapper), {args.cpu_branches}/32 branches of one hidden "
                  f"layer (C=128) of the config-2 batch (B={B}), fwd+bwd+VQ update, scaled x{32 // args.cpu_branches}; "
                  f"{steps} timed steps after {warm} warm-up")
        line = {"impl": "reference", "metric": "train nodes/sec per VQ-GNN layer", "value": v,
                "unit": "nodes/s/layer", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
                "ms_per_step": t_layer * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "nodes/s/layer", "cores": os.cpu_count(), "kind": "port",
                                 "sample": sample},
                "e2e": {"value": v, "unit": "nodes/s/layer", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ our arm
    import vq_gnn_b200 as V
    from vq_gnn_b200 import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        torch.distributed.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    _lib.require_device(torch.zeros(1, device=dev))
    pk = peaks()

    g, batches = build_workload(dev, rank, world, args.scale)
    N = g.N
    model = build_model(dev, N, distributed, args.assign_impl)
    use_graphs = not args.no_graphs
    opt = torch.optim.RMSprop(model.parameters(), lr=1e-3, alpha=0.99, capturable=use_graphs)
    opt_eager = torch.optim.RMSprop(model.parameters(), lr=1e-3, alpha=0.99) if use_graphs else opt
    warm_start(model, batches)
    plans = [model.prepare(b[1]) for b in batches]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _lib.load()

    def flush_l2():
        lib.vqgnn_flush_l2(_lib.ptr(flush), flush.numel(), _lib.stream())

    def sync_all():
        if distributed:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up, then exactly K steps, per-step CUDA events ----------
    for i in range(args.warmup):
        x, _, y = batches[i % len(batches)]
        train_step(model, opt, x, plans[i % len(plans)], y, distributed)
    sync_all()

    # ---- end-to-end through the public API from pinned host buffers (eager; measured BEFORE the CUDA graphs of
    #      the device-resident timing exist: their private memory pools slow later eager allocation down) ----
    host = []
    for x, bA, y in batches:
        pin = lambda t: None if t is None else (tuple(u.cpu().pin_memory() for u in t) if isinstance(t, tuple)
                                                else t.cpu().pin_memory())
        host.append((x.cpu().pin_memory(), tuple(pin(t) for t in bA), y.cpu().pin_memory()))

    def nbytes(t):
        if t is None:
            return 0
        if isinstance(t, tuple):
            return sum(nbytes(u) for u in t)
        return t.numel() * t.element_size()
    h2d = sum(nbytes(h[0]) + nbytes(h[1]) + nbytes(h[2]) for h in host) / len(host)

    def to_dev(t):
        if t is None:
            return None
        if isinstance(t, tuple):
            return tuple(u.to(dev, non_blocking=True) for u in t)
        return t.to(dev, non_blocking=True)

    # The user-facing loop: DevicePrefetcher uploads batch i+1 (H2D from pinned memory + batch-plan construction)
    # on a side stream while batch i trains; every step's H2D copies, plan build, forward, backward, VQ update,
    # optimiser step and the D2H read of the loss are inside the timed region.
    from vq_gnn_b200.loader import DevicePrefetcher

    def prep(b):
        x, bA, y = b
        return x, model.prepare(bA), y

    side = torch.cuda.Stream(device=dev)

    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()     # double-buffered D2H target for the step results

    def e2e_run(n):
        """Every step: H2D of the batch, plan build, train step, and a D2H copy of the step's loss.  The loss of
        step i is copied asynchronously into pinned memory and READ on the host one step later (after its event),
        so the host never idles the GPU waiting for the step it has just launched."""
        pf = DevicePrefetcher(host, dev, prepare=prep, count=n, stream=side)
        pending, last = None, 0.0
        for i in range(n):
            x, plan, y = pf.next()
            flush_l2()
            loss = train_step(model, opt_eager, x, plan, y, distributed)
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

    e2e_run(3 * len(host) + 2)   # lets the caching allocator's side-stream pool converge (no cudaMalloc in the timed run)
    sync_all()
    e2e_passes = []
    for _ in range(4):       # four passes of K steps each; the fastest one is reported: the eager e2e loop is bound by
        sync_all()           # the host's Python launch rate and a shared host shows 2x pass-to-pass noise
        t0 = time.perf_counter()
        e2e_run(args.steps)
        torch.cuda.synchronize()
        e2e_passes.append((time.perf_counter() - t0) * 1e3)
    e2e_ms = min(e2e_passes)
    e_t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if distributed:
        torch.distributed.all_reduce(e_t, op=torch.distributed.ReduceOp.MAX)
    e2e_nodes = torch.tensor([float(sum(batches[i % len(batches)][0].shape[0] for i in range(args.steps)))],
                             dtype=torch.float64, device=dev)
    if distributed:
        torch.distributed.all_reduce(e2e_nodes)
    e2e_value = float(e2e_nodes.item()) * c["layers"] / (float(e_t.item()) / 1e3)


    # The step is launch-bound (~550 kernel launches, ~5 ms of kernels): capture one CUDA graph per resident batch
    # (the plan's shapes differ per batch) and replay it, so the GPU is never waiting on Python.  Every kernel of
    # the step -- forward, backward, the VQ updates, the NCCL allreduces and the optimiser -- is inside the graph.
    graphs, graph_launches = None, []
    if use_graphs:
        try:
            graphs = []
            side_cap = torch.cuda.Stream(device=dev)
            for bi, (x, _, y) in enumerate(batches):
                side_cap.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side_cap):
                    train_step(model, opt, x, plans[bi], y, distributed)       # allocator warm-up on the side stream
                torch.cuda.current_stream().wait_stream(side_cap)
                gph = torch.cuda.CUDAGraph()
                lc0 = _lib.launch_count()
                with torch.cuda.graph(gph):
                    train_step(model, opt, x, plans[bi], y, distributed)
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
            x, _, y = batches[i % len(batches)]
            train_step(model, opt, x, plans[i % len(plans)], y, distributed)
    sync_all()
    # nvidia-smi polls every 100 ms while the device-resident steps run: ~1 s of untimed replays of the same steps first
    # (the K timed steps alone last ~0.1 s), then the timed steps.  Not started earlier: polling nvidia-smi during
    # the eager e2e loop measurably stalls its CUDA API calls (e2e dropped from 3.1 M to 1.3 M nodes/s/layer).
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i_soak in range(200):      # a FIXED count: every rank must replay the same number of (collective-carrying) steps
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
    nodes = torch.tensor([float(sum(batches[i % len(batches)][0].shape[0] for i in range(args.steps)))],
                         dtype=torch.float64, device=dev)
    if distributed:
        torch.distributed.all_reduce(t_t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(nodes)
    t_ms = float(t_t.item())
    ms_per_step = t_ms / args.steps
    value = float(nodes.item()) * c["layers"] / (t_ms / 1e3)
    model.check_status()
    replica_div = None
    if distributed:   # codebook replicas must stay identical across ranks
        from vq_gnn_b200 import dist as vdist
        replica_div = max(max(vdist.replicas_max_abs_diff(l.bank.E), vdist.replicas_max_abs_diff(l.bank.size),
                              vdist.replicas_max_abs_diff(l.bank.codes.float()))
                          for l in model.convs)

    # ---- attribution pass: CUDA events around every C-ABI launch (dominant kernel + roofline) ---
    roofline, kernel_table = None, {}
    if True:   # every rank runs the pass (its steps contain collectives); rank 0 reports
        _lib.PROFILER.enabled = True
        _lib.PROFILER.reset()
        n_attr = min(args.steps, 8)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for i in range(n_attr):
            x, _, y = batches[i % len(batches)]
            flush_l2()
            a0.record()
            train_step(model, opt_eager, x, plans[i % len(plans)], y, distributed)
            a1.record()
            torch.cuda.synchronize()
            tot += a0.elapsed_time(a1)
        summ = _lib.PROFILER.summary()
        _lib.PROFILER.enabled = False
        summ.pop("vqgnn_flush_l2", None)
        for k, (n, ms) in sorted(summ.items(), key=lambda kv: -kv[1][1]):
            kernel_table[k] = {"launches_per_step": n / n_attr, "ms_per_step": ms / n_attr,
                               "share_of_step": (ms / n_attr) / (tot / n_attr)}
        top = max(summ.items(), key=lambda kv: kv[1][1])[0]
        # per-launch algorithmic work, summed over the launches of one step (3 layers), batch 0 shapes
        plan0, B0 = plans[0], plans[0].B
        work, bound = 0.0, "hbm"
        for (cin, _cout) in [(c["C_in"], c["hidden"]), (c["hidden"], c["hidden"]), (c["hidden"], c["classes"])]:
            w_, bound = algorithmic_work(top, plan0, cin, cin // c["D"], c["M"], B0)
            work += w_
        n_launch, ms_total = summ[top]
        avg_ms = ms_total / n_launch
        per_launch_work = work / (n_launch / n_attr)
        if bound == "hbm":
            achieved, peak, unit = per_launch_work / (avg_ms * 1e-3) / 1e9, pk["hbm"], "GB/s"
        else:
            achieved, peak, unit = per_launch_work / (avg_ms * 1e-3) / 1e12, pk["tf_sust"], "TFLOP/s"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.isfile(tpath):      # ncu --set full capture of the same kernel (per launch, averaged over a step)
            tj = json.load(open(tpath)).get(top)
            if tj:
                traffic = sum(tj["per_launch_bytes"]) / len(tj["per_launch_bytes"])
        roofline = {"kernel": top, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                    "frac": achieved / peak, "traffic": traffic,
                    "traffic_note": "ncu dram bytes per launch, averaged over the three layer launches of one 5.1 M-entry "
                                    "batch; below the algorithmic bytes because the code table is partly L2-resident" if traffic else None,
                    "peak_source": pk["source"],
                    "avg_launch_ms": avg_ms, "launches_per_step": n_launch / n_attr,
                    "algorithmic_work_per_launch": per_launch_work}

    # ---- CPU baseline (rank 0, N=1 only) ---------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, t_layer, Bc = cpu_reference_step([batch_to_cpu(batches[0])], args.cpu_branches, 2, 1, N)
        cpu_baseline = {"value": v, "unit": "nodes/s/layer", "cores": os.cpu_count(), "kind": "port",
                        "sample": f"oracle port (oracle/restate.py, sparse mapper): {args.cpu_branches}/32 branches of "
                                  f"one hidden layer (C=128) of the same batch (B={Bc}), fwd+bwd+VQ update, scaled "
                                  f"x{32 // args.cpu_branches}; 2 timed steps after 1 warm-up; "
                                  f"{t_layer * 1e3:.0f} ms per layer"}

    if rank == 0:
        line = {"metric": "train nodes/sec per VQ-GNN layer", "value": value, "unit": "nodes/s/layer",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "nodes/s/layer", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": 4, "passes_ms": [round(t, 2) for t in e2e_passes],
                        "note": "K steps per pass, fastest of four passes; eager launches, DevicePrefetcher"},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
                "replica_max_abs_diff": replica_div, "cuda_graphs": graphs is not None, "kernels": kernel_table, "assign_impl": {"auto": "tcgen05 (auto: M=1024)", "1": "tcgen05", "0": "simt-fp32"}[str(args.assign_impl)]}
        print(json.dumps(line), flush=True)
    if distributed:
        torch.distributed.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(), sys.stderr.flush()
        # captured graphs hold NCCL work: tearing the process group down under them can hang, and the JSON line is
        # already out -- leave without running destructors
        os._exit(0)


if __name__ == "__main__":
    main()
