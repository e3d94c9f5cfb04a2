"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): node partitioning, the flat gradient
allreduce, and the VQ update's two exchange points.  The staged update below is the CPU
restatement of what `VQBank.run` does between its kernels on every rank
    moments -> [allreduce] -> whiten -> assign + per-codeword sums/counts -> [allreduce] -> EMA / recovery
and must reproduce the single-process oracle (`oracle.restate.OracleVQ.update`) run on the concatenated
batch: identical codes, identical (bit-exact) count histogram, state within fp32 tolerance, and identical
replicas on both ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restate
from vq_gnn_b200 import dist as vdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _staged_update(o: restate.OracleVQ, X, G, group=None):
    """One rank's share of `update` with the two sums exchanged (vq.py:204-279 split as in VQBank.run)."""
    D, W, M = o.D, o.W, o.M
    inp = torch.cat([X, G], 1).double()
    sums = torch.cat([inp.sum(0), (inp ** 2).sum(0), torch.tensor([float(X.shape[0])], dtype=torch.float64)])
    vdist.allreduce_sum_(sums, group)                                   # exchange 1: whitening moments
    n = sums[-1]
    mean = (sums[:W] / n)
    var_b = (sums[W:2 * W] / n - mean ** 2).clamp_min(0)
    var_u = var_b * n / (n - 1)
    mean_f, var_bf, var_uf = mean.float(), var_b.float(), var_u.float()
    eps = torch.cat([torch.full((D,), 1e-5), torch.full((W - D,), o.eps)])
    mom = torch.cat([torch.full((D,), 0.1), torch.full((W - D,), o.momentum)])
    run_m = torch.cat([o.feat_mean, o.grad_mean])
    run_v = torch.cat([o.feat_var, o.grad_var])
    if not o.bn_inited:                                                  # vq.py:216-221
        run_m, run_v = mean_f.clone(), var_uf.clone()
        o.bn_inited = True
    run_m = (1 - mom) * run_m + mom * mean_f
    run_v = (1 - mom) * run_v + mom * var_uf
    o.feat_mean, o.grad_mean, o.feat_var, o.grad_var = run_m[:D], run_m[D:], run_v[:D], run_v[D:]
    z = (torch.cat([X, G], 1) - mean_f) / torch.sqrt(var_bf + eps)
    z[:, D:2 * D] *= o.scale[0]
    if o.add:
        z[:, 2 * D] *= o.scale[1]
    idx = torch.argmin(o.distances(z, o._embedding), 1)
    stats = torch.zeros(M, W + 1)
    stats[:, :W].index_add_(0, idx, z)
    stats[:, W] = torch.bincount(idx, minlength=M).float()
    vdist.allreduce_sum_(stats, group)                                  # exchange 2: per-codeword sums + counts
    o._ema_size(stats[:, W])
    o._ema_w = o._ema_w * o.decay + (1 - o.decay) * stats[:, :W]
    o._embedding = o._ema_w / o._ema_cluster_size.unsqueeze(1)
    out = o._embedding.clone()
    out[:, D:2 * D] /= o.scale[0] + o.eps
    if o.add:
        out[:, 2 * D] /= o.scale[1] + o.eps
    var = torch.cat([o.feat_var + 1e-5, o.grad_var + o.eps])
    o._embedding_output = out * torch.sqrt(var).unsqueeze(0) + torch.cat([o.feat_mean, o.grad_mean]).unsqueeze(0)
    return idx, stats[:, W].clone()


def _worker(rank, world_size, port, tmp):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        # ---- partitioning ---------------------------------------------------------------------
        N = 1003
        lo, hi = vdist.partition_range(N, rank, world_size)
        spans = [torch.zeros(2, dtype=torch.long) for _ in range(world_size)]
        dist.all_gather(spans, torch.tensor([lo, hi]))
        assert spans[0][0] == 0 and spans[-1][1] == N
        for a, b in zip(spans[:-1], spans[1:]):
            assert a[1] == b[0]
        # ---- flat gradient allreduce ----------------------------------------------------------
        torch.manual_seed(0)
        lin = torch.nn.Linear(5, 3)
        for p in lin.parameters():
            p.grad = torch.full_like(p, float(rank + 1))
        vdist.allreduce_mean_grads_(lin.parameters())
        for p in lin.parameters():
            assert torch.allclose(p.grad, torch.full_like(p, (1 + world_size) / 2))
        # ---- staged VQ update with the two exchanges vs the single-process oracle ----------------
        M, D, Bfull = 32, 4, 600
        g = torch.Generator().manual_seed(5)
        torch.manual_seed(3)
        full = restate.OracleVQ(M, D, grad_normalize_scale=[1, 1], warm_up_flag=True)
        torch.manual_seed(3)
        mine = restate.OracleVQ(M, D, grad_normalize_scale=[1, 1], warm_up_flag=True)
        for step in range(3):
            X = torch.randn(Bfull, D, generator=g) * 2 + 0.5
            G = torch.randn(Bfull, D, generator=g) * 1e-2
            cut = Bfull * 2 // 5                                         # uneven shards
            sl = slice(0, cut) if rank == 0 else slice(cut, Bfull)
            idx, counts = _staged_update(mine, X[sl], G[sl])
            idx_full, _ = full.update(X, G)
            assert torch.equal(idx, idx_full.squeeze(1)[sl]), step          # codes bit-exact
            assert torch.equal(counts, torch.bincount(idx_full.squeeze(1), minlength=M).float())
            for k, v in full.dump().items():
                assert torch.allclose(mine.dump()[k], v, rtol=1e-4, atol=1e-6), (step, k)
            for k, v in mine.dump().items():                              # replicas identical on every rank
                assert vdist.replicas_max_abs_diff(v) == 0.0, (step, k)
        # ---- code-table updates: every replica applies every rank's re-assignments, in rank order ------
        Nn, nbr, Bc = 50, 3, 8
        table = torch.zeros(Nn, nbr, dtype=torch.int16)
        gg = torch.Generator().manual_seed(11)                           # same stream on both ranks
        idx_all = [torch.randperm(Nn, generator=gg)[:Bc].to(torch.int32) for _ in range(world_size)]
        idx_all[1][0] = idx_all[0][0]                                    # a node shared by two ranks' batches
        new_all = [torch.randint(0, 99, (Bc, nbr), generator=gg, dtype=torch.int16) for _ in range(world_size)]
        table[idx_all[rank].long()] = new_all[rank]                      # own update (what the assign kernel does)
        gidx = vdist.allgather_code_updates_(table, idx_all[rank], new_all[rank])
        want = torch.zeros(Nn, nbr, dtype=torch.int16)
        for rr in range(world_size):
            want[idx_all[rr].long()] = new_all[rr]
        assert torch.equal(table, want) and gidx.numel() == world_size * Bc
        assert vdist.replicas_max_abs_diff(table.float()) == 0.0
        # ---- ranks with DIFFERENT batch sizes: refused without a capacity, padded (node id -1) with one ----------
        Bs = [5, 8]
        idx_r = torch.randperm(Nn, generator=torch.Generator().manual_seed(20 + rank))[:Bs[rank]].to(torch.int32)
        new_r = torch.randint(0, 99, (Bs[rank], nbr), generator=torch.Generator().manual_seed(30 + rank),
                              dtype=torch.int16)
        try:
            vdist.allgather_code_updates(idx_r, new_r)
            raise AssertionError("unequal batch sizes must be refused when no capacity is given")
        except ValueError:
            pass
        t2 = torch.zeros(Nn, nbr, dtype=torch.int16)
        gidx2 = vdist.allgather_code_updates_(t2, idx_r, new_r, capacity=8)
        assert gidx2.numel() == world_size * 8 and int((gidx2 < 0).sum()) == 3
        want2 = torch.zeros(Nn, nbr, dtype=torch.int16)
        for rr in range(world_size):
            ir = torch.randperm(Nn, generator=torch.Generator().manual_seed(20 + rr))[:Bs[rr]]
            want2[ir] = torch.randint(0, 99, (Bs[rr], nbr), generator=torch.Generator().manual_seed(30 + rr),
                                      dtype=torch.int16)
        assert torch.equal(t2, want2) and vdist.replicas_max_abs_diff(t2.float()) == 0.0
        with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(2))


def test_single_process_helpers_are_noops():
    t = torch.ones(3)
    assert vdist.world() == (0, 1)
    assert torch.equal(vdist.allreduce_sum_(t.clone()), t)
    assert vdist.replicas_max_abs_diff(t) == 0.0
    assert vdist.partition_range(10, 0, 1) == (0, 10)
