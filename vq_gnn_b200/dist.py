"""Multi-GPU plumbing (one process per GPU, `torch.distributed`; NCCL on the GPUs, gloo in the CPU tests).

The reference has no distributed code; the design is SURVEY.md §8e:
  * every rank owns a contiguous node range (`partition_range`) and samples its mini-batch from it,
  * every rank holds a replica of each layer's codebooks and code table,
  * the ONLY exchanges are sums: the whitening moments and the per-codeword statistics inside the VQ update
    (`allreduce_sum_`, called by `VQBank.run` between its kernels) and the dense weight gradients once per
    step (`allreduce_mean_grads_`).  Counts are integers held in fp32/fp64 (< 2^24): the reduced histogram is
    bit-exact for any reduction order, which is what keeps the replicas identical.
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist

Tensor = torch.Tensor


def world(group=None) -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def partition_range(N: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous node range [lo, hi) owned by `rank` (ranges tile [0, N) without overlap)."""
    return rank * N // world_size, (rank + 1) * N // world_size


def _timed(name: str, fn):
    """Run fn() bracketed by CUDA events when bench.py's per-kernel attribution is on."""
    from . import _lib
    if not (_lib.PROFILER.enabled and torch.cuda.is_available()):
        return fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    _lib.PROFILER.records.append((name, a, b, None))
    return out


def allreduce_sum_(t: Tensor, group=None) -> Tensor:
    """In-place sum over ranks; a no-op for a single process."""
    if world(group)[1] > 1:
        _timed("nccl_allreduce_stats", lambda: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group))
    return t


def allreduce_mean_grads_(params: Iterable[Tensor], group=None) -> None:
    """One flat allreduce of every available `.grad`, averaged over ranks (data parallel)."""
    ws = world(group)[1]
    if ws == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    _timed("nccl_allreduce_grads", lambda: dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group))
    flat /= ws
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def allgather_code_updates(batch_idx: Tensor, new_codes: Tensor, group=None, capacity=None):
    """All-gather, in rank order, every rank's (re-assigned node ids [B], their new codes [B, nbc]) ->
    (gidx [world * cap], gcodes [world * cap, nbc]), or None for a single process.

    Ranks sample from their own node ranges, so their batch sizes may differ (last / partial batches, uneven
    N / world_size).  `capacity` (>= every rank's B) fixes the per-rank slot size: shorter batches are padded
    with node id -1, which the apply step skips.  With capacity=None every rank must bring the SAME B; that is
    verified with one small allreduce (skipped under CUDA-graph capture, where shapes are frozen and the eager
    warm-up of the same step has already checked them) instead of letting NCCL hang or mis-align the pairs."""
    rank, ws = world(group)
    if ws == 1:
        return None
    B, nbc = new_codes.shape
    dev = new_codes.device
    if capacity is None:
        capturing = dev.type == 'cuda' and torch.cuda.is_current_stream_capturing()
        if not capturing:
            chk = torch.tensor([B, -B], dtype=torch.int64, device=dev)
            dist.all_reduce(chk, op=dist.ReduceOp.MAX, group=group)
            hi, lo = int(chk[0]), -int(chk[1])
            if hi != lo:
                raise ValueError(f"allgather_code_updates: ranks bring different batch sizes ({lo}..{hi}); set "
                                 f"VQBank.gather_capacity to the largest one")
        cap = B
    else:
        cap = int(capacity)
        if B > cap:
            raise ValueError(f"allgather_code_updates: batch of {B} rows exceeds gather_capacity={cap}")
    if cap != B:
        pidx = torch.full((cap,), -1, dtype=batch_idx.dtype, device=dev)
        pidx[:B] = batch_idx
        pcodes = torch.zeros(cap, nbc, dtype=new_codes.dtype, device=dev)
        pcodes[:B] = new_codes
        batch_idx, new_codes = pidx, pcodes
    gidx = torch.empty(ws * cap, dtype=batch_idx.dtype, device=dev)
    gcodes = torch.empty(ws * cap, nbc, dtype=new_codes.dtype, device=dev)
    # codes travel as raw bytes: int16 is not a collective dtype on every backend
    _timed("nccl_allgather_codes", lambda: (
        dist.all_gather_into_tensor(gidx, batch_idx.contiguous(), group=group),
        dist.all_gather_into_tensor(gcodes.view(torch.uint8), new_codes.contiguous().view(torch.uint8), group=group)))
    return gidx, gcodes


def apply_code_updates_(codes: Tensor, gidx: Tensor, gcodes: Tensor, k0: int = 0) -> None:
    """Plain-torch application of gathered updates, LAST entry wins for a repeated node (so a node shared by two
    ranks' batches resolves identically on every replica).  The CUDA path uses vqgnn_codes_apply_updates; this is
    its device-agnostic restatement (CPU tests, fallback for odd dtypes)."""
    keep = gidx >= 0                       # -1 = padding of a short batch (allgather_code_updates capacity)
    gidx, gcodes = gidx[keep], gcodes[keep]
    n, nbc = gcodes.shape
    owner = torch.full((codes.shape[0],), -1, dtype=torch.long, device=codes.device)
    owner.scatter_reduce_(0, gidx.long(), torch.arange(n, device=codes.device), reduce='amax')
    win = owner[gidx.long()] == torch.arange(n, device=codes.device)
    codes[gidx[win].long(), k0:k0 + nbc] = gcodes[win]


def allgather_code_updates_(codes: Tensor, batch_idx: Tensor, new_codes: Tensor, k0: int = 0, group=None,
                            capacity=None):
    """Gather + apply (see the two functions above).  Returns the gathered node ids or None when single."""
    got = allgather_code_updates(batch_idx, new_codes, group, capacity)
    if got is None:
        return None
    apply_code_updates_(codes, got[0], got[1], k0)
    return got[0]


def replicas_max_abs_diff(t: Tensor, group=None) -> float:
    """max |t_rank - t_rank0| over ranks (0.0 for identical replicas): a cheap divergence check."""
    if world(group)[1] == 1:
        return 0.0
    ref = t.detach().clone()
    dist.broadcast(ref, src=0, group=group)
    d = (t.detach() - ref).abs().max().reshape(1).float()
    dist.all_reduce(d, op=dist.ReduceOp.MAX, group=group)
    return float(d.item())
