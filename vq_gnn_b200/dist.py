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

from typing import Iterable, Optional, Tuple

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


def allgather_code_updates_(codes: Tensor, batch_idx: Tensor, new_codes: Tensor, k0: int = 0, group=None):
    """Apply every OTHER rank's code-table updates to this rank's replica.

    codes [N, nb] int16 is the local replica (already holding this rank's own update), batch_idx [B] the global
    node ids this rank just re-assigned, new_codes [B, nbc] their codes for branches [k0, k0 + nbc).  Ranks are
    applied in rank order, so when two ranks' batches share a node every replica ends with the same (highest
    rank's) code.  Returns the gathered node ids [world * B] (for the group-major mirror), or None when single."""
    rank, ws = world(group)
    if ws == 1:
        return None
    B, nbc = new_codes.shape
    gidx = torch.empty(ws * B, dtype=batch_idx.dtype, device=batch_idx.device)
    gcodes = torch.empty(ws * B, nbc, dtype=new_codes.dtype, device=new_codes.device)
    # codes travel as raw bytes: int16 is not a collective dtype on every backend
    _timed("nccl_allgather_codes", lambda: (
        dist.all_gather_into_tensor(gidx, batch_idx.contiguous(), group=group),
        dist.all_gather_into_tensor(gcodes.view(torch.uint8), new_codes.contiguous().view(torch.uint8), group=group)))
    for r in range(ws):      # rank order => identical result on every replica (own slice included on purpose)
        sl = slice(r * B, (r + 1) * B)
        codes[gidx[sl].long(), k0:k0 + nbc] = gcodes[sl]
    return gidx


def replicas_max_abs_diff(t: Tensor, group=None) -> float:
    """max |t_rank - t_rank0| over ranks (0.0 for identical replicas): a cheap divergence check."""
    if world(group)[1] == 1:
        return 0.0
    ref = t.detach().clone()
    dist.broadcast(ref, src=0, group=group)
    d = (t.detach() - ref).abs().max().reshape(1).float()
    dist.all_reduce(d, op=dist.ReduceOp.MAX, group=group)
    return float(d.item())
