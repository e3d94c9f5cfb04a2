"""`VectorQuantizerEMA` — drop-in for vq_gnn_v{1,2}/vq.py:60-279, backed by libvqgnn (sm_100a).

Same constructor, method names, return shapes, buffer names and `state_dict()` keys as the reference.
The storage is a `VQBank`: the codebooks / EMA statistics / BatchNorm running statistics of ALL
`num_branch` quantisers of a layer stacked as `[nb, M, Wp]`, so one kernel launch covers every branch
(the reference runs ~261 ATen ops per branch per call).  A stand-alone
`VectorQuantizerEMA` owns a bank with nb = 1; the per-branch modules inside a `LowRankGNNLayer`
are views into the layer's bank.

Pipeline of one `update` / `feature_update`:
    vqgnn_vq_moments  ->  [allreduce]  ->  vqgnn_vq_whiten  ->  vqgnn_vq_assign (+ per-codeword sums)
                      ->  [allreduce]  ->  vqgnn_vq_finalize (EMA, Laplace smoothing, recovery)
The two bracketed points are where a multi-GPU run sums its statistics over ranks (NCCL), which makes
every rank's codebook replica identical.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import nn

from . import _lib, dist

Tensor = torch.Tensor


def _round4(w: int) -> int:
    return (w + 3) // 4 * 4


class VQBank:
    """Stacked state of `nb` quantisers + the kernel launches.  Not an nn.Module on purpose: the
    registered buffers live on the per-branch `VectorQuantizerEMA` modules (reference key names)
    as views into these tensors."""

    def __init__(self, nb: int, num_embeddings: int, embedding_dim: int, decay: float = 0.99,
                 epsilon: float = 1e-24, grad_normalize_scale: Sequence[float] = (1, 1),
                 warm_up_flag: bool = False, momentum: float = 0.1, add_flag: bool = False,
                 num_N: int = 0):
        self.nb, self.M, self.D = nb, num_embeddings, embedding_dim
        self.add = 1 if add_flag else 0
        self.Dg = embedding_dim + self.add
        self.W = embedding_dim + self.Dg
        self.Wp = _round4(self.W)
        self.Ws = self.Wp + 4
        self.decay, self.eps = decay, epsilon
        self.scale = [float(grad_normalize_scale[0]), float(grad_normalize_scale[1])]
        self.warm_up_flag, self.momentum = warm_up_flag, momentum
        self.num_N = num_N
        self.bn_inited = [False] * nb      # one flag per quantiser (vq.py:100), not per layer
        # True (default): per-codeword statistics by the ordered segmented sum (vqgnn_vq_segsum: bit-stable results);
        # False: float atomics fused into the assignment kernel's epilogue (order-dependent last bits)
        self.deterministic = True
        # multi-GPU: fixed capacity (rows) of the per-rank slot in the code-update all-gather; None = every rank must
        # bring the same batch size, which is then verified with one small allreduce per update
        self.gather_capacity: Optional[int] = None
        self._status_host = None           # (pinned int32, event) of the last asynchronous status read
        # True: the hook's update (which only prepares state for the NEXT step, vq_gnn_v2/models.py:39-56 returns
        # `grad` unchanged) runs on a side stream, overlapping the rest of the backward pass and -- multi-GPU -- taking
        # its three collectives off the critical path; `join()` (called by the layer's next forward, or explicitly
        # through LowRankGNN.join_vq_updates() before a CUDA-graph capture ends) orders later readers after it
        self.async_update = False
        self._pending = False
        self.side_lane = 0            # which VQ side stream this bank's asynchronous updates run on (see side_stream)
        # 0: exact-fp32 SIMT kernel (default: bit-stable codes, the parity anchor), 1: tcgen05 3xTF32 kernel,
        # 'auto': tcgen05 when it is the faster one (M >= 512 and a packed width it supports; measured on B200:
        # 0.55 vs 0.45 ms at M = 256, 0.57 vs 0.89 at M = 1024, 0.62 vs 1.41 at M = 4096)
        self.assign_impl = 0
        self.process_group = None     # set (with world_size > 1) to allreduce statistics over ranks
        self.distributed = False
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt)
        self.E, self.O, self.Wm = z(nb, self.M, self.Wp), z(nb, self.M, self.Wp), z(nb, self.M, self.Wp)
        self.size = z(nb, self.M)
        self.rm_f, self.rv_f = z(nb, self.D), torch.ones(nb, self.D)
        self.rm_g, self.rv_g = z(nb, self.Dg), torch.ones(nb, self.Dg)
        self.nbt_f, self.nbt_g = z(nb, dt=torch.int64), z(nb, dt=torch.int64)
        self.codes = z(max(num_N, 1), nb, dt=torch.int16)   # [N, nb]
        self.status = z(1, dt=torch.int32)
        self.last_idx: Optional[Tensor] = None
        # group-major mirror of the code table for the shared-memory tail kernel (csrc/mp_tail.cu)
        self.G = 0                   # branches per group; resolved on the device (vqgnn_mp_tail_group)
        self.codes_g: Optional[Tensor] = None
        self._codes_g_dirty = True
        self._owner_ws: Optional[Tensor] = None     # scratch of vqgnn_codes_apply_updates (multi-GPU)

    TENSORS = ("E", "O", "Wm", "size", "rm_f", "rv_f", "rm_g", "rv_g", "nbt_f", "nbt_g", "codes", "status")

    @property
    def device(self):
        return self.E.device

    def to(self, device):
        for n in self.TENSORS:
            setattr(self, n, getattr(self, n).to(device))
        self.codes_g, self._codes_g_dirty = None, True
        return self

    def mark_codes_dirty(self):
        """Call after writing `codes` (or a per-branch `c_indices` view) from outside the library."""
        self._codes_g_dirty = True

    def grouped_codes(self) -> Optional[Tensor]:
        """codes_g [ceil(nb/G), N, 8] int16 (None when the shape has no shared-memory path)."""
        lib = _lib.load()
        if self.G == 0:
            self.G = int(lib.vqgnn_mp_tail_group(self.M, self.D, self.Wp))
            if self.G == 0:
                self.G = -1
        if self.G < 0:
            return None
        if self.codes_g is None or self._codes_g_dirty:
            N = self.codes.shape[0]
            ng = (self.nb + self.G - 1) // self.G
            if self.codes_g is None:
                self.codes_g = torch.zeros(ng, N, 8, dtype=torch.int16, device=self.codes.device)
            _lib.check(lib.vqgnn_codes_group(_lib.ptr(self.codes), self.nb, None, N, N, self.G,
                                             _lib.ptr(self.codes_g), _lib.stream()))
            self._codes_g_dirty = False
        return self.codes_g

    # ---- the fused update ---------------------------------------------------------------------
    def run(self, x: Tensor, g: Optional[Tensor], batch_idx: Optional[Tensor], training: bool,
            k0: int = 0, nbc: Optional[int] = None, write_codes: bool = True,
            local: bool = False) -> Tensor:
        """x: [B, >= (k0+nbc)*D] (branch k reads columns k*D..), g: [B, nb*Dg] or None; with
        `local` the operands hold only the nbc selected branches, starting at column 0.
        Returns idx [B, nbc] int16 (codes chosen with the PRE-update codebook)."""
        nbc = self.nb - k0 if nbc is None else nbc
        _lib.require_device(x)
        lib = _lib.load()
        st = _lib.stream()
        D, Dg, M, Wp = self.D, self.Dg, self.M, self.Wp
        joint = g is not None
        assert x.dtype == torch.float32 and x.stride(1) == 1
        B = x.shape[0]
        dev = x.device
        c0 = 0 if local else k0
        xk = x[:, c0 * D:(c0 + nbc) * D]
        gk = None
        if joint:
            assert g.dtype == torch.float32 and g.stride(1) == 1 and g.shape[1] >= (c0 + nbc) * Dg
            gk = g[:, c0 * Dg:(c0 + nbc) * Dg]
        assert xk.shape[1] == nbc * D
        C, Cg = nbc * D, (nbc * Dg if joint else 0)
        sl = slice(k0, k0 + nbc)
        E, O, Wm, size = self.E[sl], self.O[sl], self.Wm[sl], self.size[sl]
        rm_f, rv_f, rm_g, rv_g = self.rm_f[sl], self.rv_f[sl], self.rm_g[sl], self.rv_g[sl]
        nbt_f, nbt_g = self.nbt_f[sl], self.nbt_g[sl]
        scale = torch.empty(C + Cg, device=dev)
        shift = torch.empty(C + Cg, device=dev)
        sums, d_count = None, None
        count = float(B)
        # vq.py:216-221: the FIRST update() of a quantiser seeds both BatchNorms' running statistics from the batch
        # (whatever the mode); feature_update never does
        seeds = [joint and not self.bn_inited[k] for k in range(k0, k0 + nbc)]
        seed, seed_mask = (1 if seeds[0] else 0), None
        if any(seeds) and not all(seeds):
            seed_mask = torch.tensor([int(v) for v in seeds], dtype=torch.int32, device=dev)
        if training or any(seeds):
            sums = torch.empty(2 * (C + Cg) + 1, dtype=torch.float64, device=dev)
            mws = torch.empty(int(lib.vqgnn_vq_moments_workspace_bytes(B, C, Cg)) // 8 + 1, dtype=torch.float64,
                              device=dev)
            _lib.check(lib.vqgnn_vq_moments(_lib.ptr(xk), xk.stride(0), _lib.ptr(gk),
                                            gk.stride(0) if joint else 0, B, C, Cg, _lib.ptr(sums), _lib.ptr(mws), st))
            if self.distributed:   # global batch statistics: every rank whitens identically
                sums[-1:].fill_(float(B))          # a fill kernel (capturable), not a host copy
                dist.allreduce_sum_(sums, self.process_group)
                d_count = sums[-1:]
        _lib.check(lib.vqgnn_vq_whiten(
            _lib.ptr(sums), count, _lib.ptr(d_count), nbc, D, Dg, 1 if joint else 0, _lib.ptr(rm_f), _lib.ptr(rv_f),
            _lib.ptr(rm_g) if joint else None, _lib.ptr(rv_g) if joint else None,
            1e-5, 0.1, self.eps, self.momentum, self.scale[0], self.scale[1], 1 if training else 0, seed,
            _lib.ptr(seed_mask), _lib.ptr(nbt_f) if training else None,
            _lib.ptr(nbt_g) if (training and joint) else None, _lib.ptr(scale), _lib.ptr(shift), st))
        if joint:
            for k in range(k0, k0 + nbc):
                self.bn_inited[k] = True
        idx = torch.empty(B, nbc, dtype=torch.int16, device=dev)
        stats = None
        fused_stats = training and not self.deterministic
        if training:
            stats = torch.empty(nbc, M, self.Ws, device=dev)
            if fused_stats:
                _lib.check(lib.vqgnn_fill_zero(_lib.ptr(stats), stats.numel() * 4, st))
        codes_ptr, bidx = None, None
        if write_codes and batch_idx is not None:
            assert batch_idx.dtype == torch.int32 and batch_idx.is_cuda
            bidx = batch_idx
            codes_ptr = _lib.C.c_void_p(self.codes.data_ptr() + 2 * k0)
        impl = self.assign_impl
        if impl == 'auto':
            impl = 1 if (M >= 512 and D == 4 and (D + (Dg if joint else 0)) in (4, 8, 9)) else 0
        ws, ws_bytes = None, 0
        if impl == 1:     # tcgen05 path: the codebook re-packed into MMA tiles
            ws_bytes = int(lib.vqgnn_vq_assign_workspace_bytes(nbc, M))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.vqgnn_vq_assign(
            _lib.ptr(xk), xk.stride(0), _lib.ptr(gk), gk.stride(0) if joint else 0, _lib.ptr(scale),
            _lib.ptr(shift), _lib.ptr(E), B, nbc, M, D, Dg, Wp, _lib.ptr(bidx), codes_ptr, self.nb,
            _lib.ptr(idx), _lib.ptr(stats) if fused_stats else None, int(impl), _lib.ptr(ws), ws_bytes, st))
        if training:
            if not fused_stats:    # ordered segmented sum over the assignments: no float atomics
                sb = int(lib.vqgnn_vq_segsum_workspace_bytes(B, nbc, M))
                sws = torch.empty(sb, dtype=torch.uint8, device=dev)
                _lib.check(lib.vqgnn_vq_segsum(
                    _lib.ptr(xk), xk.stride(0), _lib.ptr(gk), gk.stride(0) if joint else 0, _lib.ptr(scale),
                    _lib.ptr(shift), _lib.ptr(idx), B, nbc, M, D, Dg, Wp, _lib.ptr(stats), _lib.ptr(sws), sb, st))
            if self.distributed:
                dist.allreduce_sum_(stats, self.process_group)
            _lib.check(lib.vqgnn_vq_finalize(
                _lib.ptr(stats), nbc, M, D, Dg, Wp, 1 if joint else 0, float(self.decay),
                1 if self.warm_up_flag else 0, self.eps, self.scale[0], self.scale[1], _lib.ptr(rm_f),
                _lib.ptr(rv_f), _lib.ptr(rm_g) if joint else None, _lib.ptr(rv_g) if joint else None,
                _lib.ptr(size), _lib.ptr(Wm), _lib.ptr(E), _lib.ptr(O), _lib.ptr(self.status), st))
        if codes_ptr is not None and self.distributed:
            # every replica of the code table learns the other ranks' re-assignments (their batch nodes are this
            # rank's out-of-batch neighbours): one all-gather of (node id, codes) per update, applied by one kernel
            # with last-entry-wins semantics (identical on every rank)
            got = dist.allgather_code_updates(bidx, idx, self.process_group, capacity=self.gather_capacity)
            if got is not None:
                gidx, gcodes = got
                if self._owner_ws is None or self._owner_ws.device != dev:
                    self._owner_ws = torch.empty(self.codes.shape[0], dtype=torch.int32, device=dev)
                mirror = self.codes_g if (self.codes_g is not None and not self._codes_g_dirty) else None
                _lib.check(lib.vqgnn_codes_apply_updates(
                    _lib.ptr(gidx), _lib.ptr(gcodes), int(gidx.numel()), nbc, k0, _lib.ptr(self.codes), self.nb,
                    self.codes.shape[0], _lib.ptr(mirror), max(self.G, 0) if mirror is not None else 0,
                    _lib.ptr(self._owner_ws), st))
        elif codes_ptr is not None and self.codes_g is not None and not self._codes_g_dirty:
            if k0 == 0 and nbc == self.nb:    # keep the group-major mirror in step (re-assigned rows only)
                _lib.check(lib.vqgnn_codes_group(_lib.ptr(self.codes), self.nb, _lib.ptr(bidx), B,
                                                 self.codes.shape[0], self.G, _lib.ptr(self.codes_g), st))
            else:
                self._codes_g_dirty = True
        self.last_idx = idx
        self.last_stats = stats
        return idx

    # ---- the hook's update, optionally off the critical path ----------------------------------------
    _SIDE = {}

    @classmethod
    def side_stream(cls, dev, which: int = 0, lane: int = 0) -> "torch.cuda.Stream":
        """Side streams per device and purpose (0 = VQ updates, 1 = tail-row prefetch).  Lane 0 is shared by every bank:
        the updates (and their collectives) of all layers then stay in program order, which is the same on every rank.
        A bank with its own `side_lane` (LowRankGNN.set_async_vq_updates(per_layer_streams=True)) runs its update
        concurrently with the other layers' -- the latency-bound small kernels of one layer (moments, sort, segmented
        sums, finalize) fill the gaps of another layer's assignment kernel; multi-GPU this needs one communicator per
        lane (`process_group`), NCCL calls on one communicator must not run concurrently."""
        idx = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
        key = (idx, which, lane)
        st = cls._SIDE.get(key)
        if st is None:
            st = cls._SIDE[key] = torch.cuda.Stream(device=dev)
        return st

    def update(self, x: Tensor, g: Tensor, batch_idx: Tensor) -> None:
        """The VQ hook body: vq.update(X_B, grad); c_indices[batch] = idx  (vq_gnn_v2/models.py:39-46)."""
        if not self.async_update:
            self.run(x, g, batch_idx, True)
            return
        cur = torch.cuda.current_stream(x.device)
        side = self.side_stream(x.device, 0, self.side_lane)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self.run(x, g, batch_idx, True)
        for t in (x, g, batch_idx):
            t.record_stream(side)
        self._pending = True

    def join(self) -> None:
        """Order the current stream after a pending asynchronous update of this bank."""
        if self._pending:
            torch.cuda.current_stream(self.E.device).wait_stream(self.side_stream(self.E.device, 0, self.side_lane))
            self._pending = False

    def check_status(self):
        """'Bad Init!' check (vq.py:188-189, 253-254): one host sync."""
        self.join()
        if self.status.is_cuda and int(self.status.item()) & 1:
            self.status.zero_()
            self._status_host = None
            raise ValueError('Bad Init!')

    def poll_status(self):
        """The same check without stalling the stream: reads the status word that was copied to pinned memory after
        an EARLIER update (raising one or two steps late), then queues the next copy.  Called by the layers on every
        forward, so a bad initialisation surfaces by itself; skipped during CUDA-graph capture."""
        if not self.status.is_cuda or torch.cuda.is_current_stream_capturing():
            return
        if self._status_host is not None:
            host, ev = self._status_host
            if not ev.query():
                return
            if int(host[0]) & 1:
                self.status.zero_()
                self._status_host = None
                raise ValueError('Bad Init!')
        else:
            host = torch.zeros(1, dtype=torch.int32).pin_memory()
        host.copy_(self.status, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._status_host = (host, ev)


class VectorQuantizerEMA(nn.Module):
    """Reference signature: vq_gnn_v2/vq.py:61-62."""

    def __init__(self, num_embeddings, embedding_dim, commitment_cost=0.5, decay=0.99, epsilon=1e-24,
                 grad_normalize_scale=(1, 1), warm_up_flag=False, momentum=0.1, add_flag=False,
                 _bank: Optional[VQBank] = None, _branch: int = 0):
        super().__init__()
        if type(grad_normalize_scale) is not list:
            raise ValueError('grad scale type wrong!')          # vq.py:91-92
        self.add_flag = add_flag
        added_dim = 1 if add_flag else 0
        self._embedding_dim, self._num_embeddings = embedding_dim, num_embeddings
        self._commitment_cost, self._warm_up_flag = commitment_cost, warm_up_flag
        self._decay, self._epsilon = decay, epsilon
        self.grad_normalize_scale = grad_normalize_scale
        D, W = embedding_dim, embedding_dim * 2 + added_dim
        # same RNG consumption as the reference constructor (vq.py:73-80)
        emb = torch.randn(num_embeddings, W)
        ema_w = torch.zeros(num_embeddings, W)
        if warm_up_flag:
            ema_w.normal_()
        emb[:, D:2 * D] *= grad_normalize_scale[0]
        ema_w[:, D:2 * D] *= grad_normalize_scale[0]
        if add_flag:
            emb[:, 2 * D] *= grad_normalize_scale[1]
            ema_w[:, 2 * D] *= grad_normalize_scale[1]
        self.register_buffer('_embedding', emb)
        self.register_buffer('_embedding_output', torch.zeros(num_embeddings, W))
        self.register_buffer('_ema_cluster_size', torch.zeros(num_embeddings))
        self.register_buffer('_ema_w', ema_w)
        self.batch_norm_feat = nn.BatchNorm1d(embedding_dim, affine=False)
        self.batch_norm_grad = nn.BatchNorm1d(embedding_dim + added_dim, eps=epsilon, affine=False,
                                              momentum=momentum)
        self._owns_bank = _bank is None
        self._branch = _branch
        self.lazy_status = False
        if _bank is None:
            _bank = VQBank(1, num_embeddings, embedding_dim, decay, epsilon, grad_normalize_scale,
                           warm_up_flag, momentum, add_flag)
        object.__setattr__(self, '_bank', _bank)   # not a submodule / buffer
        if self._owns_bank:
            self._rebind()

    # ---- storage plumbing -------------------------------------------------------------------
    @property
    def bank(self) -> VQBank:
        return self._bank

    @property
    def bn_inited(self) -> bool:
        return self._bank.bn_inited[self._branch]

    @bn_inited.setter
    def bn_inited(self, v: bool):
        self._bank.bn_inited[self._branch] = bool(v)

    def _pull_into_bank(self, bank: VQBank, i: int):
        """Copy this module's (possibly just-moved / just-loaded) buffers into slot i of `bank`."""
        W = bank.W
        if bank is not self._bank:
            bank.bn_inited[i] = self._bank.bn_inited[self._branch]
        bank.E[i, :, :W] = self._embedding
        bank.O[i, :, :W] = self._embedding_output
        bank.Wm[i, :, :W] = self._ema_w
        bank.size[i] = self._ema_cluster_size
        bank.rm_f[i], bank.rv_f[i] = self.batch_norm_feat.running_mean, self.batch_norm_feat.running_var
        bank.rm_g[i], bank.rv_g[i] = self.batch_norm_grad.running_mean, self.batch_norm_grad.running_var
        bank.nbt_f[i] = self.batch_norm_feat.num_batches_tracked
        bank.nbt_g[i] = self.batch_norm_grad.num_batches_tracked

    def _bind_views(self, bank: VQBank, i: int):
        """Re-point the registered buffers at views of slot i of `bank`."""
        W = bank.W
        object.__setattr__(self, '_bank', bank)
        self._branch = i
        self._buffers['_embedding'] = bank.E[i, :, :W]
        self._buffers['_embedding_output'] = bank.O[i, :, :W]
        self._buffers['_ema_w'] = bank.Wm[i, :, :W]
        self._buffers['_ema_cluster_size'] = bank.size[i]
        bf, bg = self.batch_norm_feat, self.batch_norm_grad
        bf._buffers['running_mean'], bf._buffers['running_var'] = bank.rm_f[i], bank.rv_f[i]
        bg._buffers['running_mean'], bg._buffers['running_var'] = bank.rm_g[i], bank.rv_g[i]
        bf._buffers['num_batches_tracked'] = bank.nbt_f[i]
        bg._buffers['num_batches_tracked'] = bank.nbt_g[i]

    def _rebind(self):
        dev = self._embedding.device
        bank = self._bank
        if bank.device != dev:
            bank.to(dev)
        self._pull_into_bank(bank, 0)
        self._bind_views(bank, 0)

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        if self._owns_bank:
            self._rebind()
        return out

    # ---- reference API ----------------------------------------------------------------------
    def feature_kmeans_init(self, kmeans_centroids, kmeans_counts):   # vq.py:102-105
        D = self._embedding_dim
        self._embedding[:, :D] = kmeans_centroids
        self._ema_cluster_size.copy_(kmeans_counts)
        self._ema_w[:, :D] = kmeans_centroids * kmeans_counts.unsqueeze(1)

    def kmeans_init(self, kmeans_centroids, kmeans_counts):           # vq.py:108-118
        D = self._embedding_dim
        self._embedding.copy_(kmeans_centroids)
        self._ema_cluster_size.copy_(kmeans_counts)
        self._ema_w.copy_(kmeans_centroids * kmeans_counts.unsqueeze(1))
        self._embedding[:, D:2 * D] *= self.grad_normalize_scale[0]
        self._ema_w[:, D:2 * D] *= self.grad_normalize_scale[0]
        if self.add_flag:
            self._embedding[:, 2 * D] *= self.grad_normalize_scale[1]
            self._ema_w[:, 2 * D] *= self.grad_normalize_scale[1]

    def get(self):                 # vq.py:120
        return self._embedding_output

    def get_codebook(self):        # vq.py:123
        return self._embedding_output[:, :self._embedding_dim]

    def get_grad(self):            # vq.py:126
        return self._embedding_output[:, self._embedding_dim:]

    def get_feat_cen_norm(self):
        return torch.norm(torch.mean(self._embedding[:, :self._embedding_dim], dim=0)).item()

    def get_grad_cen_norm(self):
        return torch.norm(torch.mean(self._embedding[:, self._embedding_dim:], dim=0)).item()

    def _get_feat_embed(self):
        return self._embedding[:, :self._embedding_dim]

    def feature_update(self, X_B: Tensor) -> Tensor:
        """vq.py:160-202.  Returns encoding_indices [B, 1] int64."""
        x = X_B.detach().contiguous().float()
        idx = self._bank.run(x, None, None, self.training, k0=self._branch, nbc=1, write_codes=False,
                             local=True)
        if self.training and not self.lazy_status:
            self._bank.check_status()
        return idx.to(torch.long)

    def update(self, X_B: Tensor, grad: Tensor):
        """vq.py:204-279.  Returns (encoding_indices [B,1] int64, encodings).  `encodings` (the dense
        one-hot [B, M] no caller uses) is built lazily: a `LazyOneHot` whose
        `.dense()` materialises it."""
        x = X_B.detach().contiguous().float()
        g = grad.detach().contiguous().float()
        bank, i = self._bank, self._branch
        idx = bank.run(x, g, None, self.training, k0=i, nbc=1, write_codes=False, local=True)
        if self.training and not self.lazy_status:
            bank.check_status()
        idx = idx.to(torch.long)
        return idx, LazyOneHot(idx, self._num_embeddings)

    def forward(self, *a, **k):
        raise NotImplementedError("the reference VectorQuantizerEMA defines no forward()")


class LazyOneHot:
    def __init__(self, idx: Tensor, M: int):
        self.idx, self.M = idx, M

    def dense(self) -> Tensor:
        out = torch.zeros(self.idx.shape[0], self.M, device=self.idx.device)
        return out.scatter_(1, self.idx, 1)

    def sum(self, dim=0):
        assert dim == 0
        return torch.bincount(self.idx.view(-1), minlength=self.M).float()
