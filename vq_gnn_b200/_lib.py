"""ctypes binding of libvqgnn.so (the C-ABI in include/vqgnn.h).  No torch types cross the boundary:
tensors are passed as raw device pointers + sizes, the stream as a void*.

The product path FAILS LOUDLY when the CUDA library is missing or the device is not sm_100:
there is no CPU or PyTorch fallback for the kernels.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvqgnn.so")

_lib = None
_arch_ok = {}

vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

_SIGNATURES = {
    "vqgnn_abi_version": (C.c_int, []),
    "vqgnn_arch_check": (C.c_int, [i32]),
    "vqgnn_last_error": (C.c_char_p, []),
    "vqgnn_launch_count": (i64, []),
    "vqgnn_vq_moments_workspace_bytes": (C.c_size_t, [i64, i32, i32]),
    "vqgnn_vq_moments": (C.c_int, [vp, i64, vp, i64, i64, i32, i32, vp, vp, vp]),
    "vqgnn_vq_whiten": (C.c_int, [vp, f64, vp, i32, i32, i32, i32, vp, vp, vp, vp, f32, f32, f32, f32, f32, f32,
                                  i32, i32, vp, vp, vp, vp, vp, vp]),
    "vqgnn_vq_segsum_workspace_bytes": (C.c_size_t, [i64, i32, i32]),
    "vqgnn_vq_segsum": (C.c_int, [vp, i64, vp, i64, vp, vp, vp, i64, i32, i32, i32, i32, i32, vp, vp, C.c_size_t,
                                  vp]),
    "vqgnn_vq_assign_workspace_bytes": (C.c_size_t, [i32, i32]),
    "vqgnn_vq_assign": (C.c_int, [vp, i64, vp, i64, vp, vp, vp, i64, i32, i32, i32, i32, i32, vp, vp, i64, vp,
                                  vp, i32, vp, C.c_size_t, vp]),
    "vqgnn_vq_finalize": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, f64, i32, f32, f32, f32, vp, vp, vp, vp,
                                    vp, vp, vp, vp, vp, vp]),
    "vqgnn_mp_workspace_bytes": (C.c_size_t, [i64, i32, i32]),
    "vqgnn_mp_fwd_tail_workspace_bytes": (C.c_size_t, [i64, i32, i64, i32]),
    "vqgnn_mp_num_chunks": (i64, [i64, i32]),
    "vqgnn_mp_chunk_rows": (C.c_int, [vp, i64, i64, i32, vp, vp]),
    "vqgnn_mp_fwd": (C.c_int, [vp, vp, vp, vp, vp, i32, i64, i64, i64, vp, i64, vp, vp, vp, i32, i32, i32, i32,
                               vp, i64, i32, f32, f32, vp, i64, vp, i64, vp, vp, vp]),
    "vqgnn_tail_materialize": (C.c_int, [vp, i64, vp, vp, i32, i32, i32, i32, vp, vp, i64, vp]),
    "vqgnn_gat_fwd_rows": (C.c_int, [vp, vp, vp, vp, i32, i64, i64, i64, vp, i64, vp, i64, vp, i64, i32, vp, vp, vp, f32,
                                     f32, vp, i64, vp, vp, vp, C.c_size_t, vp]),
    "vqgnn_mp_fwd_rows": (C.c_int, [vp, vp, vp, vp, i32, i64, i64, i64, vp, i64, vp, i64, f32, vp, vp, i64, i32, f32,
                                    vp, i64, vp, vp, vp]),
    "vqgnn_mp_bwd": (C.c_int, [vp, vp, vp, vp, i32, i64, i64, vp, i64, vp, vp, vp, i32, i32, i32, i32, vp, i64,
                               i32, f32, vp, i64, f32, vp, vp, i64, vp, vp]),
    "vqgnn_mp_tail_group": (C.c_int, [i32, i32, i32]),
    "vqgnn_codes_apply_updates": (C.c_int, [vp, vp, i64, i32, i32, vp, i32, i64, vp, i32, vp, vp]),
    "vqgnn_codes_group": (C.c_int, [vp, i32, vp, i64, i64, i32, vp, vp]),
    "vqgnn_mp_fwd_tail": (C.c_int, [vp, vp, vp, vp, vp, i32, i64, vp, i64, vp, i64, vp, i64, vp, i32, i32, i32, i32,
                                    f32, f32, vp, i64, vp, i64, vp, vp, vp]),
    "vqgnn_csr_transpose_workspace_bytes": (C.c_size_t, [i64, i64]),
    "vqgnn_csr_transpose_lt": (C.c_int, [vp, vp, vp, i64, i64, i64, vp, vp, vp, vp, vp, vp]),
    "vqgnn_plan_v1_workspace_bytes": (C.c_size_t, [i64, i64, i64, i64]),
    "vqgnn_plan_v1_build": (C.c_int, [vp, vp, vp, vp, i64, vp, vp, vp, i64, vp, vp, i64, i64, i32, i32, i32, i32,
                                      vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "vqgnn_tail_materialize_slab": (C.c_int, [vp, i64, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    "vqgnn_mp_info_workspace_bytes": (C.c_size_t, [i64, i32, i32]),
    "vqgnn_csr_expand_rows": (C.c_int, [vp, i64, i64, vp, vp]),
    "vqgnn_mp_info": (C.c_int, [vp, vp, vp, i64, i64, i64, i64, vp, i64, vp, vp, i32, i32, f32, vp, vp, vp]),
    "vqgnn_khop_workspace_bytes": (C.c_size_t, [i64, i64]),
    "vqgnn_khop_mark": (C.c_int, [vp, vp, vp, i64, i64, vp, vp, vp, vp, vp]),
    "vqgnn_khop_count": (C.c_int, [vp, vp, vp, vp, i64, i64, i64, vp, vp, vp, vp]),
    "vqgnn_khop_fill": (C.c_int, [vp, vp, vp, vp, vp, i64, i64, i64, vp, vp, vp, vp, vp]),
    "vqgnn_collate_v1_count": (C.c_int, [vp, vp, vp, i64, i64, i32, vp, vp, vp, vp, vp]),
    "vqgnn_collate_v1_fill": (C.c_int, [vp, vp, vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                        vp]),
    "vqgnn_gat_scores": (C.c_int, [i64, i64, vp, i64, vp, vp, vp, i32, i32, i32, i32, vp, i64, vp, vp, vp, vp, vp,
                                   vp]),
    "vqgnn_gat_fwd": (C.c_int, [vp, vp, vp, vp, i32, i64, i64, i64, vp, i64, vp, vp, vp, i32, i32, i32, i32,
                                vp, i64, vp, vp, vp, f32, f32, vp, i64, vp, vp, vp, vp]),
    "vqgnn_gat_bwd": (C.c_int, [vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, i64, i32, i64, vp, i64, vp, vp, vp,
                                i32, i32, i32, i32, vp, vp, i64, vp, vp, vp, vp, vp, f32, vp, i64, vp, vp, i64,
                                f32, vp, vp, i64, vp, vp, vp, vp, i64, vp, vp, vp]),
    "vqgnn_gat1_scores": (C.c_int, [i64, vp, i64, vp, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp, vp]),
    "vqgnn_gat1_fwd": (C.c_int, [vp, vp, vp, vp, vp, i32, i64, i64, vp, i64, vp, vp, i32, i32, i32, i32, f32,
                                 vp, vp, vp, vp, f32, vp, i64, vp, vp, vp, vp]),
    "vqgnn_gat1_bwd": (C.c_int, [vp, vp, vp, vp, vp, i32, i64, i64, vp, i64, vp, vp, i32, i32, i32, i32, f32,
                                 vp, vp, vp, vp, vp, vp, f32, vp, i64, vp, vp, i64, vp, vp, vp, vp, vp, i64,
                                 vp, vp, vp]),
    "vqgnn_fill_zero": (C.c_int, [vp, C.c_size_t, vp]),
    "vqgnn_flush_l2": (C.c_int, [vp, C.c_size_t, vp]),
    "vqgnn_codes_pack": (C.c_int, [vp, i32, i64, vp, vp]),
}


class VQGNNLibraryError(RuntimeError):
    pass


class _Profiler:
    """Optional per-kernel CUDA-event timing (bench.py's attribution pass).  Off by default."""

    def __init__(self):
        self.enabled = False
        self.records = []      # (name, start_event, stop_event, meta)

    def reset(self):
        self.records = []

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b, _ in self.records:
            d = out.setdefault(name, [0, 0.0])
            d[0] += 1
            d[1] += a.elapsed_time(b)
        return out


PROFILER = _Profiler()


class _Proxy:
    """Attribute access returns the ctypes function, wrapped with CUDA events while profiling."""

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not PROFILER.enabled or not name.startswith("vqgnn_") or name in _NO_STREAM:
            return fn

        def timed(*args):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = fn(*args)
            b.record()
            PROFILER.records.append((name, a, b, None))
            return rc
        return timed


_NO_STREAM = {"vqgnn_abi_version", "vqgnn_arch_check", "vqgnn_last_error", "vqgnn_launch_count",
              "vqgnn_mp_workspace_bytes", "vqgnn_mp_fwd_tail_workspace_bytes", "vqgnn_mp_num_chunks", "vqgnn_vq_assign_workspace_bytes", "vqgnn_mp_tail_group",
              "vqgnn_plan_v1_workspace_bytes", "vqgnn_csr_transpose_workspace_bytes",
              "vqgnn_vq_moments_workspace_bytes", "vqgnn_vq_segsum_workspace_bytes", "vqgnn_khop_workspace_bytes",
              "vqgnn_mp_info_workspace_bytes"}


def load():
    """dlopen libvqgnn.so (built in-tree by `make` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise VQGNNLibraryError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `make` (or "
            f"`python -c 'import __graft_entry__ as g; g.build()'`) in the repo root. "
            f"vq_gnn_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here == header/library mismatch: fail loudly
        fn.restype, fn.argtypes = res, args
    if lib.vqgnn_abi_version() != 3:
        raise VQGNNLibraryError("libvqgnn.so ABI version mismatch")
    _lib = _Proxy(lib)
    return _lib


def last_error() -> str:
    return load().vqgnn_last_error().decode("utf-8", "replace")


def check(rc: int):
    if rc == 0:
        return
    msg = last_error()
    if rc == -2:
        raise ValueError(f"libvqgnn: {msg}")
    raise VQGNNLibraryError(f"libvqgnn error {rc}: {msg}")


def require_device(t: torch.Tensor):
    """Every product entry point calls this: CUDA tensor on an sm_100 device, or an exception."""
    if not t.is_cuda:
        raise VQGNNLibraryError(
            "vq_gnn_b200 kernels need CUDA tensors on a B200 (sm_100a); there is no CPU fallback")
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if dev not in _arch_ok:
        check(load().vqgnn_arch_check(dev))
        _arch_ok[dev] = True


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(load().vqgnn_launch_count())
