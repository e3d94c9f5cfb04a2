"""Host -> device batch pipeline: uploads (and optionally prepares) mini-batch i+1 on a side CUDA stream while
mini-batch i trains, so the PCIe copy of the reference-format batch (int64 COO triples, ~130 MB per step at the
Reddit-shaped config; vq_gnn_v1/main_node.py:27-41 `prepare`, vq_gnn_v2/utils/misc.py:57-75) and the
batch-plan construction overlap the kernels instead of serialising with them."""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch


def _walk(obj, fn):
    if obj is None:
        return None
    if isinstance(obj, torch.Tensor):
        return fn(obj)
    if isinstance(obj, tuple):
        return tuple(_walk(o, fn) for o in obj)
    if isinstance(obj, list):
        return [_walk(o, fn) for o in obj]
    if isinstance(obj, dict):
        return {k: _walk(v, fn) for k, v in obj.items()}
    if hasattr(obj, "tensors") and callable(obj.tensors):    # graph.BatchPlan
        for t in obj.tensors():
            fn(t)
        return obj
    if hasattr(obj, "csr") and hasattr(obj, "sparse_sizes"):   # CSRAdj / SparseTensor-like (v2 batches)
        from .graph import CSRAdj
        rowptr, col, val = obj.csr()
        return CSRAdj(fn(rowptr), fn(col), fn(val), obj.sparse_sizes())
    return obj


class DevicePrefetcher:
    """Iterates `host_batches` (pinned CPU tensors in nested tuples) cyclically, returning device copies.

    Batch i+1 is uploaded -- and run through `prepare(batch_on_device)` (e.g. `(x, model.prepare(batch_A), y)`) --
    on a side CUDA stream while batch i trains.  Two modes:
      threaded=False (default): everything is enqueued from the calling thread right after batch i is handed out.
          Right when `prepare` does not synchronise with the device (v1 plans: csrc/plan.cu builds them on the
          device), because then it costs only launch time and there is no second Python thread to fight for the GIL.
      threaded=True: a worker thread does the upload and `prepare`; use it when `prepare` contains host syncs
          (torch boolean-mask / sort based plan builders), so those do not stall the training thread.
    Tensors handed out are registered with the consumer stream (`record_stream`), so the caching allocator
    does not recycle them while the consumer still reads them."""

    def __init__(self, host_batches: Sequence, device, prepare: Optional[Callable] = None,
                 count: Optional[int] = None, stream: Optional[torch.cuda.Stream] = None, threaded: bool = False):
        import queue
        import threading
        self.host, self.device, self.prepare = host_batches, device, prepare
        # reuse one side stream across epochs: the caching allocator keeps a pool per stream
        self.stream = stream if stream is not None else torch.cuda.Stream(device=device)
        self.count = count
        self.threaded = threaded
        self._i = 0
        self._next = None
        self._stop = False
        self._err = None
        if threaded:
            self._q = queue.Queue(maxsize=1)
            self._thread = threading.Thread(target=self._work, daemon=True)
            self._thread.start()

    def _upload(self, i: int):
        with torch.cuda.stream(self.stream):
            dev = self.device
            b = _walk(self.host[i % len(self.host)], lambda t: t.to(dev, non_blocking=True))
            if self.prepare is not None:
                b = self.prepare(b)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return b, ev

    def _work(self):
        import queue
        try:
            torch.cuda.set_device(self.device)
            i = 0
            while not self._stop and (self.count is None or i < self.count):
                item = self._upload(i)
                while not self._stop:
                    try:
                        self._q.put(item, timeout=0.05)
                        break
                    except queue.Full:
                        pass
                i += 1
        except BaseException as e:   # surfaced to the consumer by next()
            self._err = e
            self._q.put((None, None))

    def next(self):
        if self.threaded:
            b, ev = self._q.get()
            if b is None:
                raise RuntimeError("DevicePrefetcher worker failed") from self._err
        else:
            if self._next is None:
                self._next = self._upload(self._i)
            b, ev = self._next
            self._i += 1
            more = self.count is None or self._i < self.count
            self._next = self._upload(self._i) if more else None    # overlaps with the caller's next launches
        main = torch.cuda.current_stream(self.device)
        main.wait_event(ev)
        _walk(b, lambda t: (t.record_stream(main), t)[1] if t.is_cuda else t)
        return b

    def drain(self):
        self._stop = True
        self._next = None
        if self.threaded:
            try:
                while True:
                    self._q.get_nowait()
            except Exception:
                pass
            self._thread.join(timeout=5)
