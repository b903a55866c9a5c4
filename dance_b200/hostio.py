"""Host ↔ device staging for the numpy-in / numpy-out model handlers.

The reference handlers take and return numpy arrays (``feature_AE_handler`` scgnn2.py:275-335 returns ``X_embed`` and the
full ``X_recon`` — 8 GB at 1 M × 2 000).  At that size the PCIe copies are a visible part of what a ``.fit()`` user waits for,
so they are pipelined against the training kernels instead of bracketing them:

  * uploads go batch by batch on a side stream straight into the batch's rows of the device-resident matrix (pinned source:
    one async copy; pageable source: staged through a small pinned ring);
  * downloads of per-batch results go on a second side stream into pinned host buffers while the next batch trains.

``IOPool`` keeps the pinned buffers alive between handler calls (``cudaHostAlloc`` of 8 GB costs seconds): arrays returned
from a pool are views that stay valid until the next request for the same tag — exactly the lifetime the EM loop of
``ScGNN2.fit`` needs.  Without a pool every call allocates fresh pinned memory.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch


class IOPool:
    """Reusable pinned host buffers keyed by tag."""

    def __init__(self):
        self._bufs: Dict[str, torch.Tensor] = {}

    def pinned(self, tag: str, shape: Tuple[int, ...], dtype=torch.float32) -> torch.Tensor:
        numel = int(np.prod(shape)) if len(shape) else 1
        buf = self._bufs.get(tag)
        if buf is None or buf.dtype != dtype or buf.numel() < numel:
            buf = torch.empty(max(numel, 1), dtype=dtype, pin_memory=True)
            self._bufs[tag] = buf
        return buf[:numel].view(*shape)

    def release(self):
        self._bufs.clear()


def pinned_empty(pool: Optional[IOPool], tag: str, shape, dtype=torch.float32) -> torch.Tensor:
    if pool is not None:
        return pool.pinned(tag, tuple(shape), dtype)
    return torch.empty(tuple(shape), dtype=dtype, pin_memory=True)


def as_host_tensor(a) -> torch.Tensor:
    """numpy / torch-CPU array → contiguous float32 CPU tensor sharing memory when possible."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class Uploader:
    """Row-range uploads of a host matrix on a side stream, each range signalled by an event the compute stream waits on."""

    RING = 3

    def __init__(self, host: torch.Tensor, device, pool: Optional[IOPool] = None, max_rows: int = 0):
        self.host, self.device = host, device
        self.stream = torch.cuda.Stream(device=device)
        self.pinned = host.is_pinned()
        self.ring = None
        if not self.pinned:
            rows = max(1, max_rows)
            self.ring = [pinned_empty(pool, f"upload_ring{i}", (rows, host.shape[1])) for i in range(self.RING)]
            self.ring_free = [None] * self.RING
            self._slot = 0

    def copy_rows(self, b0: int, b1: int, dst: torch.Tensor) -> torch.cuda.Event:
        """Queue host[b0:b1] → dst (device rows); returns the event that marks its completion."""
        if self.pinned:
            src = self.host[b0:b1]
        else:
            s = self._slot
            self._slot = (s + 1) % self.RING
            if self.ring_free[s] is not None:
                self.ring_free[s].synchronize()            # the previous H2D out of this slot has finished
            src = self.ring[s][:b1 - b0]
            src.copy_(self.host[b0:b1])                    # multi-threaded host memcpy into pinned memory
        with torch.cuda.stream(self.stream):
            dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        if not self.pinned:
            self.ring_free[s] = ev
        return ev


class Downloader:
    """Device → pinned-host copies on a side stream, ordered after the producing work of the compute stream."""

    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.device = device

    def copy(self, src: torch.Tensor, dst_host: torch.Tensor) -> torch.cuda.Event:
        """Queue src → dst_host after everything already submitted to the current stream; returns the completion event
        (the compute stream must wait on it before overwriting `src`)."""
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            dst_host.copy_(src, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        return done

    def synchronize(self):
        self.stream.synchronize()
