"""Stack assembly on the input side of the deskew boundary.

The reference collects a z-stack as ``img.copy()`` per frame into a Python list
(``shrimpy/dynatrack/manager.py:357-384``), ``np.stack``s it (``shrimpy/dynatrack/tracking.py:1041``),
pickles it through an ``mp.Queue`` and finally does a pageable host-to-device copy with a separate
uint16 -> float32 convert kernel (``shrimpy/preprocessing.py:316``): three host copies of 0.7-1 GB and
a synchronous transfer after the last frame has arrived.

``FrameStack`` removes all of that: every frame is copied ONCE, into its slice of a pinned ``(Z, Y, X)``
buffer, and consecutive slices are shipped to the GPU in batches on a side stream *while the
acquisition is still running*.  When the last frame lands only the tail batch is left to copy, and
``finish()`` hands out a uint16 device tensor that ``fast_deskew_zyx`` consumes directly (the convert is
fused into the deskew kernel).  ``slots`` buffers rotate so the next stack can be collected while the
previous one is still being processed.
"""

from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

__all__ = ["FrameStack"]


class FrameStack:
    """Pinned landing buffer for the frames of one z-stack with overlapped host-to-device transfer."""

    def __init__(self, zyx_shape: Sequence[int], dtype=np.uint16, device=None, h2d_batch: int = 32, slots: int = 2):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("FrameStack needs a CUDA device (pinned memory + async copies); there is no CPU path")
        if len(zyx_shape) != 3 or min(zyx_shape) <= 0:
            raise ValueError(f"zyx_shape must be three positive sizes, got {tuple(zyx_shape)}")
        dtype = np.dtype(dtype)
        tmap = {np.dtype(np.uint16): torch.uint16, np.dtype(np.float32): torch.float32}
        if dtype not in tmap:
            raise TypeError(f"FrameStack holds uint16 or float32 frames, got {dtype}")
        self._torch = torch
        self.shape = tuple(int(s) for s in zyx_shape)
        self.dtype = dtype
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.h2d_batch = max(1, int(h2d_batch))
        with torch.cuda.device(self.device):
            self._host = [torch.empty(self.shape, dtype=tmap[dtype]).pin_memory() for _ in range(max(1, slots))]
            self._dev = [torch.empty(self.shape, dtype=tmap[dtype], device=self.device) for _ in range(max(1, slots))]
            self._stream = torch.cuda.Stream(device=self.device)
        self._views = [h.numpy() for h in self._host]
        self._slot = 0
        self._consumed = [None] * len(self._host)     # event: the device buffer of this slot may be overwritten
        self.h2d_bytes = 0
        self._begin()

    # ---- one stack ----------------------------------------------------------------------------------
    def _begin(self) -> None:
        self._have = np.zeros(self.shape[0], dtype=bool)
        self._shipped = 0            # slices [0, _shipped) are already on their way to the GPU
        done = self._consumed[self._slot]
        if done is not None:         # do not overwrite a device buffer a consumer kernel may still read
            self._stream.wait_event(done)

    @property
    def complete(self) -> bool:
        return bool(self._have.all())

    def put(self, z: int, frame: np.ndarray) -> bool:
        """Copy one camera frame into slice ``z`` (any arrival order); returns True when the stack is complete."""
        Z, Y, X = self.shape
        if not 0 <= z < Z:
            raise IndexError(f"slice {z} outside a stack of {Z}")
        frame = np.asarray(frame)
        if frame.shape != (Y, X):
            raise ValueError(f"frame shape {frame.shape} != {(Y, X)}")
        np.copyto(self._views[self._slot][z], frame, casting="same_kind")   # the single host copy
        self._have[z] = True
        self._ship(final=False)
        return self.complete

    def _ship(self, final: bool) -> None:
        """Send the contiguous prefix of arrived slices that has not been shipped yet, in batches."""
        torch = self._torch
        ready = int(np.argmin(self._have)) if not self._have.all() else self.shape[0]   # first missing slice
        while ready - self._shipped >= self.h2d_batch or (final and ready > self._shipped):
            a = self._shipped
            b = ready if final else a + self.h2d_batch
            with torch.cuda.stream(self._stream):
                self._dev[self._slot][a:b].copy_(self._host[self._slot][a:b], non_blocking=True)
            self.h2d_bytes += (b - a) * self.shape[1] * self.shape[2] * self.dtype.itemsize
            self._shipped = b

    def finish(self):
        """Flush the tail and return the ``(Z, Y, X)`` device tensor; the current stream waits for the copies.

        The tensor stays valid until this slot comes round again (``slots`` stacks later).
        """
        if not self.complete:
            missing = np.flatnonzero(~self._have)
            raise RuntimeError(f"stack incomplete: {missing.size} slices missing (first {int(missing[0])})")
        torch = self._torch
        self._ship(final=True)
        ev = torch.cuda.Event()
        ev.record(self._stream)
        torch.cuda.current_stream(self.device).wait_event(ev)
        out = self._dev[self._slot]
        done = torch.cuda.Event()
        self._pending_release = (self._slot, done)
        self._slot = (self._slot + 1) % len(self._host)
        self._begin()
        return out

    def release(self) -> None:
        """Mark the tensor returned by the last ``finish()`` as consumed on the current stream
        (call after the kernels reading it were launched)."""
        slot, done = getattr(self, "_pending_release", (None, None))
        if slot is not None:
            done.record(self._torch.cuda.current_stream(self.device))
            self._consumed[slot] = done
            self._pending_release = (None, None)
