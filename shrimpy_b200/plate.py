"""Streaming deskew of an OME-Zarr plate: pinned-host chunk loader, three CUDA streams, no collective.

The unit of work is one ``(position, time, channel)`` stack; a rank takes every
``world_size``-th unit (``sharding.shard_units``).  For each unit, on one GPU:

    IO threads   zarr chunks --readinto--> pinned host stack        (``ZarrArray.read_stack_into``: z-chunks land in place)
    stream h2d   pinned host --> device raw slot                     (uint16 as stored; no float32 staging)
    stream run   deskew_tma_kernel (fused convert + lerp + average)  (``shrimpy_deskew_device``)
    stream d2h   device out slot --> pinned host result
    IO threads   pinned result --> output zarr chunks / callback

``depth`` slots of (pinned in, device in, device out, pinned out) rotate, so the load of unit i+2, the
H2D of i+1, the kernel of i and the D2H + write of i-1 overlap.  This is the end-to-end path
BASELINE.json configs[3] describes; it is PCIe/host-bound by construction (the kernel is ~1.5 % of a
unit's time), so the loader's job is to keep both copy engines busy.
"""

from __future__ import annotations

import threading
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .deskew import deskew_geometry, deskew_zyx
from .settings import DeskewSettings
from .sharding import shard_units
from .zarr_io import Position, create_plate

__all__ = ["PlateStats", "list_units", "deskew_plate", "create_deskewed_plate"]


@dataclass
class PlateStats:
    units: int = 0
    seconds: float = 0.0
    raw_voxels: int = 0
    out_voxels: int = 0
    disk_read_bytes: int = 0
    disk_write_bytes: int = 0
    h2d_bytes: int = 0
    d2h_bytes: int = 0
    load_seconds: float = 0.0      # summed over IO threads
    write_seconds: float = 0.0
    launches: int = 0
    per_unit: List[Tuple[str, int, int]] = field(default_factory=list)

    @property
    def gvoxel_out_per_s(self) -> float:
        return self.out_voxels / self.seconds / 1e9 if self.seconds else 0.0

    def as_dict(self) -> dict:
        d = {k: getattr(self, k) for k in ("units", "seconds", "raw_voxels", "out_voxels", "disk_read_bytes",
                                           "disk_write_bytes", "h2d_bytes", "d2h_bytes", "load_seconds",
                                           "write_seconds", "launches")}
        d["gvoxel_out_per_s"] = self.gvoxel_out_per_s
        d["pcie_gbs"] = (self.h2d_bytes + self.d2h_bytes) / self.seconds / 1e9 if self.seconds else 0.0
        return d


def list_units(positions: Sequence[Position]) -> List[Tuple[int, int, int]]:
    """All ``(position index, t, c)`` stacks of a plate, position-major (the acquisition order)."""
    units = []
    for i, pos in enumerate(positions):
        T, C = pos.array.shape[:2]
        units += [(i, t, c) for t in range(T) for c in range(C)]
    return units


def create_deskewed_plate(path, src: Sequence[Position], settings: DeskewSettings, z_chunk: int = 50,
                          zstd_level: Optional[int] = None, blosc: Optional[dict] = None,
                          shard_z: Optional[int] = None) -> List[Position]:
    """Output plate mirroring ``src``: float32, deskewed shape, chunks ``(1, 1, z_chunk, Y', X')`` and the scale
    transform ``(1, 1) + voxel_size`` exactly as ``scripts/measure_psf.py:273-287`` writes it.  ``blosc`` /
    ``shard_z`` (z extent of one shard file, a multiple of ``z_chunk``) select the acquisition's own storage layout
    (``sharding_indexed`` -> ``blosc``; ``shrimpy/mantis/mantis_engine.py:474-481``)."""
    T, C, Z, Y, X = src[0].array.shape
    g = deskew_geometry((Z, Y, X), settings.ls_angle_deg, settings.px_to_scan_ratio, settings.keep_overhang,
                        settings.average_n_slices, settings.pixel_size_um)
    Zd, Yd, Xd = g.out_shape
    zc = min(z_chunk, Zd)
    shard = None if shard_z is None else (1, 1, zc * max(1, -(-min(shard_z, Zd) // zc)), Yd, Xd)
    return create_plate(path, [p.name for p in src], (T, C, Zd, Yd, Xd), shard or (1, 1, zc, Yd, Xd), np.float32,
                        channel_names=src[0].channel_names, scale=(1.0, 1.0) + tuple(float(v) for v in g.voxel_size),
                        zstd_level=zstd_level, blosc=blosc, shard_inner=(1, 1, zc, Yd, Xd) if shard else None)


def deskew_plate(src: Sequence[Position], settings, dst: Optional[Sequence[Position]] = None, *, rank: int = 0,
                 world_size: int = 1, device: Optional[int] = None, depth: int = 3, io_threads: int = 4,
                 cval: float = 0.0, on_result: Optional[Callable] = None,
                 units: Optional[Sequence[Tuple[int, int, int]]] = None) -> PlateStats:
    """Deskew this rank's share of a plate.  ``settings`` is a ``DeskewSettings`` or the ``deskew:`` dict.

    Results go to ``dst`` (positions of an output plate, see ``create_deskewed_plate``) and/or to
    ``on_result(position_name, t, c, array)`` -- the array is a view of a pinned buffer that is recycled
    after the callback returns.
    """
    import torch

    from . import _cabi

    if not torch.cuda.is_available():
        raise RuntimeError("shrimpy_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if not isinstance(settings, DeskewSettings):
        settings = DeskewSettings(**settings)
    if device is None:
        device = torch.cuda.current_device()
    dev = torch.device("cuda", device)
    mine = shard_units(list(units) if units is not None else list_units(src), world_size, rank)
    stats = PlateStats()
    if not mine:
        return stats
    shapes = {tuple(src[i].array.shape[2:]) for i, _, _ in mine}
    dtypes = {src[i].array.dtype for i, _, _ in mine}
    if len(shapes) != 1 or len(dtypes) != 1:
        raise ValueError("all stacks of one run must share shape and dtype")
    zyx, dtype = shapes.pop(), dtypes.pop()
    if dtype not in (np.dtype(np.uint16), np.dtype(np.float32)):
        raise TypeError(f"plate dtype {dtype} is not supported (uint16 or float32)")
    tdtype = torch.uint16 if dtype == np.uint16 else torch.float32
    args = (settings.ls_angle_deg, settings.px_to_scan_ratio, settings.keep_overhang, settings.average_n_slices)
    g = deskew_geometry(zyx, *args)
    depth = max(1, min(depth, len(mine)))

    with torch.cuda.device(dev):
        h_in = [torch.empty(zyx, dtype=tdtype).pin_memory() for _ in range(depth)]
        h_out = [torch.empty(g.out_shape, dtype=torch.float32).pin_memory() for _ in range(depth)]
        d_in = [torch.empty(zyx, dtype=tdtype, device=dev) for _ in range(depth)]
        d_out = [torch.empty(g.out_shape, dtype=torch.float32, device=dev) for _ in range(depth)]
        s_h2d, s_run, s_d2h = (torch.cuda.Stream(device=dev) for _ in range(3))
        free = [threading.Semaphore(1) for _ in range(depth)]      # slot i is free for a new load
        lock = threading.Lock()
        launches0 = _cabi.launch_count()

        def load(k: int):
            slot = k % depth
            free[slot].acquire()
            i, t, c = mine[k]
            t0 = time.perf_counter()
            nbytes = src[i].array.read_stack_into(t, c, h_in[slot].numpy(), pool=io_pool)
            with lock:
                stats.disk_read_bytes += nbytes
                stats.load_seconds += time.perf_counter() - t0
            return slot

        def finish(k: int, slot: int, done):
            try:
                done.synchronize()                      # D2H of this unit landed in h_out[slot]
                i, t, c = mine[k]
                result = h_out[slot].numpy()
                t0 = time.perf_counter()
                written = dst[i].array.write_stack(t, c, result, pool=io_pool) if dst is not None else 0
                if on_result is not None:
                    on_result(src[i].name, t, c, result)
                with lock:
                    stats.disk_write_bytes += written
                    stats.write_seconds += time.perf_counter() - t0
                    stats.per_unit.append((src[i].name, t, c))
            finally:
                free[slot].release()                    # a failed write must not leave load(k + depth) waiting for ever

        def raise_failed(tails):
            """Surface a writer's / callback's exception now instead of after (or instead of) the next blocking load."""
            for f in tails:
                if f.done() and f.exception() is not None:
                    raise f.exception()

        t_start = time.perf_counter()
        # `pool` runs the per-unit stages (blocked loads never exceed `depth`, so 2*depth+1 workers cannot starve the
        # writers); `io_pool` runs the byte-range pieces those stages fan out, and nothing in it ever blocks.
        with ThreadPoolExecutor(max_workers=2 * depth + 1) as pool, ThreadPoolExecutor(max_workers=max(1, io_threads)) as io_pool:
            loads = {k: pool.submit(load, k) for k in range(min(depth, len(mine)))}
            tails = []
            for k in range(len(mine)):
                raise_failed(tails)
                slot = loads.pop(k).result()
                if k + depth < len(mine):
                    loads[k + depth] = pool.submit(load, k + depth)      # blocks in its thread until the slot frees
                with torch.cuda.stream(s_h2d):
                    d_in[slot].copy_(h_in[slot], non_blocking=True)
                    ev_in = torch.cuda.Event()
                    ev_in.record()
                with torch.cuda.stream(s_run):
                    s_run.wait_event(ev_in)
                    deskew_zyx(d_in[slot], *args, cval=cval, out=d_out[slot])
                    ev_run = torch.cuda.Event()
                    ev_run.record()
                with torch.cuda.stream(s_d2h):
                    s_d2h.wait_event(ev_run)
                    h_out[slot].copy_(d_out[slot], non_blocking=True)
                    ev_out = torch.cuda.Event()
                    ev_out.record()
                tails.append(pool.submit(finish, k, slot, ev_out))
            for f in tails:
                f.result()
        torch.cuda.synchronize(dev)
        stats.seconds = time.perf_counter() - t_start
        stats.launches = _cabi.launch_count() - launches0
    stats.units = len(mine)
    stats.raw_voxels = len(mine) * int(np.prod(zyx))
    stats.out_voxels = len(mine) * int(np.prod(g.out_shape))
    stats.h2d_bytes = stats.raw_voxels * dtype.itemsize
    stats.d2h_bytes = stats.out_voxels * 4
    return stats
