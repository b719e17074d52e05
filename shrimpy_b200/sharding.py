"""Multi-GPU layout of the deskew path: one process per GPU, no data-path collective where the
work shards naturally, one halo exchange where it does not.

* ``shard_units``      -- (position, time, channel) volumes are independent
  (``docs/data_structure.md:64-65`` describes this per-position parallelism upstream): static
  round-robin over ranks, nothing is exchanged.
* ``plan_scan_split`` / ``deskew_scan_split`` -- ONE oversized volume (BASELINE configs[4]) is cut
  along the scan axis: rank g owns a contiguous range of raw scan slices and computes a contiguous
  range of output columns (o2).  Because of the shear, its columns also read up to
  ``r cos(theta) (Y-1) + 1`` slices below its own range (owned by the previous rank(s)) and at most
  two above; those halos travel peer-to-peer (``torch.distributed`` send/recv: NCCL over
  NVLink/NVSwitch on GPUs, gloo in the CPU tests).  Output columns whose taps lie entirely in the
  rank's own slices are computed while the halo is in flight.

The arithmetic is always the full-stack geometry (window calls), so the concatenated result is
bit-identical to a single-GPU deskew.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .deskew import DeskewGeometry, window_needs

__all__ = ["shard_units", "ScanShard", "plan_scan_split", "exchange_halos", "deskew_scan_split"]


def shard_units(units: Sequence, world_size: int, rank: int) -> list:
    """Static round-robin of independent work units ((position, time, channel) volumes) over ranks."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return [u for i, u in enumerate(units) if i % world_size == rank]


@dataclass(frozen=True)
class ScanShard:
    """What one rank owns, computes and needs in a scan-axis split."""

    rank: int
    own_z: Tuple[int, int]        # raw scan slices this rank holds before the exchange  [z0, z1)
    cols: Tuple[int, int]         # output columns (o2) this rank computes               [c0, c1)
    need_z: Tuple[int, int]       # raw scan slices its columns read                      [z0, z1)
    interior_cols: Tuple[int, int]  # sub-range of cols whose taps lie inside own_z       [c0, c1) (may be empty)

    @property
    def halo_below(self) -> Tuple[int, int]:
        return (self.need_z[0], max(self.need_z[0], min(self.own_z[0], self.need_z[1])))

    @property
    def halo_above(self) -> Tuple[int, int]:
        return (min(self.need_z[1], max(self.own_z[1], self.need_z[0])), self.need_z[1])


def plan_scan_split(g: DeskewGeometry, world_size: int, align: int = 32) -> List[ScanShard]:
    """Cut the output columns into ``world_size`` contiguous ranges (multiples of ``align`` columns) and
    derive the raw slices each range reads.  Raw ownership boundaries follow the column boundaries
    (the scan index of a range's first column at tilt row 0), so the halo is almost entirely below."""
    Z, Y, _ = g.raw_shape
    Yn, _, Xp = g.out_shape
    if world_size < 1:
        raise ValueError("world_size must be positive")
    edges = [min(Xp, int(round(Xp * i / world_size / align)) * align) for i in range(world_size)] + [Xp]
    for i in range(1, len(edges)):
        edges[i] = max(edges[i], edges[i - 1])
    # ownership boundary of rank i: scan slice of its first column at tilt row o0 = 0, clipped to [0, Z]
    own_edges = [0]
    for i in range(1, world_size):
        z = (g.shift + 0.0 * g.m00) + edges[i] * g.m02
        own_edges.append(int(min(max(np.floor(z), own_edges[-1]), Z)))
    own_edges.append(Z)
    shards = []
    for i in range(world_size):
        c0, c1 = edges[i], edges[i + 1]
        own = (own_edges[i], own_edges[i + 1])
        if c1 > c0:
            _, need = window_needs(g, 0, Yn, c0, c1 - c0)
        else:
            need = (own[0], own[0])
        if need[1] <= need[0]:
            need = (own[0], own[0])
        # widest run of columns, starting anywhere in [c0, c1), that needs nothing outside own_z
        lo, hi = c0, c0
        if c1 > c0 and own[1] > own[0]:
            cols = np.arange(c0, c1)
            zmin = (g.shift + (Y - 1) * g.m00) + cols * g.m02       # smallest scan coordinate of a column
            zmax = (g.shift + 0.0 * g.m00) + cols * g.m02           # largest
            inside_vol = (zmax >= 0) & (zmin <= Z - 1)
            lo_need = np.floor(np.clip(zmin, 0, Z - 1))
            hi_need = np.minimum(np.floor(np.clip(zmax, 0, Z - 1)) + 1, Z - 1)
            ok = (~inside_vol) | ((lo_need >= own[0]) & (hi_need < own[1]))
            if ok.any():
                # columns are monotone in z, so the admissible set is one run
                first = int(np.argmax(ok))
                last = first
                while last + 1 < ok.size and ok[last + 1]:
                    last += 1
                lo, hi = c0 + first, c0 + last + 1
        shards.append(ScanShard(i, own, (c0, c1), (int(need[0]), int(need[1])), (lo, hi)))
    return shards


def exchange_halos(own_slab, shards: Sequence[ScanShard], rank: int, *, group=None):
    """Assemble this rank's needed slab ``raw[need_z[0]:need_z[1]]`` from its own slices and its peers'.

    ``own_slab`` is a torch tensor ``(own_z1-own_z0, Y, X)`` (CPU with gloo, CUDA with NCCL).  Every
    rank posts the sends its peers need and the receives it needs (``batch_isend_irecv``), so the
    pattern is a sparse neighbour exchange, not a collective.  Returns ``(slab, requests)`` -- the
    caller may compute interior columns from ``own_slab`` before waiting on ``requests``.
    """
    import torch
    import torch.distributed as dist

    me = shards[rank]
    z0, z1 = me.need_z
    slab = torch.empty((max(z1 - z0, 0),) + tuple(own_slab.shape[1:]), dtype=own_slab.dtype, device=own_slab.device)
    # own part
    a, b = max(z0, me.own_z[0]), min(z1, me.own_z[1])
    if b > a:
        slab[a - z0:b - z0].copy_(own_slab[a - me.own_z[0]:b - me.own_z[0]])
    ops = []
    for other in shards:
        if other.rank == rank:
            continue
        # what I need from `other`
        a, b = max(z0, other.own_z[0]), min(z1, other.own_z[1])
        if b > a:
            ops.append(dist.P2POp(dist.irecv, slab[a - z0:b - z0], other.rank, group=group))
        # what `other` needs from me
        a, b = max(other.need_z[0], me.own_z[0]), min(other.need_z[1], me.own_z[1])
        if b > a:
            ops.append(dist.P2POp(dist.isend, own_slab[a - me.own_z[0]:b - me.own_z[0]].contiguous(), other.rank,
                                  group=group))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    return slab, reqs


def deskew_scan_split(own_slab, g: DeskewGeometry, shards: Sequence[ScanShard], rank: int, *, cval: float = 0.0,
                      group=None, window_fn: Optional[Callable] = None):
    """Deskew this rank's output columns of one volume that is split along the scan axis.

    Returns the compact tensor ``out[:, :, c0:c1]``.  ``window_fn(slab, g, p_begin, p_count, c_begin,
    c_count, y_origin, z_origin, cval)`` defaults to the CUDA window kernel; the CPU tests inject a
    stand-in.  Interior columns are computed from the rank's own slices while the halo travels.
    """
    import torch

    if window_fn is None:
        from .deskew import deskew_window

        def window_fn(slab, g, p_begin, p_count, c_begin, c_count, y_origin, z_origin, cval):
            return deskew_window(slab, g, p_begin=p_begin, p_count=p_count, c_begin=c_begin, c_count=c_count,
                                 y_origin=y_origin, z_origin=z_origin, cval=cval)

    me = shards[rank]
    Yn, X, _ = g.out_shape
    c0, c1 = me.cols
    out = torch.empty((Yn, X, c1 - c0), dtype=torch.float32, device=own_slab.device)
    slab, reqs = exchange_halos(own_slab, shards, rank, group=group)
    if slab.shape[0] == 0 and own_slab.shape[0] == 0:   # nothing of the volume maps into these columns
        for r in reqs:
            r.wait()
        return out.fill_(cval)
    i0, i1 = me.interior_cols
    if i1 > i0:      # overlaps with the exchange
        out[:, :, i0 - c0:i1 - c0] = window_fn(own_slab, g, 0, Yn, i0, i1 - i0, 0, me.own_z[0], cval)
    for r in reqs:
        r.wait()
    for a, b in ((c0, i0 if i1 > i0 else c1), (i1, c1) if i1 > i0 else (c1, c1)):
        if b > a:
            z_origin = me.need_z[0]
            src = slab if slab.shape[0] else own_slab
            if slab.shape[0] == 0:
                z_origin = me.own_z[0]
            out[:, :, a - c0:b - c0] = window_fn(src, g, 0, Yn, a, b - a, 0, z_origin, cval)
    return out
