"""Multi-GPU layout of the deskew path: one process per GPU, no data-path collective where the
work shards naturally, one halo exchange where it does not.

* ``shard_units``      -- (position, time, channel) volumes are independent
  (``docs/data_structure.md:64-65`` describes this per-position parallelism upstream): static
  round-robin over ranks, nothing is exchanged.
* ``plan_scan_split`` / ``deskew_scan_split`` -- ONE oversized volume (BASELINE configs[4]) is cut
  along the scan axis: rank g owns a contiguous range of raw scan slices and computes a contiguous
  range of output columns (o2).  Because of the shear, its columns also read up to
  ``r cos(theta) (Y-1) + 1`` slices below its own range (owned by the previous rank(s)) and at most
  two above; those halos travel peer-to-peer: PULLED from the neighbour's HBM by a device copy over
  peer-mapped (symmetric) memory when the slab is a ``PeerSlab`` (NVLink / NVSwitch loads, no NCCL
  call on the data path), else ``torch.distributed`` send/recv (NCCL on GPUs, gloo in the CPU tests).  Output columns whose taps lie entirely in the
  rank's own slices are computed while the halo is in flight.

The arithmetic is always the full-stack geometry (window calls), so the concatenated result is
bit-identical to a single-GPU deskew.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .deskew import DeskewGeometry, window_needs

__all__ = ["shard_units", "ScanShard", "plan_scan_split", "exchange_halos", "deskew_scan_split", "PeerSlab",
           "exchange_halos_peer"]


def shard_units(units: Sequence, world_size: int, rank: int) -> list:
    """Static round-robin of independent work units ((position, time, channel) volumes) over ranks."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return [u for i, u in enumerate(units) if i % world_size == rank]


@dataclass(frozen=True)
class ScanShard:
    """What one rank owns, computes and needs in a scan-axis split."""

    rank: int
    own_z: Tuple[int, int]          # raw scan slices this rank holds before the exchange  [z0, z1)
    cols: Tuple[int, int]           # output columns (o2) this rank computes               [c0, c1)
    need_z: Tuple[int, int]         # raw scan slices its columns read                      [z0, z1)
    interior_cols: Tuple[int, int]  # sub-range of cols whose taps lie inside own_z         [c0, c1) (may be empty)
    low_need_z: Tuple[int, int]     # slices read by the columns below the interior run  (cols[0] .. interior[0])
    high_need_z: Tuple[int, int]    # slices read by the columns above the interior run  (interior[1] .. cols[1])

    @property
    def low_cols(self) -> Tuple[int, int]:
        return (self.cols[0], self.interior_cols[0])

    @property
    def high_cols(self) -> Tuple[int, int]:
        return (self.interior_cols[1], self.cols[1])

    @property
    def halo_below(self) -> Tuple[int, int]:
        return (self.need_z[0], max(self.need_z[0], min(self.own_z[0], self.need_z[1])))

    @property
    def halo_above(self) -> Tuple[int, int]:
        return (min(self.need_z[1], max(self.own_z[1], self.need_z[0])), self.need_z[1])


def _needs(g: DeskewGeometry, c0: int, c1: int) -> Tuple[int, int]:
    if c1 <= c0:
        return (0, 0)
    _, need = window_needs(g, 0, g.out_shape[0], c0, c1 - c0)
    return (int(need[0]), int(need[1])) if need[1] > need[0] else (0, 0)


def plan_scan_split(g: DeskewGeometry, world_size: int, align: int = 32) -> List[ScanShard]:
    """Cut the output columns into ``world_size`` contiguous ranges (multiples of ``align`` columns) and
    derive the raw slices each range reads.  Raw ownership boundaries follow the column boundaries
    (the scan index of a range's first column at tilt row 0), so the halo is almost entirely below."""
    Z, Y, _ = g.raw_shape
    Xp = g.out_shape[2]
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
        need = _needs(g, c0, c1)
        if need == (0, 0):
            need = (own[0], own[0])
        # the run of columns that needs nothing outside own_z (columns are monotone in z: one run)
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
                first = int(np.argmax(ok))
                last = first
                while last + 1 < ok.size and ok[last + 1]:
                    last += 1
                lo, hi = c0 + first, c0 + last + 1
        shards.append(ScanShard(i, own, (c0, c1), need, (lo, hi), _needs(g, c0, lo), _needs(g, hi, c1)))
    return shards


def exchange_halos(own_slab, shards: Sequence[ScanShard], rank: int, *, group=None):
    """Start the neighbour exchange for this rank's two boundary column ranges.

    ``own_slab`` is a torch tensor ``(own_z1-own_z0, Y, X)`` (CPU with gloo, CUDA with NCCL).  For the
    columns below and above the interior run, a small slab ``raw[z0:z1]`` is assembled from this rank's
    own slices (copied) and its peers' (received).  Every rank posts the sends its peers need and the
    receives it needs in one ``batch_isend_irecv`` -- a sparse neighbour exchange, not a collective;
    both sides enumerate (peer, low, high) in the same order so the pairs match.
    Returns ``((low_slab, low_z0), (high_slab, high_z0), requests)``.
    """
    import torch
    import torch.distributed as dist

    def wire(t):   # torch's NCCL backend has no 16-bit integer types: ship the same bytes as uint8
        return t.view(torch.uint8) if t.dtype in (torch.uint16, torch.int16) else t

    me = shards[rank]
    ops, slabs = [], []
    for z0, z1 in (me.low_need_z, me.high_need_z):
        slab = torch.empty((max(z1 - z0, 0),) + tuple(own_slab.shape[1:]), dtype=own_slab.dtype, device=own_slab.device)
        a, b = max(z0, me.own_z[0]), min(z1, me.own_z[1])
        if b > a:
            slab[a - z0:b - z0].copy_(own_slab[a - me.own_z[0]:b - me.own_z[0]])
        slabs.append((slab, z0))
    for other in shards:
        if other.rank == rank:
            continue
        for which, (z0, z1) in enumerate((me.low_need_z, me.high_need_z)):      # what I need from `other`
            a, b = max(z0, other.own_z[0]), min(z1, other.own_z[1])
            if b > a:
                ops.append(dist.P2POp(dist.irecv, wire(slabs[which][0][a - z0:b - z0]), other.rank, group=group))
        for z0, z1 in (other.low_need_z, other.high_need_z):                     # what `other` needs from me
            a, b = max(z0, me.own_z[0]), min(z1, me.own_z[1])
            if b > a:
                ops.append(dist.P2POp(dist.isend, wire(own_slab[a - me.own_z[0]:b - me.own_z[0]]), other.rank,
                                      group=group))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    return slabs[0], slabs[1], reqs


class PeerSlab:
    """This rank's raw slices in peer-mapped (symmetric) device memory: every rank of the box can read them with
    plain loads over NVLink / NVSwitch, so the halo of a scan-axis split is PULLED by a device copy -- no NCCL call,
    no staging, nothing for the owner to do.  One allocation of the same size on every rank (the largest slab),
    one rendezvous; ``tensor`` is the view holding this rank's ``own_z`` slices."""

    def __init__(self, shards: Sequence[ScanShard], rank: int, frame_shape: Tuple[int, int], dtype, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.shards, self.rank, self.frame_shape, self.dtype = list(shards), rank, tuple(frame_shape), dtype
        self.itemsize = torch.empty((), dtype=dtype).element_size()
        frame_bytes = frame_shape[0] * frame_shape[1] * self.itemsize
        nbytes = max(1, max(s.own_z[1] - s.own_z[0] for s in shards)) * frame_bytes
        self._buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self._hdl = symm_mem.rendezvous(self._buf, group if group is not None else dist.group.WORLD)
        self.tensor = self._view(self._buf, shards[rank])

    def _view(self, flat, shard: ScanShard):
        nz = shard.own_z[1] - shard.own_z[0]
        nbytes = nz * self.frame_shape[0] * self.frame_shape[1] * self.itemsize
        return flat[:nbytes].view(self.dtype).view((nz,) + self.frame_shape)

    def peer(self, other_rank: int):
        """Rank ``other_rank``'s slices as a tensor on THIS device (loads go over NVLink)."""
        import torch

        flat = self._hdl.get_buffer(other_rank, (self._buf.numel(),), torch.uint8)
        return self._view(flat, self.shards[other_rank])

    def barrier(self) -> None:
        """Device-side barrier over the ranks on the current stream (call after filling ``tensor``, before peers read)."""
        self._hdl.barrier()


def exchange_halos_peer(slab: "PeerSlab", shards: Sequence[ScanShard], rank: int, stream=None):
    """The peer-memory form of ``exchange_halos``: the slabs for the boundary column ranges are assembled by device
    copies that read the neighbours' slices directly from their HBM (on ``stream`` when given, so that the interior
    columns are computed meanwhile).  Returns ``((low_slab, low_z0), (high_slab, high_z0), event)``."""
    import torch

    me = shards[rank]
    own = slab.tensor
    ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream())
    if stream is not None:
        stream.wait_stream(torch.cuda.current_stream())
    out = []
    with ctx:
        for z0, z1 in (me.low_need_z, me.high_need_z):
            piece = torch.empty((max(z1 - z0, 0),) + tuple(own.shape[1:]), dtype=own.dtype, device=own.device)
            for other in shards:
                a, b = max(z0, other.own_z[0]), min(z1, other.own_z[1])
                if b > a:
                    src = own if other.rank == rank else slab.peer(other.rank)
                    piece[a - z0:b - z0].copy_(src[a - other.own_z[0]:b - other.own_z[0]], non_blocking=True)
            out.append((piece, z0))
        event = torch.cuda.Event()
        event.record()
    return out[0], out[1], event


def deskew_scan_split(own_slab, g: DeskewGeometry, shards: Sequence[ScanShard], rank: int, *, cval: float = 0.0,
                      group=None, window_fn: Optional[Callable] = None, peer_stream=None):
    """Deskew this rank's output columns of one volume that is split along the scan axis.

    Returns the compact tensor ``out[:, :, c0:c1]``.  ``window_fn(slab, g, p_begin, p_count, c_begin,
    c_count, y_origin, z_origin, cval)`` defaults to the CUDA window kernel (which writes straight into
    the strided output view); the CPU tests inject a stand-in.  The interior columns are computed from
    the rank's own slices while the halo for the boundary columns is in flight.
    """
    import torch

    in_place = window_fn is None
    if in_place:
        from .deskew import deskew_window

        def window_fn(slab, g, p_begin, p_count, c_begin, c_count, y_origin, z_origin, cval, out=None):
            return deskew_window(slab, g, p_begin=p_begin, p_count=p_count, c_begin=c_begin, c_count=c_count,
                                 y_origin=y_origin, z_origin=z_origin, cval=cval, out=out)

    me = shards[rank]
    Yn, X, _ = g.out_shape
    c0, c1 = me.cols
    device = own_slab.tensor.device if isinstance(own_slab, PeerSlab) else own_slab.device
    out = torch.empty((Yn, X, c1 - c0), dtype=torch.float32, device=device)

    def run(slab, z_origin, a, b):
        if b <= a:
            return
        view = out[:, :, a - c0:b - c0]
        if slab.shape[0] == 0:              # these columns read nothing from the volume
            view.fill_(cval)
        elif in_place:
            window_fn(slab, g, 0, Yn, a, b - a, 0, z_origin, cval, out=view)
        else:
            view.copy_(window_fn(slab, g, 0, Yn, a, b - a, 0, z_origin, cval))

    if isinstance(own_slab, PeerSlab):
        # halo pulled from the neighbours' HBM by device copies on a side stream (no NCCL on the data path)
        peer, own_slab = own_slab, own_slab.tensor
        (low, low_z0), (high, high_z0), ready = exchange_halos_peer(peer, shards, rank, stream=peer_stream)
        run(own_slab, me.own_z[0], *me.interior_cols)     # overlaps with the pull
        torch.cuda.current_stream().wait_event(ready)
        run(low, low_z0, *me.low_cols)
        run(high, high_z0, *me.high_cols)
        return out
    (low, low_z0), (high, high_z0), reqs = exchange_halos(own_slab, shards, rank, group=group)
    run(own_slab, me.own_z[0], *me.interior_cols)     # overlaps with the exchange
    for r in reqs:
        r.wait()
    run(low, low_z0, *me.low_cols)
    run(high, high_z0, *me.high_cols)
    return out
