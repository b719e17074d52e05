"""One oversized raw stack spread over the GPUs of a box as ONE virtual array (BASELINE configs[4]).

``sharding.deskew_scan_split`` exchanges the halo of a scan-axis split in a step of its own (device copies out of the
neighbour's peer-mapped memory, or NCCL send/recv) and then launches the window kernel three times (interior, low and
high columns).  Here there is no exchange step and no second launch: the stack is cut into **granularity-aligned byte
pages** (CUDA virtual memory management, ``cuMemCreate``), every rank owns a contiguous run of pages as physical memory
on its GPU, and maps the pages of its neighbours that its output columns read *next to its own in one contiguous
virtual range* (``cuMemAddressReserve`` + ``cuMemMap`` of the handles imported from the neighbours).  The raw slices a
rank needs are then one ordinary ``(nz, Y, X)`` array at one base address, so the unchanged TMA deskew kernel computes
all of the rank's columns in ONE launch and its tile loads of halo rows are NVLink reads of the neighbour's HBM, issued
tile by tile while the other tiles compute -- the transfer is fused into the kernel by the address map, not by a
second code path.

Cutting by bytes rather than by slices is what makes the pieces adjacent: a physical handle can only be mapped at
multiples of the allocation granularity (2 MiB), and a mantis slice (300 x 2048 uint16 = 1 228 800 B) is not one; the
slice that straddles a page boundary simply has its head on one GPU and its tail on the next.

A physical handle can only be mapped WHOLE (``cuMemMap`` wants ``offset == 0``: the first B200 run of this module
answered ``CUDA_ERROR_NOT_SUPPORTED`` to a partial mapping), so a rank's bytes are not one allocation but a few
**segments**, cut wherever some rank's window begins or ends: a reader's window is then a union of whole segments, and
only the segments a neighbour reads are exported.

Covered on CPU (``tests/test_paged_stack.py``): the layout arithmetic (every byte a rank's columns read is mapped,
windows are tiled by whole segments), stitched windows equal the stack, the 3-process fd hand-over over unix sockets,
the whole call sequence against an emulation of the driver's virtual-memory calls.  On GPUs:
``tests/test_paged_stack_gpu.py`` (one device) and ``tools/scan_split_bench.py --transport vmm`` / ``bench.py`` (N > 1).
"""

from __future__ import annotations

import os
import socket
import threading
import uuid
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

from .deskew import DeskewGeometry
from .sharding import plan_scan_split

__all__ = ["PagedShard", "plan_paged_split", "exchange_descriptors", "PagedWindow", "PagedStack", "DeviceSlab",
           "deskew_paged_split"]


@dataclass(frozen=True)
class PagedShard:
    """What one rank owns, maps and computes when the stack is cut into byte pages."""

    rank: int
    cols: Tuple[int, int]            # output columns (o2) this rank computes                         [c0, c1)
    need_z: Tuple[int, int]          # raw scan slices those columns read                             [z0, z1)
    own_bytes: Tuple[int, int]       # bytes of the flattened stack held in this rank's HBM           [lo, hi), page aligned
    window_bytes: Tuple[int, int]    # bytes of the flattened stack visible in this rank's window     [lo, hi), page aligned
    stack_bytes: int                 # Z * frame_bytes: the last page may reach beyond it
    maps: Tuple[Tuple[int, int, int, int], ...]   # (owner rank, offset in owner's pages, offset in window, size):
    #                                               each entry is one WHOLE segment of its owner
    segments: Tuple[Tuple[int, int], ...] = ()    # this rank's bytes as separate physical allocations [lo, hi)

    @property
    def remote_bytes(self) -> int:
        """Bytes of the window that live in another GPU's HBM."""
        return sum(size for owner, _, _, size in self.maps if owner != self.rank)


def plan_paged_split(g: DeskewGeometry, world_size: int, frame_bytes: int, granularity: int,
                     align: int = 32) -> List[PagedShard]:
    """Columns and needed slices as in ``plan_scan_split``; ownership boundaries rounded down to whole pages of the
    flattened stack.  Every rank's window is the page-aligned hull of what it owns and what its columns read."""
    if frame_bytes <= 0 or granularity <= 0:
        raise ValueError("frame_bytes and granularity must be positive")
    Z = g.raw_shape[0]
    base = plan_scan_split(g, world_size, align)
    G = int(granularity)
    total = -(-Z * frame_bytes // G) * G
    edges = [0]
    for s in base[1:]:
        edges.append(min(max(s.own_z[0] * frame_bytes // G * G, edges[-1]), total))
    edges.append(total)
    owned = [(edges[i], edges[i + 1]) for i in range(world_size)]
    windows = []
    for s in base:
        lo, hi = owned[s.rank]
        if s.need_z[1] > s.need_z[0]:
            need_lo = s.need_z[0] * frame_bytes // G * G
            need_hi = min(-(-s.need_z[1] * frame_bytes // G) * G, total)
            windows.append((min(lo, need_lo), max(hi, need_hi)) if hi > lo else (need_lo, need_hi))
        else:
            windows.append((lo, hi))
    # a handle is mapped whole: cut every rank's bytes wherever any window begins or ends
    cuts = sorted({v for w in windows for v in w} | set(edges))
    segments = [tuple((a, b) for a, b in zip(cuts[:-1], cuts[1:]) if lo <= a and b <= hi) for lo, hi in owned]
    shards = []
    for s in base:
        wlo, whi = windows[s.rank]
        maps = tuple((r, a - owned[r][0], a - wlo, b - a)
                     for r in range(world_size) for a, b in segments[r] if wlo <= a and b <= whi)
        shards.append(PagedShard(s.rank, s.cols, s.need_z, owned[s.rank], (wlo, whi), Z * frame_bytes, maps,
                                 segments[s.rank]))
    return shards


def exchange_descriptors(my_fds: Dict[int, int], shards: Sequence[PagedShard], rank: int,
                         group=None) -> Dict[Tuple[int, int], int]:
    """Hand this rank's exported file descriptors (``{offset of the segment in this rank's bytes: fd}``) to every
    rank that maps one of its segments, and collect the descriptors of the segments this rank maps (``SCM_RIGHTS``
    over unix sockets: a descriptor is only meaningful inside the process that received it this way).  A client
    names the segments it wants; the owner answers with exactly those.  Returns ``{(owner, offset): fd}``."""
    import json

    import torch.distributed as dist

    token = [uuid.uuid4().hex if rank == 0 else None]
    dist.broadcast_object_list(token, src=0, group=group)
    path = lambda r: f"/tmp/shrimpy_b200_pages_{token[0]}_{r}.sock"     # noqa: E731
    wanted: Dict[int, List[int]] = {}
    for owner, h_off, _, _ in shards[rank].maps:
        if owner != rank:
            wanted.setdefault(owner, []).append(h_off)
    clients = sum(1 for s in shards if s.rank != rank and any(owner == rank for owner, _, _, _ in s.maps))
    server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    server.bind(path(rank))
    server.listen(max(1, len(shards)))
    failure: List[BaseException] = []

    def serve():
        try:
            for _ in range(clients):
                conn, _ = server.accept()
                with conn:
                    offsets = json.loads(conn.recv(65536).decode())
                    socket.send_fds(conn, [b"p"], [my_fds[o] for o in offsets])
        except BaseException as exc:   # surfaced on the caller's thread below
            failure.append(exc)

    worker = threading.Thread(target=serve, daemon=True)
    try:
        dist.barrier(group=group)              # every listener is bound before anyone connects
        worker.start()
        got: Dict[Tuple[int, int], int] = {}
        for owner, offsets in wanted.items():
            with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
                c.connect(path(owner))
                c.sendall(json.dumps(offsets).encode())
                _, fds, _, _ = socket.recv_fds(c, 16, len(offsets))
                if len(fds) != len(offsets):
                    raise RuntimeError(f"rank {owner} sent {len(fds)} descriptors instead of {len(offsets)}")
                for o, fd in zip(offsets, fds):
                    got[(owner, o)] = fd
        worker.join(timeout=120)
        if worker.is_alive() or failure:
            raise RuntimeError(f"descriptor hand-over failed on rank {rank}: {failure or 'peer never connected'}")
        dist.barrier(group=group)
        return got
    finally:
        server.close()
        try:
            os.unlink(path(rank))
        except OSError:
            pass


class _DeviceBytes:
    """``__cuda_array_interface__`` carrier: lets torch alias a mapped virtual range without owning it."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class DeviceSlab:
    """A contiguous ``(nz, Y, X)`` array in device address space, described the way ``deskew_window`` reads a tensor
    (``data_ptr``, ``shape``, ``stride``, ``dtype``, ``device``) without being one."""

    def __init__(self, ptr: int, shape: Tuple[int, int, int], dtype, device):
        self._ptr, self.shape, self.dtype, self.device = int(ptr), tuple(shape), dtype, device

    def data_ptr(self) -> int:
        return self._ptr

    def stride(self, axis: int) -> int:
        return (self.shape[1] * self.shape[2], self.shape[2], 1)[axis]


def _driver():
    """The CUDA driver API (cuda-python).  A seam: the CPU tests substitute an emulation of the virtual-memory calls."""
    from cuda.bindings import driver

    return driver


def _make_current(device_index: int) -> None:
    """Make the device's primary context (torch's) current on this thread."""
    import torch

    torch.cuda.set_device(device_index)
    torch.empty(1, device="cuda")


def _alias_bytes(ptr: int, nbytes: int, device_index: int):
    """uint8 tensor over ``nbytes`` of LOCAL device memory at ``ptr`` (no copy, torch does not own it)."""
    import torch

    t = torch.as_tensor(_DeviceBytes(ptr, nbytes), device=f"cuda:{device_index}")
    if t.data_ptr() != ptr:
        raise RuntimeError("torch copied the mapped pages instead of aliasing them")
    return t


def _ck(result, what: str):
    """cuda-python returns ``(CUresult, values...)``; raise on anything but success, return the values."""
    err, rest = result[0], result[1:]
    if int(err) != 0:
        raise RuntimeError(f"{what} failed with CUresult {int(err)} ({err!s})")
    return rest[0] if len(rest) == 1 else rest


class PagedWindow:
    """Window arithmetic shared by ``PagedStack`` and the host stand-in of the CPU tests: ``_window`` is a uint8 tensor
    over ``shard.window_bytes`` of the flattened stack."""

    shard: PagedShard
    frame_shape: Tuple[int, int]
    frame_bytes: int

    @property
    def device(self):
        return self._window.device

    def slices(self, z0: int, z1: int):
        """Raw slices ``[z0, z1)`` as one ``(z1 - z0, Y, X)`` tensor on this device (remote pages are read over NVLink)."""
        wlo, whi = self.shard.window_bytes
        a, b = z0 * self.frame_bytes - wlo, z1 * self.frame_bytes - wlo
        if z1 < z0 or a < 0 or b > whi - wlo:
            raise ValueError(f"slices [{z0},{z1}) lie outside this rank's window")
        return self._window[a:b].view(self.dtype).view((z1 - z0,) + self.frame_shape)

    def fill_own(self, slices_of) -> None:
        """Load this rank's bytes: ``slices_of(z0, z1)`` returns raw slices ``[z0, z1)`` as a tensor on this device."""
        import torch

        lo, hi = self.shard.own_bytes
        hi = min(hi, self.shard.stack_bytes)           # the last page may be padding beyond the stack
        if hi <= lo:
            return
        z0, z1 = lo // self.frame_bytes, -(-hi // self.frame_bytes)
        src = slices_of(z0, z1).contiguous().view(torch.uint8).reshape(-1)
        start = lo - z0 * self.frame_bytes
        self.own[:hi - lo].copy_(src[start:start + hi - lo])


class PagedStack(PagedWindow):
    """This rank's pages of the stack (physical memory on its GPU) and its window onto the neighbours' pages.

    ``own``     uint8 tensor over the bytes this rank holds (``shard.own_bytes`` of the flattened stack): the loader's
                destination.  Call ``barrier()`` after filling it and before any rank computes.
    ``slices``  raw slices inside the window as one ``(z1 - z0, Y, X)`` array at one base address, whoever holds them
                (a ``DeviceSlab``: an address and a shape, which is all the window kernel takes from a tensor).
    """

    def __init__(self, shards: Sequence[PagedShard], rank: int, frame_shape: Tuple[int, int], dtype, device_index: int,
                 granularity: int, group=None, _local_handles=None):
        import torch

        drv = _driver()
        self._drv, self._torch, self._group = drv, torch, group
        self.shards, self.rank, self.shard = list(shards), rank, shards[rank]
        self.frame_shape, self.dtype, self.device_index = tuple(frame_shape), dtype, int(device_index)
        self.itemsize = torch.empty((), dtype=dtype).element_size()
        self.frame_bytes = self.frame_shape[0] * self.frame_shape[1] * self.itemsize
        self._handles: Dict[Tuple[int, int], object] = {}    # (owner, offset of the segment in the owner's bytes)
        self._release: List[Tuple[int, int]] = []            # the handles this object has to release (not borrowed ones)
        self._local = _local_handles is not None
        self._fds: List[int] = []
        self._va, self._mapped = None, []
        _make_current(self.device_index)
        lo, hi = self.shard.own_bytes
        wlo, whi = self.shard.window_bytes
        if (hi - lo) % granularity or (whi - wlo) % granularity:
            raise ValueError("the plan was made for another granularity")
        kind = drv.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        try:
            if self._local:            # every rank's segments were created in this process (``on_one_device``)
                self._handles = dict(_local_handles)
                self._release = [k for k in self._handles if k[0] == rank]
            else:
                read_by_others = {h_off for s in self.shards if s.rank != rank
                                  for owner, h_off, _, _ in s.maps if owner == rank}
                my_fds: Dict[int, int] = {}
                for a, b in self.shard.segments:
                    key = (rank, a - lo)
                    self._handles[key] = _ck(drv.cuMemCreate(b - a, self.allocation_prop(self.device_index), 0),
                                             "cuMemCreate")
                    self._release.append(key)
                    if key[1] in read_by_others:
                        fd = int(_ck(drv.cuMemExportToShareableHandle(self._handles[key], kind, 0),
                                     "cuMemExportToShareableHandle"))
                        my_fds[key[1]] = fd
                        self._fds.append(fd)
                for key, fd in exchange_descriptors(my_fds, self.shards, rank, group).items():
                    self._fds.append(fd)
                    self._handles[key] = _ck(drv.cuMemImportFromShareableHandle(fd, kind),
                                             "cuMemImportFromShareableHandle")
                    self._release.append(key)
            if whi > wlo:
                self._va = int(_ck(drv.cuMemAddressReserve(whi - wlo, granularity, 0, 0), "cuMemAddressReserve"))
                for owner, h_off, w_off, size in self.shard.maps:      # every entry is one whole handle: offset 0
                    _ck(drv.cuMemMap(self._va + w_off, size, 0, self._handles[(owner, h_off)], 0), "cuMemMap")
                    self._mapped.append((self._va + w_off, size))
                access = drv.CUmemAccessDesc()
                access.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
                access.location.id = self.device_index
                access.flags = drv.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
                _ck(drv.cuMemSetAccess(self._va, whi - wlo, [access], 1), "cuMemSetAccess")
        except BaseException:
            self.close()
            raise
        # torch only ever aliases the rank's OWN pages (local physical memory: the pointer's device is this one).  A
        # tensor over the whole window would start in a neighbour's pages, torch would attribute it to that GPU and
        # ``device=`` would then mean a copy; the window is handed to the kernels as a raw address (``DeviceSlab``).
        try:
            self.own = (_alias_bytes(self._va + (lo - wlo), hi - lo, self.device_index) if hi > lo
                        else torch.empty(0, dtype=torch.uint8))
        except BaseException:
            self.close()
            raise

    @property
    def device(self):
        return self._torch.device("cuda", self.device_index)

    def slices(self, z0: int, z1: int) -> "DeviceSlab":
        """Raw slices ``[z0, z1)`` of the window as one ``(z1 - z0, Y, X)`` array at one address (remote pages are read
        over NVLink); accepted by ``deskew_window`` in place of a tensor."""
        wlo, whi = self.shard.window_bytes
        a, b = z0 * self.frame_bytes - wlo, z1 * self.frame_bytes - wlo
        if z1 < z0 or a < 0 or b > whi - wlo or self._va is None:
            raise ValueError(f"slices [{z0},{z1}) lie outside this rank's window")
        return DeviceSlab(self._va + a, (z1 - z0,) + self.frame_shape, self.dtype, self.device)

    @classmethod
    def on_one_device(cls, shards: Sequence[PagedShard], frame_shape: Tuple[int, int], dtype, device_index: int,
                      granularity: int) -> List["PagedStack"]:
        """Every rank's pages on ONE GPU of one process, each rank's window stitched from them: the layout, the
        mapping calls and the kernel's reads through a stitched range, without the descriptor hand-over and without
        NVLink (a one-GPU check of everything else)."""
        drv = _driver()
        _make_current(device_index)
        handles = {}
        for s in shards:
            for a, b in s.segments:
                handles[(s.rank, a - s.own_bytes[0])] = _ck(
                    drv.cuMemCreate(b - a, cls.allocation_prop(device_index), 0), "cuMemCreate")
        return [cls(shards, s.rank, frame_shape, dtype, device_index, granularity, _local_handles=handles)
                for s in shards]

    @staticmethod
    def allocation_prop(device_index: int):
        drv = _driver()

        prop = drv.CUmemAllocationProp()
        prop.type = drv.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
        prop.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        prop.location.id = int(device_index)
        prop.requestedHandleTypes = drv.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        return prop

    @staticmethod
    def granularity(device_index: int) -> int:
        """Allocation granularity of exportable device memory on this GPU (the page size of the plan)."""
        drv = _driver()
        _make_current(device_index)
        flag = drv.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_RECOMMENDED
        return int(_ck(drv.cuMemGetAllocationGranularity(PagedStack.allocation_prop(device_index), flag),
                       "cuMemGetAllocationGranularity"))

    def barrier(self) -> None:
        """All ranks have filled their pages and the copies have landed (call before the first deskew)."""
        import torch.distributed as dist

        self._torch.cuda.synchronize()
        if not self._local:
            dist.barrier(group=self._group)

    def close(self) -> None:
        drv = self._drv
        self.own = None
        for ptr, size in self._mapped:
            drv.cuMemUnmap(ptr, size)
        self._mapped = []
        for key in self._release:
            drv.cuMemRelease(self._handles[key])
        self._handles, self._release = {}, []
        if self._va is not None:
            wlo, whi = self.shard.window_bytes
            drv.cuMemAddressFree(self._va, whi - wlo)
            self._va = None
        for fd in self._fds:
            try:
                os.close(fd)
            except OSError:
                pass
        self._fds = []


def deskew_paged_split(stack, g: DeskewGeometry, shard: PagedShard, *, cval: float = 0.0, window_fn=None):
    """This rank's output columns ``out[:, :, c0:c1]`` in ONE launch over its window of the paged stack.

    ``stack`` needs ``slices(z0, z1)`` (``PagedStack``, or a host stand-in in the CPU tests); ``window_fn`` defaults
    to the CUDA window kernel."""
    import torch

    if window_fn is None:
        from .deskew import deskew_window

        def window_fn(slab, g, p_begin, p_count, c_begin, c_count, y_origin, z_origin, cval):
            return deskew_window(slab, g, p_begin=p_begin, p_count=p_count, c_begin=c_begin, c_count=c_count,
                                 y_origin=y_origin, z_origin=z_origin, cval=cval)

    # (Dispatching the column tiles outermost, so that every tile that reads remote pages is in flight at the start of
    # the launch, was tried on 2 GPUs: 2.56 -> 4.54 ms -- the output writes lose their locality -- and removed;
    # profiles/r02_scan_split_n2_with_block_order_experiment.json.)
    Yn, X, _ = g.out_shape
    c0, c1 = shard.cols
    z0, z1 = shard.need_z
    if c1 <= c0 or z1 <= z0:      # no columns, or columns that read nothing from the volume
        return torch.full((Yn, X, max(c1 - c0, 0)), float(cval), dtype=torch.float32, device=stack.device)
    return window_fn(stack.slices(z0, z1), g, 0, Yn, c0, c1 - c0, 0, z0, cval)
