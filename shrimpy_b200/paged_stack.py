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

Status: the layout arithmetic and the descriptor exchange are covered by CPU tests (``tests/test_paged_stack.py``:
every byte a rank's columns read is mapped, stitched windows equal the stack, 3-process fd hand-over over unix
sockets, the whole call sequence against an emulation of the driver's virtual-memory calls).  The driver calls were written after round 1's GPU minutes ran out and have NOT run on a GPU yet:
``tools/scan_split_bench.py --transport vmm`` is the first thing to run (2 GPUs) before anything relies on it.
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
    maps: Tuple[Tuple[int, int, int, int], ...]   # (owner rank, offset in owner's pages, offset in window, size)

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
    shards = []
    for s in base:
        lo, hi = owned[s.rank]
        if s.need_z[1] > s.need_z[0]:
            need_lo = s.need_z[0] * frame_bytes // G * G
            need_hi = min(-(-s.need_z[1] * frame_bytes // G) * G, total)
            wlo, whi = (min(lo, need_lo), max(hi, need_hi)) if hi > lo else (need_lo, need_hi)
        else:
            wlo, whi = lo, hi
        maps = []
        for r, (a, b) in enumerate(owned):
            x0, x1 = max(a, wlo), min(b, whi)
            if x1 > x0:
                maps.append((r, x0 - a, x0 - wlo, x1 - x0))
        shards.append(PagedShard(s.rank, s.cols, s.need_z, (lo, hi), (wlo, whi), Z * frame_bytes, tuple(maps)))
    return shards


def exchange_descriptors(my_fd: int, shards: Sequence[PagedShard], rank: int, group=None) -> Dict[int, int]:
    """Hand this rank's exported file descriptor to every rank that maps its pages, and collect the descriptors of
    the ranks whose pages this rank maps (``SCM_RIGHTS`` over unix sockets: a descriptor is only meaningful inside
    the process that received it this way).  ``my_fd < 0`` means this rank owns no pages.  Returns {owner: fd}."""
    import torch.distributed as dist

    token = [uuid.uuid4().hex if rank == 0 else None]
    dist.broadcast_object_list(token, src=0, group=group)
    path = lambda r: f"/tmp/shrimpy_b200_pages_{token[0]}_{r}.sock"     # noqa: E731
    wanted = [owner for owner, _, _, _ in shards[rank].maps if owner != rank]
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
                    socket.send_fds(conn, [b"p"], [my_fd])
        except BaseException as exc:   # surfaced on the caller's thread below
            failure.append(exc)

    worker = threading.Thread(target=serve, daemon=True)
    try:
        dist.barrier(group=group)              # every listener is bound before anyone connects
        worker.start()
        got: Dict[int, int] = {}
        for owner in wanted:
            with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
                c.connect(path(owner))
                _, fds, _, _ = socket.recv_fds(c, 16, 1)
                if len(fds) != 1:
                    raise RuntimeError(f"rank {owner} sent {len(fds)} descriptors instead of one")
                got[owner] = fds[0]
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
        self._handles: Dict[int, object] = {}
        self._release: List[int] = []          # the handles this object has to release (all but borrowed ones)
        self._local = _local_handles is not None
        self._fds: List[int] = []
        self._va, self._mapped = None, []
        _make_current(self.device_index)
        lo, hi = self.shard.own_bytes
        wlo, whi = self.shard.window_bytes
        if (hi - lo) % granularity or (whi - wlo) % granularity:
            raise ValueError("the plan was made for another granularity")
        kind = drv.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        my_fd = -1
        try:
            if self._local:            # every rank's pages were created in this process (``on_one_device``)
                self._handles = dict(_local_handles)
                self._release = [rank] if rank in self._handles else []
            else:
                if hi > lo:
                    self._handles[rank] = _ck(drv.cuMemCreate(hi - lo, self.allocation_prop(self.device_index), 0),
                                              "cuMemCreate")
                    self._release.append(rank)
                    my_fd = int(_ck(drv.cuMemExportToShareableHandle(self._handles[rank], kind, 0),
                                    "cuMemExportToShareableHandle"))
                    self._fds.append(my_fd)
                for owner, fd in exchange_descriptors(my_fd, self.shards, rank, group).items():
                    self._fds.append(fd)
                    self._handles[owner] = _ck(drv.cuMemImportFromShareableHandle(fd, kind),
                                               "cuMemImportFromShareableHandle")
                    self._release.append(owner)
            if whi > wlo:
                self._va = int(_ck(drv.cuMemAddressReserve(whi - wlo, granularity, 0, 0), "cuMemAddressReserve"))
                for owner, h_off, w_off, size in self.shard.maps:
                    _ck(drv.cuMemMap(self._va + w_off, size, h_off, self._handles[owner], 0), "cuMemMap")
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
            lo, hi = s.own_bytes
            if hi > lo:
                handles[s.rank] = _ck(drv.cuMemCreate(hi - lo, cls.allocation_prop(device_index), 0), "cuMemCreate")
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
        for owner in self._release:
            drv.cuMemRelease(self._handles[owner])
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

    Yn, X, _ = g.out_shape
    c0, c1 = shard.cols
    z0, z1 = shard.need_z
    if c1 <= c0 or z1 <= z0:      # no columns, or columns that read nothing from the volume
        return torch.full((Yn, X, max(c1 - c0, 0)), float(cval), dtype=torch.float32, device=stack.device)
    return window_fn(stack.slices(z0, z1), g, 0, Yn, c0, c1 - c0, 0, z0, cval)
