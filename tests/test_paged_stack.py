"""Byte-paged scan split (``shrimpy_b200/paged_stack.py``) on CPU: the layout arithmetic, a host stand-in for the
stitched window (same ``slices`` / ``fill_own`` code as the CUDA class) driven through the numpy window kernel, and the
file-descriptor hand-over between processes.  The CUDA driver calls themselves need GPUs (tools/scan_split_bench.py
--transport vmm)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import shrimpy_b200 as sb
from helpers import synthetic_stack
from oracle import deskew_oracle as o
from shrimpy_b200 import paged_stack as ps
from test_sharding import numpy_window


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("keep", [True, False])
@pytest.mark.parametrize("granularity", [2 << 20, 64 << 10])
def test_plan_partitions_the_stack_into_pages_and_maps_every_needed_byte(world, keep, granularity):
    g = sb.deskew_geometry((4000, 300, 2048), 30.0, 0.39, keep, 1)          # BASELINE configs[4] geometry
    F = 300 * 2048 * 2
    shards = ps.plan_paged_split(g, world, F, granularity)
    total = -(-4000 * F // granularity) * granularity
    assert shards[0].own_bytes[0] == 0 and shards[-1].own_bytes[1] == total
    assert shards[0].cols[0] == 0 and shards[-1].cols[1] == g.out_shape[2]
    for a, b in zip(shards[:-1], shards[1:]):
        assert a.own_bytes[1] == b.own_bytes[0] and a.cols[1] == b.cols[0]
    for s in shards:
        assert all(v % granularity == 0 for v in s.own_bytes + s.window_bytes) and s.stack_bytes == 4000 * F
        wlo, whi = s.window_bytes
        if s.need_z[1] > s.need_z[0]:
            assert wlo <= s.need_z[0] * F and s.need_z[1] * F <= whi          # every byte the columns read is visible
        assert wlo <= s.own_bytes[0] and s.own_bytes[1] <= whi                # and everything the rank has to load
        # the maps tile the window in order, without gaps, from inside each owner's pages
        at = 0
        for owner, h_off, w_off, size in s.maps:
            assert w_off == at and size > 0 and size % granularity == 0 and h_off % granularity == 0
            lo, hi = shards[owner].own_bytes
            assert lo + h_off == wlo + w_off and lo + h_off + size <= hi
            at += size
        assert at == whi - wlo
        # a handle is only ever mapped whole: every map is exactly one segment of its owner
        for owner, h_off, _, size in s.maps:
            lo = shards[owner].own_bytes[0]
            assert (lo + h_off, lo + h_off + size) in shards[owner].segments
        assert sum(b - a for a, b in s.segments) == s.own_bytes[1] - s.own_bytes[0]
        if world > 1:
            # the halo: about r*cos(theta)*(Y-1)+1 slices below, two above, rounded out to pages
            assert s.remote_bytes <= (int(0.39 * np.cos(np.pi / 6) * 299) + 5) * F + 3 * granularity


class HostPagedStack(ps.PagedWindow):
    """Stand-in for ``PagedStack``: every rank's pages are numpy bytes, the window is stitched by copying them in the
    order of ``shard.maps`` (what ``cuMemMap`` does with address translation)."""

    def __init__(self, shards, rank, frame_shape, dtype, pages_of_rank):
        self.shards, self.rank, self.shard = shards, rank, shards[rank]
        self.frame_shape, self.dtype = tuple(frame_shape), dtype
        self.frame_bytes = frame_shape[0] * frame_shape[1] * torch.empty((), dtype=dtype).element_size()
        wlo, whi = self.shard.window_bytes
        window = np.zeros(whi - wlo, dtype=np.uint8)
        for owner, h_off, w_off, size in self.shard.maps:
            window[w_off:w_off + size] = pages_of_rank[owner][h_off:h_off + size]
        self._window = torch.from_numpy(window)
        lo, hi = self.shard.own_bytes
        self.own = self._window[lo - wlo:hi - wlo]


@pytest.mark.parametrize("params,world,granularity", [
    (((120, 13, 8), 30.0, 0.39, True, 1), 3, 64),        # pages far smaller than a slice (13*8*2 = 208 B)
    (((150, 10, 8), 30.0, 0.39, False, 3), 2, 4096),     # pages of ~25 slices
    (((60, 9, 16), 30.0, 0.39, True, 2), 4, 1 << 20),    # one page holds the whole stack: rank 0 owns it all
], ids=["small-pages", "medium-pages", "one-page"])
def test_stitched_windows_deskew_like_the_whole_stack(params, world, granularity):
    shape, ang, r, keep, n = params
    raw = synthetic_stack(shape, seed=21)
    g = sb.deskew_geometry(shape, ang, r, keep, n)
    F = shape[1] * shape[2] * 2
    shards = ps.plan_paged_split(g, world, F, granularity, align=8)
    flat = np.zeros(shards[-1].own_bytes[1], dtype=np.uint8)
    flat[:raw.nbytes] = raw.reshape(-1).view(np.uint8)
    # every rank loads its own pages through fill_own (slices in, bytes out) ...
    pages = []
    for s in shards:
        loader = HostPagedStack(shards, s.rank, shape[1:], torch.uint16,
                                [np.zeros(t.own_bytes[1] - t.own_bytes[0], np.uint8) for t in shards])
        loader.fill_own(lambda z0, z1: torch.from_numpy(raw[z0:z1].copy()))
        mine = loader.own.numpy().copy()
        valid = min(s.own_bytes[1], raw.nbytes) - s.own_bytes[0]
        assert np.array_equal(mine[:max(valid, 0)], flat[s.own_bytes[0]:s.own_bytes[0] + max(valid, 0)])
        pages.append(mine)
    # ... and deskews its columns in one window call over the stitched range
    pieces = []
    for s in shards:
        stack = HostPagedStack(shards, s.rank, shape[1:], torch.uint16, pages)
        if s.need_z[1] > s.need_z[0]:
            assert np.array_equal(stack.slices(*s.need_z).numpy(), raw[s.need_z[0]:s.need_z[1]])
        with pytest.raises(ValueError):
            stack.slices(0, shape[0] + 10**6)
        pieces.append(ps.deskew_paged_split(stack, g, s, cval=-3.0, window_fn=numpy_window).numpy())
    got = np.concatenate(pieces, axis=2)
    whole = o.deskew_data(raw, ang, r, keep, n, cval=-3.0)
    assert got.shape == whole.shape
    assert np.array_equal(got == -3.0, whole == -3.0)
    assert np.max(np.abs(got - whole)) <= 2e-6 * float(whole.max() - whole.min())


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fd_worker(rank, world, port, outdir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = sb.deskew_geometry((4000, 30, 64), 30.0, 0.39, True, 1)
        shards = ps.plan_paged_split(g, world, 30 * 64 * 2, 64 << 10)
        lo = shards[rank].own_bytes[0]
        mine = {}
        for a, b in shards[rank].segments:              # memfds stand for the exported CUDA allocation handles
            mine[a - lo] = os.memfd_create(f"pages{rank}_{a - lo}")
            os.write(mine[a - lo], f"segment {a - lo} of rank {rank}".encode())
        got = ps.exchange_descriptors(mine, shards, rank)
        assert sorted(got) == sorted((owner, off) for owner, off, _, _ in shards[rank].maps if owner != rank)
        for (owner, off), peer_fd in got.items():
            assert os.pread(peer_fd, 64, 0) == f"segment {off} of rank {owner}".encode()
            os.close(peer_fd)
        for fd in mine.values():
            os.close(fd)
        with open(os.path.join(outdir, f"ok{rank}"), "w") as f:
            f.write(str(len({owner for owner, _ in got})))
    finally:
        dist.destroy_process_group()


def test_descriptor_hand_over_world3(tmp_path):
    world = 3
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_fd_worker, args=(r, world, port, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    counts = [int((tmp_path / f"ok{r}").read_text()) for r in range(world)]
    assert counts == [1, 2, 1]          # the middle rank maps both neighbours, the ends one each


def test_device_slab_describes_a_contiguous_stack_like_a_tensor():
    slab = ps.DeviceSlab(0x7f0000000000, (5, 13, 8), torch.uint16, torch.device("cpu"))
    like = torch.empty((5, 13, 8), dtype=torch.uint16)
    assert slab.shape == tuple(like.shape) and slab.dtype == like.dtype
    assert [slab.stride(i) for i in range(3)] == list(like.stride()) and slab.data_ptr() == 0x7f0000000000


class EmulatedDriver:
    """The virtual-memory calls ``PagedStack`` makes, emulated on host bytes: physical handles are numpy arrays, a
    reserved range is a number, ``cuMemMap`` records which bytes of which handle answer at which address.  Enums and
    structs are cuda-python's own (they import without a driver), so a misspelt field fails here as it would there."""

    def __init__(self):
        from cuda.bindings import driver

        self._real = driver
        self.storage, self.mappings, self.reserved, self.access, self.released = {}, [], {}, [], []
        self._arena, self._arena_used, self._arena_at = np.zeros(64 << 20, dtype=np.uint8), 0, {}
        self._next_handle, self._next_va = 1, 0x7F0000000000

    def __getattr__(self, name):
        return getattr(self._real, name)

    def cuMemGetAllocationGranularity(self, prop, option):
        assert prop.requestedHandleTypes == self._real.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        return (self._real.CUresult.CUDA_SUCCESS, 4096)

    def cuMemCreate(self, size, prop, flags):
        assert size > 0 and size % 4096 == 0 and flags == 0
        assert prop.location.type == self._real.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        handle, self._next_handle = self._next_handle, self._next_handle + 1
        # handles created one after the other are neighbours in one arena, so that a run of a rank's segments
        # mapped side by side can be handed out as ONE writable host view (what `own` is on the device)
        self.storage[handle] = self._arena[self._arena_used:self._arena_used + size]
        self._arena_at[handle] = self._arena_used
        self._arena_used += size
        assert self._arena_used <= self._arena.size
        return (self._real.CUresult.CUDA_SUCCESS, handle)

    def cuMemAddressReserve(self, size, alignment, addr, flags):
        va, self._next_va = self._next_va, self._next_va + size + (1 << 30)
        self.reserved[va] = size
        return (self._real.CUresult.CUDA_SUCCESS, self._real.CUdeviceptr(va))

    def cuMemMap(self, ptr, size, offset, handle, flags):
        base = max(v for v in self.reserved if v <= ptr)
        assert ptr + size <= base + self.reserved[base], "mapping outside the reserved range"
        assert offset == 0 and size == self.storage[handle].size, "the driver maps whole handles only (offset 0)"
        assert all(ptr + size <= p or p + n <= ptr for p, n, _, _ in self.mappings), "address mapped twice"
        self.mappings.append((ptr, size, handle, offset))
        return (self._real.CUresult.CUDA_SUCCESS,)

    def cuMemSetAccess(self, ptr, size, desc, count):
        assert count == 1 and desc[0].flags == self._real.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
        self.access.append((ptr, size, desc[0].location.id))
        return (self._real.CUresult.CUDA_SUCCESS,)

    def cuMemUnmap(self, ptr, size):
        self.mappings = [m for m in self.mappings if not (m[0] == ptr and m[1] == size)]
        return (self._real.CUresult.CUDA_SUCCESS,)

    def cuMemRelease(self, handle):
        self.released.append(handle)
        return (self._real.CUresult.CUDA_SUCCESS,)

    def cuMemAddressFree(self, ptr, size):
        assert self.reserved.pop(ptr) == size
        return (self._real.CUresult.CUDA_SUCCESS,)

    # what the hardware does with the mappings
    def view(self, ptr, nbytes):
        at, start = ptr, None
        while at < ptr + nbytes:
            p, n, handle, offset = next(m for m in self.mappings if m[0] <= at < m[0] + m[1])
            here = self._arena_at[handle] + offset + at - p
            if start is None:
                start = here
            assert here == start + (at - ptr), "the mapped handles are not neighbours in the arena"
            at = p + n
        return self._arena[start:start + nbytes]

    def read(self, ptr, nbytes):
        out = np.empty(nbytes, dtype=np.uint8)
        done = 0
        while done < nbytes:
            p, n, handle, offset = next(m for m in self.mappings if m[0] <= ptr + done < m[0] + m[1])
            take = min(nbytes - done, p + n - (ptr + done))
            out[done:done + take] = self.storage[handle][offset + ptr + done - p:offset + ptr + done - p + take]
            done += take
        return out


@pytest.mark.parametrize("world,keep,n", [(1, True, 1), (3, True, 1), (4, False, 3)])
def test_paged_stack_call_sequence_on_an_emulated_driver(monkeypatch, world, keep, n):
    """``PagedStack.on_one_device`` end to end with the driver emulated: the addresses ``slices`` hands to the kernel
    resolve, through the recorded mappings, to exactly the raw slices; everything is unmapped and released on close."""
    drv = EmulatedDriver()
    monkeypatch.setattr(ps, "_driver", lambda: drv)
    monkeypatch.setattr(ps, "_make_current", lambda device_index: None)
    monkeypatch.setattr(ps, "_alias_bytes", lambda ptr, nbytes, device_index: torch.from_numpy(drv.view(ptr, nbytes)))
    shape = (260, 12, 64)                                   # 1536-byte slices, 4096-byte pages
    raw = synthetic_stack(shape, seed=33)
    g = sb.deskew_geometry(shape, 30.0, 0.39, keep, n)
    page = ps.PagedStack.granularity(0)
    shards = ps.plan_paged_split(g, world, shape[1] * shape[2] * 2, page, align=8)
    stacks = ps.PagedStack.on_one_device(shards, shape[1:], torch.uint16, 0, page)
    for st in stacks:
        st.fill_own(lambda z0, z1: torch.from_numpy(raw[z0:z1].copy()))
    assert sorted(a[0] for a in drv.access) == sorted(st._va for st in stacks)

    def window_through_the_mappings(slab, g, p_begin, p_count, c_begin, c_count, y_origin, z_origin, cval):
        assert isinstance(slab, ps.DeviceSlab)
        data = drv.read(slab.data_ptr(), int(np.prod(slab.shape)) * 2).view(np.uint16).reshape(slab.shape)
        assert np.array_equal(data, raw[z_origin:z_origin + slab.shape[0]])
        return numpy_window(torch.from_numpy(data), g, p_begin, p_count, c_begin, c_count, y_origin, z_origin, cval)

    pieces = [ps.deskew_paged_split(st, g, st.shard, cval=-3.0, window_fn=window_through_the_mappings).numpy()
              for st in stacks if st.shard.need_z[1] > st.shard.need_z[0]]
    whole = o.deskew_data(raw, 30.0, 0.39, keep, n, cval=-3.0)
    got = np.concatenate(pieces, axis=2)
    assert got.shape == whole.shape and np.array_equal(got == -3.0, whole == -3.0)
    assert np.max(np.abs(got - whole)) <= 2e-6 * float(whole.max() - whole.min())
    for st in stacks:
        st.close()
    assert not drv.mappings and not drv.reserved and sorted(drv.released) == sorted(drv.storage)
