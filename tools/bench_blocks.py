"""Measurement blocks of ``bench.py`` beyond the headline number (also used by the stand-alone tools in this
directory).  Every function is collective over the ranks of the run (``dist`` is ``torch.distributed`` or ``None`` at
one GPU), times on the device or from a barrier to a barrier, and reduces with MAX over ranks.

* ``scan_split_block``   BASELINE.json configs[4]: ONE oversized volume (4000, 300, 2048) keep_overhang=True, n=1, cut
                         along the scan axis over the ranks, halo over NVLink (``sharding.py`` / ``paged_stack.py``)
* ``plate_block``        BASELINE.json configs[3]: a 96-position x 10-timepoint plate streamed from an OME-Zarr store
                         in RAM through ``plate.deskew_plate(rank, world_size)``
* ``e2e_block``          the headline workload through the public numpy API: pinned (the contract ``e2e``), pageable
                         (what ``scripts/measure_psf.py:239-246`` passes), online (``preprocessing.py:316`` ->
                         ``:408-413``: uint16 up, result stays on the device) and the copy-only floor of the box at N ranks
* ``affine_block``       BASELINE.json configs[2]
"""

from __future__ import annotations

import os
import shutil
import time
from pathlib import Path

import numpy as np


def _max_over_ranks(dist, value: float) -> float:
    import torch

    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(dist, value: float) -> float:
    import torch

    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def _all_true(dist, flag: bool) -> bool:
    import torch

    t = torch.tensor([int(bool(flag))], device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def _barrier(dist):
    import torch

    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def _timed(dist, fn, reps: int):
    """Median over ``reps`` of the max-over-ranks device time of ``fn()`` (barrier + synchronize on both sides)."""
    import torch

    times, last = [], None
    for _ in range(reps):
        _barrier(dist)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        last = fn()
        b.record()
        torch.cuda.synchronize()
        times.append(_max_over_ranks(dist, a.elapsed_time(b)))
    return sorted(times)[len(times) // 2], last


def _nvlink_rx_bytes(local: int):
    """Cumulative NVLink payload bytes RECEIVED by this GPU over all its links (NVML field
    NVLINK_THROUGHPUT_DATA_RX, KiB), or None when the driver does not expose it.  A hardware counter outside our code:
    the evidence that a kernel's loads of mapped neighbour pages really travel over NVLink."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip().isdigit()]
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(vis[local]) if local < len(vis) else local)
        vals = pynvml.nvmlDeviceGetFieldValues(handle, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF)])
        if vals[0].nvmlReturn != 0:
            return None
        return int(vals[0].value.ullVal) * 1024
    except Exception:
        return None


# --------------------------------------------------------------------------------------------------
# config 5: one oversized volume split along the scan axis
# --------------------------------------------------------------------------------------------------
def scan_split_block(dist, rank: int, world: int, local: int, peak_gbs: float, *, shape=(4000, 300, 2048), reps: int = 5,
                     transports=("peer", "nccl", "vmm"), check: bool = True) -> dict:
    import torch

    import shrimpy_b200 as sb
    from shrimpy_b200 import sharding

    g = sb.deskew_geometry(shape, 30.0, 0.39, True, 1)
    shards = sharding.plan_scan_split(g, world)
    me = shards[rank]
    Yn, X, Xp = g.out_shape
    vin, vout = g.algorithmic_bytes
    alg_bytes = vin * 2 + vout * 4
    dev = torch.device("cuda", local)

    def slab_of(r):
        z0, z1 = shards[r].own_z
        gen = torch.Generator(device="cuda").manual_seed(1000 + r)
        return torch.randint(100, 60000, (z1 - z0,) + tuple(shape[1:]), dtype=torch.int32, device="cuda",
                             generator=gen).to(torch.uint16)

    res = {
        "workload": f"ONE volume {tuple(shape)} uint16, ls_angle 30, px_to_scan_ratio 0.39, keep_overhang=True, n=1 -> "
                    f"float32 {g.out_shape}; rank g holds a contiguous range of raw scan slices and computes a contiguous "
                    "range of output columns; the slices its columns read below/above its own come from the neighbours",
        "n_gpus": world, "algorithmic_bytes": alg_bytes,
        "columns_per_rank": [s.cols[1] - s.cols[0] for s in shards],
        "interior_columns_frac": round(sum(s.interior_cols[1] - s.interior_cols[0] for s in shards) / Xp, 4),
        "transports": {},
    }
    own = slab_of(rank)
    halo_bytes = (me.halo_below[1] - me.halo_below[0] + me.halo_above[1] - me.halo_above[0]) * shape[1] * shape[2] * 2
    res["halo_mb_max_per_rank"] = round(_max_over_ranks(dist, halo_bytes) / 1e6, 2)

    ref = None
    if check:
        # the single-GPU answer for this rank's columns: the same window of the FULL stack's geometry, computed from
        # the whole stack (every rank regenerates it from the per-slab seeds)
        full = torch.cat([own if r == rank else slab_of(r) for r in range(world)], dim=0)
        ref = sb.deskew_window(full, g, p_begin=0, p_count=Yn, c_begin=me.cols[0], c_count=me.cols[1] - me.cols[0],
                               y_origin=0, z_origin=0)
        del full
        torch.cuda.synchronize()

    def record(name, run, extra=None, close=None):
        try:
            run()                                    # warm-up (NCCL channels, first mappings)
            torch.cuda.synchronize()
            rx0 = _nvlink_rx_bytes(local)
            ms, piece = _timed(dist, run, reps)
            rx1 = _nvlink_rx_bytes(local)
            ok = None if ref is None else _all_true(dist, torch.equal(piece, ref))
            del piece
            entry = {"ms": round(ms, 4), "algorithmic_gbs_total": round(alg_bytes / ms / 1e6, 1),
                     "algorithmic_gbs_per_gpu": round(alg_bytes / world / ms / 1e6, 1),
                     "frac_of_hbm_peak_per_gpu": round(alg_bytes / world / ms / 1e6 / peak_gbs, 4),
                     "bit_equal_to_single_gpu_window": ok}
            if rx0 is not None and rx1 is not None:
                # max over ranks of what this GPU RECEIVED over NVLink per launch, by the driver's link counters
                # (includes the few KiB of the barrier / timing all-reduces between the repetitions)
                entry["nvlink_rx_mb_per_launch_max_rank_nvml"] = round(_max_over_ranks(dist, (rx1 - rx0) / reps) / 1e6, 2)
            if extra:
                entry.update(extra())
            res["transports"][name] = entry
        except Exception as exc:      # noqa: BLE001  (recorded, the other transports still run)
            res["transports"][name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        finally:
            if close:
                close()
        torch.cuda.empty_cache()

    if world == 1:
        record("single_gpu_window", lambda: sharding.deskew_scan_split(own, g, shards, 0))
        out = torch.empty(g.out_shape, dtype=torch.float32, device=dev)
        ms, _ = _timed(dist, lambda: sb.deskew_zyx(own, 30.0, 0.39, True, 1, out=out), reps)
        res["single_gpu_ms"] = round(ms, 4)
        res["single_gpu_frac_of_hbm_peak"] = round(alg_bytes / ms / 1e6 / peak_gbs, 4)
        del out
        return res

    # the single-GPU time of the same job on this box, so that the line carries its own speed-up
    if rank == 0:
        full = torch.cat([own] + [slab_of(r) for r in range(1, world)], dim=0)
        out = torch.empty(g.out_shape, dtype=torch.float32, device=dev)
        sb.deskew_zyx(full, 30.0, 0.39, True, 1, out=out)
        evs = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            sb.deskew_zyx(full, 30.0, 0.39, True, 1, out=out)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        single = sorted(a.elapsed_time(b) for a, b in evs)[reps // 2]
        del full, out
        torch.cuda.empty_cache()
    else:
        single = 0.0
    single = _max_over_ranks(dist, single)
    res["single_gpu_ms"] = round(single, 4)
    res["target_ms"] = round(single / world * 1.15, 4)

    if "peer" in transports:
        state = {}

        def setup_peer():
            peer = sharding.PeerSlab(shards, rank, tuple(shape[1:]), torch.uint16, dev)
            peer.tensor.copy_(own)
            torch.cuda.synchronize()
            peer.barrier()                 # every rank's slices are in place before anyone pulls
            torch.cuda.synchronize()
            state["peer"], state["side"] = peer, torch.cuda.Stream()

        def halo_only_peer():
            def pull():
                return sharding.exchange_halos_peer(state["peer"], shards, rank)
            pull()
            ms, _ = _timed(dist, pull, reps)
            return {"halo_pull_ms": round(ms, 4), "nvlink_gbs_max_rank": round(res["halo_mb_max_per_rank"] / ms, 1),
                    "data_path": "device copies out of the neighbours' peer-mapped (symmetric) memory, no NCCL call"}

        try:
            setup_peer()
            record("peer", lambda: sharding.deskew_scan_split(state["peer"], g, shards, rank, peer_stream=state["side"]),
                   extra=halo_only_peer)
        except Exception as exc:      # noqa: BLE001
            res["transports"]["peer"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        state.clear()
        torch.cuda.empty_cache()

    if "nccl" in transports:
        def halo_only_nccl():
            def xchg():
                low, high, reqs = sharding.exchange_halos(own, shards, rank)
                for r in reqs:
                    r.wait()
                return low, high
            xchg()
            ms, _ = _timed(dist, xchg, reps)
            return {"halo_exchange_ms": round(ms, 4), "nvlink_gbs_max_rank": round(res["halo_mb_max_per_rank"] / ms, 1),
                    "data_path": "torch.distributed batch_isend_irecv (NCCL send/recv pairs)"}

        record("nccl", lambda: sharding.deskew_scan_split(own, g, shards, rank), extra=halo_only_nccl)

    if "vmm" in transports:
        from shrimpy_b200 import paged_stack

        state = {}
        try:
            frame_bytes = shape[1] * shape[2] * 2
            page = paged_stack.PagedStack.granularity(local)
            pshards = paged_stack.plan_paged_split(g, world, frame_bytes, page)
            stack = paged_stack.PagedStack(pshards, rank, tuple(shape[1:]), torch.uint16, local, page)
            state["stack"] = stack

            def slices_of(z0, z1):     # the same per-rank slabs as the other transports, cut where the pages are
                parts = []
                for r in range(world):
                    a, b = shards[r].own_z
                    lo, hi = max(a, z0), min(b, z1)
                    if hi > lo:
                        parts.append((own if r == rank else slab_of(r))[lo - a:hi - a])
                return torch.cat(parts, dim=0)

            stack.fill_own(slices_of)
            stack.barrier()
            remote = _max_over_ranks(dist, pshards[rank].remote_bytes)
            record("vmm", lambda: paged_stack.deskew_paged_split(stack, g, pshards[rank]),
                   extra=lambda: {"remote_mb_mapped_max_rank": round(remote / 1e6, 2), "launches_per_rank": 1,
                                  "data_path": "no exchange step: the neighbours' pages are mapped next to the rank's "
                                               "own (cuMemMap) and the kernel's TMA tile loads read them over NVLink"})
        except Exception as exc:      # noqa: BLE001
            res["transports"]["vmm"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        finally:
            if "stack" in state:
                _barrier(dist)
                state["stack"].close()
        torch.cuda.empty_cache()

    good = {k: v["ms"] for k, v in res["transports"].items() if "ms" in v}
    if good:
        best = min(good, key=good.get)
        res["best_transport"], res["ms"] = best, good[best]
        res["speedup_vs_single_gpu"] = round(single / good[best], 3)
        res["meets_target"] = bool(good[best] <= res["target_ms"])
    return res


# --------------------------------------------------------------------------------------------------
# config 4: a plate streamed from OME-Zarr
# --------------------------------------------------------------------------------------------------
def _mem_available_bytes() -> int:
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 0


def plate_block(dist, rank: int, world: int, local: int, *, positions: int = 96, timepoints: int = 10,
                root: str = "/dev/shm/shrimpy_b200_bench_plate", store_cap_bytes: int = 96 << 30, depth: int = 4,
                io_threads: int = 0) -> dict:
    """960 (position, time) stacks of (600, 300, X) uint16 in an uncompressed zarr-v3 HCS plate on tmpfs, every rank
    writes then streams its own share (unit i goes to rank i % world).  X is the widest power of two <= 2048 for which
    the whole plate fits the RAM-backed store (the 707 GB of full (600, 300, 2048) stacks do not)."""
    import torch

    from shrimpy_b200 import plate, zarr_io
    from shrimpy_b200.settings import DeskewSettings

    Z, Y = 600, 300
    units = positions * timepoints
    if io_threads <= 0:       # the loader is a host-memory copy: give it the cores, shared between the ranks of the box
        io_threads = max(4, min(14, (os.cpu_count() or 8) // world - 2))
    try:
        free = shutil.disk_usage(os.path.dirname(root) or "/").free
    except OSError:
        free = 0
    budget = min(int(free * 0.45), int(_mem_available_bytes() * 0.30), store_cap_bytes)
    X = 2048
    while X > 32 and units * Z * Y * X * 2 > budget:
        X //= 2
    fits = units * Z * Y * X * 2 <= budget
    fits = _all_true(dist, fits)
    res = {"workload": f"{positions} positions x {timepoints} timepoints x 1 channel = {units} stacks of uint16 ({Z}, {Y}, {X}) "
                       "in a zarr-v3 / NGFF 0.5 HCS plate (chunks (1,1,512,Y,X), uncompressed) on tmpfs; "
                       "ls_angle 30, px_to_scan_ratio 0.39, keep_overhang=False, average_n_slices=3; "
                       "every stack: store -> pinned host -> GPU -> deskew -> pinned host result",
           "n_gpus": world, "units": units, "volume_shape": [Z, Y, X], "store": root,
           "store_bytes": units * Z * Y * X * 2,
           "why_this_x": f"widest power-of-two X whose {units}-stack plate fits min(45 % of the store's free space, 30 % of "
                         f"MemAvailable, {store_cap_bytes >> 30} GiB) = {budget / 2**30:.1f} GiB"}
    if not fits:
        res["error"] = "the RAM-backed store cannot hold the plate even at X = 32"
        return res
    Xs = _max_over_ranks(dist, X)
    if Xs != X:      # ranks see the same box; be safe anyway
        X = int(-_max_over_ranks(dist, -X))
    store = Path(root)
    try:
        if rank == 0:
            shutil.rmtree(store, ignore_errors=True)
            names = [f"{'ABCDEFGH'[i // 12]}/{i % 12 + 1}/fov0" for i in range(positions)]
            zarr_io.create_plate(store / "raw.zarr", names, (timepoints, 1, Z, Y, X), (1, 1, min(512, Z), Y, X),
                                 np.uint16, channel_names=["GFP"])
        _barrier(dist)
        src = zarr_io.open_plate(store / "raw.zarr")
        all_units = plate.list_units(src)
        mine = [u for i, u in enumerate(all_units) if i % world == rank]
        stack = np.random.default_rng(100 + rank).integers(100, 60000, size=(Z, Y, X), dtype=np.uint16)
        import threading
        from concurrent.futures import ThreadPoolExecutor
        private = threading.local()

        def write_unit(k):           # every writer thread stamps its own copy of the stack: every stack differs
            if not hasattr(private, "stack"):
                private.stack = stack.copy()
            i, t, c = mine[k]
            private.stack[0, 0, :8] = np.frombuffer(np.int64(k * world + rank).tobytes(), np.uint16).repeat(2)
            src[i].array.write_stack(t, c, private.stack)

        t0 = time.perf_counter()
        with ThreadPoolExecutor(max(1, io_threads)) as wpool:
            list(wpool.map(write_unit, range(len(mine))))
        write_s = time.perf_counter() - t0
        settings = DeskewSettings(ls_angle_deg=30.0, pixel_size_um=0.116, px_to_scan_ratio=0.39, keep_overhang=False,
                                  average_n_slices=3)
        plate.deskew_plate(src, settings, units=all_units[rank:rank + 1], depth=1)   # warm-up: pinned allocations, first launch
        checks = {}

        first = (src[mine[0][0]].name, mine[0][1], mine[0][2])

        def keep_one(name, t, c, array):          # parity spot check: the first unit of every rank, against the device path
            if (name, t, c) == first:
                checks["sum"] = float(array[::7, ::5, ::3].astype(np.float64).sum())

        _barrier(dist)
        stats = plate.deskew_plate(src, settings, rank=rank, world_size=world, depth=depth, io_threads=io_threads,
                                   on_result=keep_one)
        secs = _max_over_ranks(dist, stats.seconds)
        out_vox = _sum_over_ranks(dist, stats.out_voxels)
        raw_vox = _sum_over_ranks(dist, stats.raw_voxels)
        done = _sum_over_ranks(dist, stats.units)
        # spot check: the same unit through the device path
        import shrimpy_b200 as sb
        i, t, c = mine[0]
        host = np.empty((Z, Y, X), np.uint16)
        src[i].array.read_stack_into(t, c, host)
        dev = sb.deskew_zyx(torch.from_numpy(host).cuda(), 30.0, 0.39, False, 3).cpu().numpy()
        same = float(dev[::7, ::5, ::3].astype(np.float64).sum()) == checks.get("sum")
        res.update({
            "units_done": int(done), "seconds": round(secs, 3), "value_gvoxel_out_per_s": round(out_vox / secs / 1e9, 3),
            "gvoxel_in_per_s": round(raw_vox / secs / 1e9, 3), "per_gpu_gvoxel_out_per_s": round(out_vox / secs / 1e9 / world, 3),
            "host_device_gbs_total": round((raw_vox * 2 + out_vox * 4) / secs / 1e9, 1),
            "store_read_gbs_total": round(raw_vox * 2 / secs / 1e9, 1), "plate_write_s_max_rank": round(_max_over_ranks(dist, write_s), 2),
            "launches_rank0": int(stats.launches), "first_unit_matches_device_path": _all_true(dist, same),
            "depth": depth, "io_threads_per_rank": io_threads,
        })
    finally:
        _barrier(dist)
        if rank == 0:
            shutil.rmtree(store, ignore_errors=True)
    return res


# --------------------------------------------------------------------------------------------------
# end to end through the public API, and the box's copy-only floor
# --------------------------------------------------------------------------------------------------
def e2e_block(dist, rank: int, world: int, local: int, raws, outs, params, steps: int, numa_bound) -> dict:
    """``raws`` / ``outs``: the device-resident channels of the headline step (outs hold the device path's result)."""
    import torch

    import shrimpy_b200 as sb

    angle, ratio, keep, navg = params
    C = len(raws)
    raw_shape, out_shape = tuple(raws[0].shape), tuple(outs[0].shape)
    vox_in, vox_out = int(np.prod(raw_shape)), int(np.prod(out_shape))
    h2d_bytes, d2h_bytes = C * vox_in * 2, C * vox_out * 4
    h_raw = [torch.empty(raw_shape, dtype=torch.uint16).pin_memory() for _ in range(C)]
    h_out = [torch.empty(out_shape, dtype=torch.float32).pin_memory() for _ in range(C)]
    for c in range(C):
        h_raw[c].copy_(raws[c])
    np_raw = [t.numpy() for t in h_raw]
    np_out = [t.numpy() for t in h_out]
    device = f"cuda:{local}"

    def wall(fn, n):
        fn()
        _barrier(dist)
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        mine = time.perf_counter() - t0
        _barrier(dist)
        return _max_over_ranks(dist, mine) / n

    def value(sec_per_step):
        return world * C * vox_out / sec_per_step / 1e9

    # (1) the contract e2e: page-locked numpy arrays in and out
    def pinned_step():
        for c in range(C):
            sb.deskew_data(np_raw[c], angle, ratio, keep, navg, device=device, out=np_out[c])

    s_pinned = wall(pinned_step, steps)
    ok = bool(torch.equal(h_out[0], outs[0].cpu()))

    # (2) the copy-only floor of this box at this many ranks: the same bytes, the same buffers, no kernel --
    # every channel's H2D on one stream and D2H on another, all ranks at once
    d_raw = [torch.empty_like(r) for r in raws]
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

    def copies(up=True, down=True):
        for c in range(C):
            if up:
                with torch.cuda.stream(s_up):
                    d_raw[c].copy_(h_raw[c], non_blocking=True)
            if down:
                with torch.cuda.stream(s_down):
                    h_out[c].copy_(outs[c], non_blocking=True)
        torch.cuda.synchronize()

    s_floor = wall(copies, steps)
    s_up_only = wall(lambda: copies(True, False), max(2, steps // 2))
    s_down_only = wall(lambda: copies(False, True), max(2, steps // 2))
    del d_raw

    # (3) what a caller of the reference's deskew_data passes: an ordinary (pageable) array; the result is whatever
    # the call returns (scripts/measure_psf.py:239-246)
    pageable = [np.array(a, copy=True) for a in np_raw]
    got = {}

    def pageable_step():
        for c in range(C):
            got[c] = sb.deskew_data(pageable[c], angle, ratio, keep, navg, device=device)

    t0 = time.perf_counter()
    pageable_step()                   # the very first such call of the process: pins its staging ring and two results
    cold_ms = 1e3 * (time.perf_counter() - t0)
    pageable_step()                   # a caller that rebinds its result holds two blocks per channel for a moment
    s_page = wall(pageable_step, max(2, steps // 2))
    ok_page = bool(np.array_equal(got[0], np_out[0]))
    got.clear()

    # (4) the reference's online path (preprocessing.py:316 -> :408-413): the stack goes up, the result STAYS on
    # the device for the next step.  (a) as this package would be fed: uint16 from a pinned buffer, convert fused into
    # the kernel; (b) the reference's literal hand-over: torch.as_tensor(pageable uint16, device, float32)
    dev = torch.device(device)

    def online_step():
        for c in range(C):
            d = h_raw[c].to(dev, non_blocking=True)
            got[c] = sb.fast_deskew_zyx(raw_data=d, ls_angle_deg=angle, px_to_scan_ratio=ratio, keep_overhang=keep,
                                        average_n_slices=navg)
        torch.cuda.synchronize()

    s_online = wall(online_step, steps)
    ok_online = bool(torch.equal(got[0], outs[0]))
    got.clear()

    def handover_step():
        for c in range(C):
            d = torch.as_tensor(pageable[c], device=dev, dtype=torch.float32)
            got[c] = sb.fast_deskew_zyx(raw_data=d, ls_angle_deg=angle, px_to_scan_ratio=ratio, keep_overhang=keep,
                                        average_n_slices=navg)
        torch.cuda.synchronize()

    try:
        s_hand = wall(handover_step, max(2, steps // 2))
        ok_hand = bool(torch.equal(got[0], outs[0]))
    except (RuntimeError, TypeError) as exc:       # a torch build without the uint16 -> float32 copy
        s_hand, ok_hand = None, f"{type(exc).__name__}: {exc}"[:200]
    got.clear()
    torch.cuda.empty_cache()

    return {
        "value": value(s_pinned), "unit": "GVoxel/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
        "steps": steps, "ms_per_step": 1e3 * s_pinned,
        "api": "shrimpy_b200.deskew_data(numpy pinned) -> shrimpy_deskew_host (H2D | kernel | D2H on 3 streams)",
        "matches_device_path": ok, "numa_bound": numa_bound,
        "floor_ms": 1e3 * s_floor, "floor_frac": s_floor / s_pinned,
        "floor": {"what": "copy-only replay at this many ranks, all ranks at once: the step's H2D bytes on one stream and "
                          "D2H bytes on another from/to the same pinned buffers, no kernel; floor_frac = floor_ms / ms_per_step",
                  "both_ms": 1e3 * s_floor, "h2d_only_ms": 1e3 * s_up_only, "d2h_only_ms": 1e3 * s_down_only,
                  "host_device_gbs_all_ranks": world * (h2d_bytes + d2h_bytes) / s_floor / 1e9,
                  "h2d_only_gbs_per_rank": h2d_bytes / s_up_only / 1e9, "d2h_only_gbs_per_rank": d2h_bytes / s_down_only / 1e9},
        "pageable": {"value": value(s_page), "ms_per_step": 1e3 * s_page, "h2d_bytes_per_step": h2d_bytes,
                     "d2h_bytes_per_step": d2h_bytes, "matches_device_path": ok_page,
                     "first_call_ms": cold_ms,
                     "what": "ordinary np.ndarray in (as scripts/measure_psf.py:239-246 passes), the returned array out; the "
                             "pipeline gathers the pageable stack into its page-locked ring on host threads"},
        "online": {"value": value(s_online), "ms_per_step": 1e3 * s_online, "h2d_bytes_per_step": h2d_bytes,
                   "d2h_bytes_per_step": 0, "matches_device_path": ok_online,
                   "what": "uint16 pinned stack -> device -> fast_deskew_zyx, result left on the device "
                           "(preprocessing.py:316 -> :408-413 with the convert fused into the kernel)"},
        "online_reference_handover": {"value": value(s_hand) if s_hand else None, "ms_per_step": 1e3 * s_hand if s_hand else None,
                                      "h2d_bytes_per_step": 2 * h2d_bytes, "d2h_bytes_per_step": 0,
                                      "matches_device_path": ok_hand,
                                      "what": "the reference's literal hand-over: torch.as_tensor(pageable uint16, device, "
                                              "float32) then fast_deskew_zyx on the float32 tensor"},
    }


# --------------------------------------------------------------------------------------------------
# config 3: affine registration resample
# --------------------------------------------------------------------------------------------------
def touched_input_voxels(in_shape, M, out_shape) -> int:
    """How many input voxels the trilinear resample reads at least once (every tap of every output voxel that lies
    inside), counted exactly on the device plane by plane -- the input side of the algorithmic bytes (SURVEY.md 8d,
    cfg 3: only the touched input counts)."""
    import torch

    iz, iy, ix = in_shape
    oz, oy, ox = out_shape
    M = torch.as_tensor(np.asarray(M, dtype=np.float64), device="cuda")
    hit = torch.zeros(iz * iy * ix, dtype=torch.bool, device="cuda")
    o1 = torch.arange(oy, dtype=torch.float64, device="cuda")[:, None]
    o2 = torch.arange(ox, dtype=torch.float64, device="cuda")[None, :]
    for o0 in range(oz):
        c = [((M[a, 3] + o0 * M[a, 0]) + o1 * M[a, 1]) + o2 * M[a, 2] for a in range(3)]
        inside = ((c[0] >= 0) & (c[0] <= iz - 1) & (c[1] >= 0) & (c[1] <= iy - 1) & (c[2] >= 0) & (c[2] <= ix - 1))
        f = [torch.floor(v).to(torch.int64) for v in c]
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    z = torch.clamp(f[0] + dz, max=iz - 1)
                    y = torch.clamp(f[1] + dy, max=iy - 1)
                    x = torch.clamp(f[2] + dx, max=ix - 1)
                    hit[((z * iy + y) * ix + x)[inside]] = True
    return int(hit.sum().item())


def affine_block(peak_gbs: float) -> dict:
    """BASELINE.json configs[2]: float32 (107, 2048, 2048) resampled with a 4x4 matrix, device resident, CUDA events,
    median of 8 launches after 3 warm-ups (inputs + outputs of 2.4-3.6 GB exceed the L2)."""
    import torch

    from shrimpy_b200 import register

    shape = (107, 2048, 2048)
    vol = torch.randn(shape, device="cuda")
    a, b, c = np.deg2rad([2.0, 1.0, 3.0])
    Rz = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rx = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    Mg = np.eye(4)
    Mg[:3, :3] = Rz @ Ry @ Rx @ np.diag([1.03, 0.97, 1.1])
    Mg[:3, 3] = [0.4, -1.2, 2.3]
    M90 = np.array([[1.0, 0, 0, 3.5], [0, 0, -1.288, 2040.0], [0, 1.288, 0, -20.0], [0, 0, 0, 1]])
    M90t = M90.copy()
    M90t[0, 1:3] = [0.02, -0.015]
    M90t[1, 0], M90t[2, 0] = 0.03, -0.02
    cases = (("in_plane_identity_like", np.eye(4), shape, "affine_stream_kernel"),
             ("in_plane_rot90_x1.288_onto_deskewed_grid", M90, (100, 2048, 1279), "affine_stream_kernel (lanes along o1)"),
             ("general_rot_2_1_3_deg_aniso_scale", Mg, shape, "affine_tilt_kernel"),
             ("rot90_x1.288_with_tilt_onto_deskewed_grid", M90t, (100, 2048, 1279), "affine_tilt_kernel (lanes along o1)"))
    res = {}
    for name, M, oshape, kern in cases:
        out = torch.empty(oshape, device="cuda")
        for _ in range(3):
            register.affine_transform_zyx(vol, M, oshape, out=out)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
        for e0, e1 in ev:
            e0.record()
            register.affine_transform_zyx(vol, M, oshape, out=out)
            e1.record()
        torch.cuda.synchronize()
        ms = float(np.median([e0.elapsed_time(e1) for e0, e1 in ev]))
        inside = float((out != 0).float().mean())
        touched = touched_input_voxels(shape, M, oshape)
        nbytes = (touched + out.numel()) * 4
        whole = (vol.numel() + out.numel()) * 4
        res[name] = {"kernel": kern, "out_shape": list(oshape), "ms": ms, "gvoxel_out_per_s": out.numel() / ms / 1e6,
                     "touched_input_voxels": touched, "touched_input_frac": touched / vol.numel(),
                     "algorithmic_bytes": nbytes, "algorithmic_gbs": nbytes / ms / 1e6,
                     "frac_of_hbm_peak": nbytes / ms / 1e6 / peak_gbs,
                     "frac_if_whole_input_were_charged": whole / ms / 1e6 / peak_gbs, "inside_fraction": inside}
        del out
    return {"workload": "affine registration resample of a float32 label-free volume (107,2048,2048) with a 4x4 matrix "
                        "(BASELINE.json configs[2]); algorithmic bytes = (input voxels read at least once + output voxels) x 4 "
                        "(SURVEY.md 8d cfg 3: only the touched input counts)", "cases": res}
