"""Host logic of the multi-GPU layout, exercised on CPU: plans, and a world_size-2 gloo run of the
scan-axis split with halo exchange (the CUDA window kernel is replaced by a numpy stand-in that
follows the same window contract and asserts that every tap it reads is present in the slab)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import shrimpy_b200 as sb
from helpers import synthetic_stack
from oracle import deskew_oracle as o
from shrimpy_b200 import sharding


def test_shard_units_partition():
    units = [(p, t, c) for p in range(5) for t in range(3) for c in range(2)]
    parts = [sharding.shard_units(units, 4, r) for r in range(4)]
    assert sorted(sum(parts, [])) == sorted(units)
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        sharding.shard_units(units, 2, 2)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("keep", [True, False])
def test_scan_split_plan_covers_everything(world, keep):
    g = sb.deskew_geometry((4000, 300, 2048), 30.0, 0.39, keep, 1)      # BASELINE configs[4] geometry
    shards = sharding.plan_scan_split(g, world)
    Z = g.raw_shape[0]
    assert shards[0].cols[0] == 0 and shards[-1].cols[1] == g.out_shape[2]
    assert shards[0].own_z[0] == 0 and shards[-1].own_z[1] == Z
    for a, b in zip(shards[:-1], shards[1:]):
        assert a.cols[1] == b.cols[0] and a.own_z[1] == b.own_z[0]
    for s in shards:
        assert 0 <= s.need_z[0] <= s.need_z[1] <= Z
        assert s.cols[0] <= s.interior_cols[0] <= s.interior_cols[1] <= s.cols[1]
        if world > 1 and s.cols[1] > s.cols[0]:
            # halo below is about r*cos(theta)*(Y-1)+1 slices, above at most two
            assert s.halo_below[1] - s.halo_below[0] <= int(0.39 * np.cos(np.pi / 6) * 299) + 3
            assert s.halo_above[1] - s.halo_above[0] <= 2
    if world == 8:
        assert sum(s.interior_cols[1] - s.interior_cols[0] for s in shards) > 0.6 * g.out_shape[2]


def numpy_window(slab, g, p_begin, p_count, c_begin, c_count, y_origin, z_origin, cval):
    """CPU stand-in for shrimpy_deskew_window_device (same contract, same float64 geometry)."""
    slab = slab.numpy().astype(np.float64)
    Z, Y, X = g.raw_shape
    n = g.n_avg
    out = np.empty((p_count, X, c_count), dtype=np.float32)
    o2 = np.arange(c_begin, c_begin + c_count, dtype=np.float64)
    for p in range(p_begin, p_begin + p_count):
        planes = []
        for k in range(n):
            o0 = min(n * p + k, Y - 1)
            z_in = (g.shift + np.float64(o0) * g.m00) + o2 * g.m02
            inside = (z_in >= 0) & (z_in <= Z - 1)
            zc = np.where(inside, z_in, float(z_origin))
            z0 = np.floor(zc).astype(np.int64)
            z1 = np.minimum(z0 + 1, Z - 1)
            w = zc - z0
            y = Y - 1 - o0 - y_origin
            assert 0 <= y < slab.shape[1]
            if inside.any():
                assert z0[inside].min() - z_origin >= 0 and z1[inside].max() - z_origin < slab.shape[0], "halo missing"
            z0c = np.clip(z0 - z_origin, 0, slab.shape[0] - 1)
            z1c = np.clip(z1 - z_origin, 0, slab.shape[0] - 1)
            col = slab[:, y, ::-1]
            v = (1.0 - w)[None, :] * col[z0c, :].T + w[None, :] * col[z1c, :].T
            planes.append(np.where(inside[None, :], v, cval).astype(np.float32))
        out[p - p_begin] = np.mean(np.stack(planes), axis=0)
    return torch.from_numpy(out)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, params, outdir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shape, ang, r, keep, n = params
        raw = synthetic_stack(shape, seed=42)                      # every rank can regenerate the volume ...
        g = sb.deskew_geometry(shape, ang, r, keep, n)
        shards = sharding.plan_scan_split(g, world, align=8)
        me = shards[rank]
        own = torch.from_numpy(raw[me.own_z[0]:me.own_z[1]].copy())   # ... but only holds its own slices
        piece = sharding.deskew_scan_split(own, g, shards, rank, cval=-3.0, window_fn=numpy_window)
        np.save(os.path.join(outdir, f"piece{rank}.npy"), piece.numpy())
        # units sharding needs no communication at all: check the partition with one all_gather of counts
        units = list(range(23))
        mine = torch.tensor([len(sharding.shard_units(units, world, rank))])
        counts = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
        dist.all_gather(counts, mine)
        assert int(sum(counts)) == len(units)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("params", [((120, 13, 8), 30.0, 0.39, True, 1), ((150, 10, 8), 30.0, 0.39, False, 3)],
                         ids=["keep-n1", "crop-n3"])
def test_scan_split_world2_gloo_matches_single(tmp_path, params):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, params, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    shape, ang, r, keep, n = params
    whole = o.deskew_data(synthetic_stack(shape, seed=42), ang, r, keep, n, cval=-3.0)
    got = np.concatenate([np.load(tmp_path / f"piece{r}.npy") for r in range(world)], axis=2)
    assert got.shape == whole.shape
    assert np.array_equal(got == -3.0, whole == -3.0)
    assert np.max(np.abs(got - whole)) <= 2e-6 * float(whole.max() - whole.min())
