"""Streaming OME-Zarr plate deskew (pinned loader + 3 streams) against the oracle, on a small plate."""

import numpy as np
import pytest

from helpers import TIGHT_TOL, assert_close_range, synthetic_stack

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["zstd", "acquisition"])
def small_plate(tmp_path, request):
    """``acquisition``: the reference's own storage layout -- blosc(zstd, shuffle) inside sharding_indexed shards
    (shrimpy/mantis/mantis_engine.py:474-481), 40 slices in shards of 32 with inner chunks of 8."""
    from shrimpy_b200 import zarr_io

    names = ["A/1/fov0", "A/2/fov0", "B/1/fov0"]
    shape = (2, 2, 40, 12, 64)
    if request.param == "acquisition":
        layout = dict(chunks=(1, 1, 32, 12, 64), blosc={"cname": "zstd", "clevel": 1, "shuffle": "shuffle"},
                      shard_inner=(1, 1, 8, 12, 64))
    else:
        layout = dict(chunks=(1, 1, 16, 12, 64), zstd_level=3 if zarr_io.zstd_available() else None)
    positions = zarr_io.create_plate(tmp_path / "raw.zarr", names, shape, layout.pop("chunks"), np.uint16,
                                     channel_names=["BF", "GFP"], scale=(1, 1, 0.3, 0.116, 0.116), **layout)
    truth = {}
    for i, pos in enumerate(positions):
        for t in range(2):
            for c in range(2):
                raw = synthetic_stack(shape[2:], seed=100 * i + 10 * t + c)
                pos.array.write_stack(t, c, raw)
                truth[(pos.name, t, c)] = raw
    return tmp_path, truth


def test_plate_streaming_matches_oracle(small_plate):
    from oracle import deskew_oracle as o
    from shrimpy_b200 import plate, zarr_io
    from shrimpy_b200.settings import DeskewSettings

    root, truth = small_plate
    src = zarr_io.open_plate(root / "raw.zarr")
    settings = DeskewSettings(ls_angle_deg=30.0, pixel_size_um=0.116, scan_step_um=0.3, keep_overhang=False,
                              average_n_slices=3)
    sharded = src[0].array.shard_inner is not None           # write back in the layout the plate came in
    dst = plate.create_deskewed_plate(root / "deskewed.zarr", src, settings, z_chunk=3,
                                      blosc={"cname": "zstd", "clevel": 1, "shuffle": "shuffle"} if sharded else None,
                                      shard_z=6 if sharded else None)
    seen = []
    stats = plate.deskew_plate(src, settings, dst, depth=3, io_threads=3,
                               on_result=lambda name, t, c, arr: seen.append((name, t, c, float(arr.sum()))))
    assert stats.units == 12 and stats.launches == 12 and len(seen) == 12
    assert stats.h2d_bytes == 12 * 40 * 12 * 64 * 2
    reopened = zarr_io.open_plate(root / "deskewed.zarr")
    args = (settings.ls_angle_deg, settings.px_to_scan_ratio, settings.keep_overhang, settings.average_n_slices)
    for pos in reopened:
        out = np.empty(pos.array.shape[2:], np.float32)
        for t in range(2):
            for c in range(2):
                pos.array.read_stack_into(t, c, out)
                want = o.deskew_data(truth[(pos.name, t, c)], *args)
                assert_close_range(out, want, TIGHT_TOL, f"{pos.name} t{t} c{c}")


def test_plate_rank_sharding_covers_all_units_once(small_plate):
    from shrimpy_b200 import plate, zarr_io

    root, _ = small_plate
    src = zarr_io.open_plate(root / "raw.zarr")
    settings = {"ls_angle_deg": 30.0, "pixel_size_um": 0.116, "scan_step_um": 0.3, "keep_overhang": True,
                "average_n_slices": 1}
    done = []
    for rank in range(2):           # the two ranks of a world of 2, run one after the other on this GPU
        stats = plate.deskew_plate(src, settings, rank=rank, world_size=2, depth=2)
        done += stats.per_unit
        assert stats.units == 6
    assert sorted(done) == sorted((p.name, t, c) for p in src for t in range(2) for c in range(2))
