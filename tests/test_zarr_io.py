"""Minimal Zarr v3 / NGFF 0.5 reader-writer used by the streaming loader.  CPU only."""

import json
import zlib

import numpy as np
import pytest

from shrimpy_b200 import plate as plate_mod
from shrimpy_b200 import zarr_io
from shrimpy_b200.settings import DeskewSettings


def _pattern(shape, p=0):
    """Pixel value encodes (p, t, c, z) like the reference's replay-camera fixture (tests/test_replay_camera.py:33-47)."""
    T, C, Z, Y, X = shape
    t, c, z = np.meshgrid(np.arange(T), np.arange(C), np.arange(Z), indexing="ij")
    base = (p * 1000 + t * 100 + c * 10 + z).astype(np.uint16)
    return np.broadcast_to(base[..., None, None], shape).copy() + np.arange(X, dtype=np.uint16)


@pytest.mark.parametrize("zstd", [None, 3])
def test_array_roundtrip_and_partial_last_chunk(tmp_path, zstd):
    if zstd is not None and not zarr_io.zstd_available():
        pytest.skip("libzstd.so.1 not present")
    shape, chunks = (2, 2, 11, 6, 8), (1, 1, 4, 6, 8)          # 11 % 4 != 0 -> last z-chunk is partial
    data = _pattern(shape)
    arr = zarr_io.ZarrArray.create(tmp_path / "a", shape, chunks, np.uint16, zstd_level=zstd,
                                   dimension_names=("t", "c", "z", "y", "x"))
    for t in range(2):
        for c in range(2):
            arr.write_stack(t, c, data[t, c])
    again = zarr_io.ZarrArray.open(tmp_path / "a")
    assert again.shape == shape and again.chunks == chunks and again.dtype == np.uint16 and again.zstd == (zstd is not None)
    assert again.grid == (2, 2, 3, 1, 1)
    meta = json.loads((tmp_path / "a" / "zarr.json").read_text())
    assert meta["zarr_format"] == 3 and meta["node_type"] == "array"
    assert (tmp_path / "a" / "c" / "1" / "0" / "2" / "0" / "0").exists()   # default key encoding
    out = np.empty(shape[2:], dtype=np.uint16)
    for t in range(2):
        for c in range(2):
            again.read_stack_into(t, c, out)
            assert np.array_equal(out, data[t, c])


def test_missing_chunk_reads_fill_value_and_bad_buffers_raise(tmp_path):
    arr = zarr_io.ZarrArray.create(tmp_path / "a", (1, 1, 8, 4, 4), (1, 1, 4, 4, 4), np.uint16, fill_value=7)
    arr.write_chunk((0, 0, 0, 0, 0), np.ones((1, 1, 4, 4, 4), np.uint16))
    out = np.zeros((8, 4, 4), np.uint16)
    arr.read_stack_into(0, 0, out)
    assert np.all(out[:4] == 1) and np.all(out[4:] == 7)
    with pytest.raises(ValueError):
        arr.read_stack_into(0, 0, np.zeros((8, 4, 4), np.float32))
    with pytest.raises(ValueError):
        arr.read_stack_into(0, 0, np.zeros((8, 4, 5), np.uint16))


def test_unsupported_codecs_are_named(tmp_path):
    arr = zarr_io.ZarrArray.create(tmp_path / "a", (1, 1, 4, 4, 4), (1, 1, 4, 4, 4), np.uint16)
    meta = json.loads((tmp_path / "a" / "zarr.json").read_text())
    for tail, word in (({"name": "gzip", "configuration": {"level": 5}}, "gzip"),
                       ({"name": "blosc", "configuration": {"cname": "blosclz", "clevel": 1, "shuffle": "shuffle"}}, "blosclz"),
                       ({"name": "blosc", "configuration": {"cname": "snappy", "clevel": 1, "shuffle": "shuffle"}}, "snappy")):
        meta["codecs"] = [{"name": "bytes", "configuration": {"endian": "little"}}, tail]
        (tmp_path / "a" / "zarr.json").write_text(json.dumps(meta))
        with pytest.raises(NotImplementedError, match=word):
            zarr_io.ZarrArray.open(tmp_path / "a")


@pytest.mark.parametrize("index_at_end", [True, False])
def test_reads_sharding_indexed(tmp_path, index_at_end):
    """A hand-built shard (inner chunks + (offset, nbytes) index + crc32c) as the reference writer lays it out."""
    shape, shard, inner = (1, 1, 8, 4, 6), (1, 1, 8, 4, 6), (1, 1, 2, 4, 6)
    data = _pattern(shape)
    arr = zarr_io.ZarrArray.create(tmp_path / "s", shape, shard, np.uint16)
    meta = json.loads((tmp_path / "s" / "zarr.json").read_text())
    meta["codecs"] = [{"name": "sharding_indexed", "configuration": {
        "chunk_shape": list(inner), "codecs": [{"name": "bytes", "configuration": {"endian": "little"}}],
        "index_codecs": [{"name": "bytes", "configuration": {"endian": "little"}}, {"name": "crc32c"}],
        "index_location": "end" if index_at_end else "start"}}]
    (tmp_path / "s" / "zarr.json").write_text(json.dumps(meta))
    n_inner = 4
    index_bytes = n_inner * 16 + 4
    body, index = b"", []
    start = 0 if index_at_end else index_bytes
    for k in (2, 0, 3):                                        # out of order, inner chunk 1 missing
        payload = data[:, :, 2 * k:2 * k + 2].tobytes()
        index.append((k, start + len(body), len(payload)))
        body += payload
    table = np.full((n_inner, 2), 2**64 - 1, dtype="<u8")
    for k, off, n in index:
        table[k] = (off, n)
    raw_index = table.tobytes() + zarr_io.crc32c(table.tobytes()).to_bytes(4, "little")
    path = arr.chunk_path((0, 0, 0, 0, 0))
    path.parent.mkdir(parents=True)
    path.write_bytes(body + raw_index if index_at_end else raw_index + body)
    sharded = zarr_io.ZarrArray.open(tmp_path / "s")
    assert sharded.shard_inner == inner
    out = np.empty((8, 4, 6), np.uint16)
    sharded.read_stack_into(0, 0, out)
    want = data[0, 0].copy()
    want[2:4] = 0                                              # the missing inner chunk reads as fill_value
    assert np.array_equal(out, want)
    blob = bytearray(path.read_bytes())                        # a damaged index is refused (crc32c index codec)
    blob[-10 if index_at_end else 5] ^= 0x40
    path.write_bytes(bytes(blob))
    with pytest.raises(IOError, match="checksum"):
        sharded.read_stack_into(0, 0, out)


def test_plate_metadata_roundtrip_and_units(tmp_path):
    names = ["A/1/fov0", "A/1/fov1", "B/2/fov0"]
    positions = zarr_io.create_plate(tmp_path / "p.zarr", names, (2, 2, 9, 4, 8), (1, 1, 4, 4, 8), np.uint16,
                                     channel_names=["BF", "GFP"], scale=(1, 1, 0.174, 0.1133, 0.1133))
    for i, pos in enumerate(positions):
        data = _pattern(pos.array.shape, p=i)
        for t in range(2):
            for c in range(2):
                pos.array.write_stack(t, c, data[t, c])
    opened = zarr_io.open_plate(tmp_path / "p.zarr")
    assert [p.name for p in opened] == names
    assert opened[0].channel_names == ("BF", "GFP") and opened[0].scale[2:] == (0.174, 0.1133, 0.1133)
    plate_meta = json.loads((tmp_path / "p.zarr" / "zarr.json").read_text())["attributes"]["ome"]
    assert plate_meta["version"] == "0.5" and len(plate_meta["plate"]["wells"]) == 2
    out = np.empty((9, 4, 8), np.uint16)
    opened[2].array.read_stack_into(1, 0, out)
    assert np.array_equal(out, _pattern((2, 2, 9, 4, 8), p=2)[1, 0])
    units = plate_mod.list_units(opened)
    assert len(units) == 3 * 2 * 2 and units[0] == (0, 0, 0) and units[-1] == (2, 1, 1)
    single = zarr_io.open_plate(tmp_path / "p.zarr" / "A" / "1" / "fov1")       # a lone FOV group also opens
    assert len(single) == 1 and single[0].array.shape == (2, 2, 9, 4, 8)


def test_deskewed_plate_layout(tmp_path):
    src = zarr_io.create_plate(tmp_path / "raw.zarr", ["A/1/fov0"], (1, 2, 60, 12, 16), (1, 1, 32, 12, 16), np.uint16)
    s = DeskewSettings(ls_angle_deg=30, pixel_size_um=0.116, px_to_scan_ratio=0.39, keep_overhang=True,
                       average_n_slices=3)
    dst = plate_mod.create_deskewed_plate(tmp_path / "dsk.zarr", src, s, z_chunk=50)
    arr = dst[0].array
    assert arr.shape == (1, 2, 4, 16, 165) and arr.dtype == np.float32
    assert arr.chunks == (1, 1, 4, 16, 165)
    assert dst[0].scale == pytest.approx((1, 1, 3 * 0.5 * 0.116, 0.116, 0.116))


@pytest.mark.parametrize("layout", ["raw", "zstd", "sharded"])
def test_chunks_that_tile_y_and_x(tmp_path, layout):
    """A writer's default chunking may split the frame: (1, 1, 4, 5, 8) chunks of a (1, 1, 10, 12, 20) array, with
    partial chunks on every axis; hand-written chunk files (the writer here only emits z-slab chunks)."""
    if layout == "zstd" and not zarr_io.zstd_available():
        pytest.skip("libzstd.so.1 not present")
    shape, chunks = (1, 1, 10, 12, 20), (1, 1, 4, 5, 8)
    data = _pattern(shape) + (np.arange(12, dtype=np.uint16) * 1000)[None, None, None, :, None]
    inner = (1, 1, 2, 5, 4) if layout == "sharded" else None
    codec = zarr_io.Codec("zstd", 3) if layout == "zstd" else zarr_io.Codec()
    arr = zarr_io.ZarrArray.create(tmp_path / "a", shape, chunks, np.uint16, zstd_level=3 if layout == "zstd" else None,
                                   shard_inner=inner)
    for kz in range(3):
        for ky in range(3):
            for kx in range(3):
                block = np.zeros(chunks, np.uint16)
                part = data[:, :, kz * 4:(kz + 1) * 4, ky * 5:(ky + 1) * 5, kx * 8:(kx + 1) * 8]
                block[tuple(slice(0, n) for n in part.shape)] = part
                if layout == "sharded":
                    arr.write_chunk((0, 0, kz, ky, kx), block)          # shard writer: inner chunks + index + crc32c
                else:
                    path = arr.chunk_path((0, 0, kz, ky, kx))
                    path.parent.mkdir(parents=True, exist_ok=True)
                    path.write_bytes(bytes(codec.encode(block)))
    out = np.empty(shape[2:], np.uint16)
    opened = zarr_io.ZarrArray.open(tmp_path / "a")
    opened.read_stack_into(0, 0, out)
    assert np.array_equal(out, data[0, 0])
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(3) as pool:
        out[:] = 0
        opened.read_stack_into(0, 0, out, pool=pool)
    assert np.array_equal(out, data[0, 0])
