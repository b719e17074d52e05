"""Blosc-1 frame codec of the chunk loader (``csrc/blosc_frame.cu``, host code; CPU only).

The reference acquisition writes ``blosc-zstd`` inside zarr-v3 shards (``shrimpy/mantis/mantis_engine.py:474-481``,
``shrimpy/tests/test_mantis_integration.py:182-188``).  No blosc library exists offline, so the decoder is pinned
against frames assembled HERE, byte by byte, from the published container layout with numpy and the system's
libzstd / liblz4 / zlib -- an independent restatement, not the library's own encoder -- plus encode -> decode round
trips and the zarr layouts the acquisition uses.
"""

import ctypes
import json
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from shrimpy_b200 import _cabi, zarr_io

LZ4, ZLIB, ZSTD = 1, 3, 4


def _libs():
    out = {}
    try:
        z = ctypes.CDLL("libzstd.so.1")
        z.ZSTD_compress.restype = ctypes.c_size_t
        z.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        z.ZSTD_compressBound.restype = ctypes.c_size_t
        z.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
        out[ZSTD] = z
    except OSError:
        pass
    try:
        l = ctypes.CDLL("liblz4.so.1")
        l.LZ4_compress_default.restype = ctypes.c_int
        l.LZ4_compress_default.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        l.LZ4_compressBound.restype = ctypes.c_int
        l.LZ4_compressBound.argtypes = [ctypes.c_int]
        out[LZ4] = l
    except OSError:
        pass
    out[ZLIB] = zlib
    return out


LIBS = _libs()


def _stream(codec: int, raw: bytes) -> bytes:
    if codec == ZSTD:
        z = LIBS[ZSTD]
        cap = z.ZSTD_compressBound(len(raw))
        buf = ctypes.create_string_buffer(cap)
        n = z.ZSTD_compress(buf, cap, raw, len(raw), 1)
        return buf.raw[:n]
    if codec == LZ4:
        l = LIBS[LZ4]
        cap = l.LZ4_compressBound(len(raw))
        buf = ctypes.create_string_buffer(cap)
        n = l.LZ4_compress_default(raw, buf, len(raw), cap)
        return buf.raw[:n]
    return zlib.compress(raw, 1)


def _shuffle(block: bytes, ts: int, mode: int) -> bytes:
    a = np.frombuffer(block, np.uint8)
    ne = len(a) // ts
    if mode == 1 and ts > 1:
        return a[:ne * ts].reshape(ne, ts).T.tobytes() + a[ne * ts:].tobytes()
    if mode == 2 and len(a) >= ts:
        ne8 = ne - ne % 8
        bits = np.unpackbits(a[:ne8 * ts].reshape(ne8, ts), axis=1, bitorder="little")      # (element, bit of element)
        rows = np.packbits(bits.T, axis=1, bitorder="little")                                 # (bit, element / 8)
        return rows.tobytes() + a[ne8 * ts:].tobytes()
    return block


def assemble(data: bytes, ts: int, codec: int, shuffle: int, blocksize: int, split: bool) -> bytes:
    """A blosc-1 frame from the container layout: header | bstarts | per block: (int32 csize | stream) x nsplits."""
    flags = (codec << 5) | (0 if split else 0x10) | (1 if shuffle == 1 and ts > 1 else 0) | (4 if shuffle == 2 else 0)
    nblocks = -(-len(data) // blocksize)
    body, starts = b"", []
    for b in range(nblocks):
        block = data[b * blocksize:(b + 1) * blocksize]
        leftover = len(block) != blocksize
        filtered = _shuffle(block, ts, shuffle)
        nsplits = ts if (split and ts <= 16 and blocksize // ts >= 128 and not leftover) else 1
        ne = len(block) // nsplits
        starts.append(16 + 4 * nblocks + len(body))
        for j in range(nsplits):
            part = filtered[j * ne:(j + 1) * ne]
            comp = _stream(codec, part)
            if len(comp) >= len(part):
                comp = part                                  # stored stream: csize == length
            body += struct.pack("<i", len(comp)) + comp
    head = bytes([2, 1, flags, ts]) + struct.pack("<iii", len(data), blocksize, 16 + 4 * nblocks + len(body))
    return head + struct.pack(f"<{nblocks}i", *starts) + body


def decode(frame: bytes, nbytes: int, threads: int = 1) -> bytes:
    out = np.empty(nbytes, np.uint8)
    src = np.frombuffer(frame, np.uint8)
    _cabi.check(_cabi.lib().shrimpy_blosc_decode(src.ctypes.data, src.nbytes, out.ctypes.data, out.nbytes, threads))
    return out.tobytes()


def encode(data: np.ndarray, codec=ZSTD, level=1, shuffle=1, blocksize=0, split=0) -> bytes:
    lib = _cabi.lib()
    cap = lib.shrimpy_blosc_encode_bound(data.nbytes, blocksize, data.dtype.itemsize)
    buf = np.empty(cap, np.uint8)
    n = ctypes.c_size_t(0)
    _cabi.check(lib.shrimpy_blosc_encode(data.ctypes.data, data.nbytes, data.dtype.itemsize, codec, level, shuffle, blocksize,
                                         split, buf.ctypes.data, cap, ctypes.byref(n)))
    return buf[:n.value].tobytes()


def _camera_like(n, dtype, seed=0):
    rng = np.random.default_rng(seed)
    smooth = 400 + 300 * np.sin(np.arange(n) / 37.0) + rng.normal(0, 6, n)
    return smooth.astype(dtype)


@pytest.mark.parametrize("codec", [ZSTD, LZ4, ZLIB])
@pytest.mark.parametrize("ts,dtype", [(2, np.uint16), (4, np.float32), (1, np.uint8), (8, np.float64)])
@pytest.mark.parametrize("shuffle", [0, 1, 2])
@pytest.mark.parametrize("split", [False, True])
def test_decodes_hand_assembled_frames(codec, ts, dtype, shuffle, split):
    if codec not in LIBS:
        pytest.skip("codec library not present")
    data = _camera_like(5000, dtype, seed=ts).tobytes() + b"\x07" * 3         # ragged: last block short, 3 stray bytes
    for blocksize in (1024 * ts, 4096 * ts + 8 * ts, len(data) // ts * ts):    # blosc keeps blocks whole elements
        frame = assemble(data, ts, codec, shuffle, blocksize, split)
        assert decode(frame, len(data)) == data
        assert decode(frame, len(data), threads=4) == data


def test_known_answer_frame_bytes():
    """A frame small enough to check by eye: 4 uint16 values, byte shuffle, every stream stored (incompressible)."""
    values = np.array([0x0102, 0x0304, 0x0506, 0x0708], "<u2")
    flags = (ZSTD << 5) | 0x10 | 0x1
    stored = bytes([0x02, 0x04, 0x06, 0x08, 0x01, 0x03, 0x05, 0x07])           # low bytes, then high bytes
    frame = bytes([2, 1, flags, 2]) + struct.pack("<iii", 8, 8, 16 + 4 + 4 + 8) + struct.pack("<i", 20) + \
        struct.pack("<i", 8) + stored
    assert decode(frame, 8) == values.tobytes()
    memcpyed = bytes([2, 1, flags | 0x2, 2]) + struct.pack("<iii", 8, 8, 24) + values.tobytes()
    assert decode(memcpyed, 8) == values.tobytes()
    info = [ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()]
    _cabi.check(_cabi.lib().shrimpy_blosc_info(frame, len(frame), *[ctypes.byref(v) for v in info]))
    assert [v.value for v in info] == [8, 32, 8, 2, flags]


@pytest.mark.parametrize("codec", [ZSTD, LZ4, ZLIB])
@pytest.mark.parametrize("shuffle", [0, 1, 2])
@pytest.mark.parametrize("split", [0, 1])
def test_encoder_round_trips_and_matches_the_assembler(codec, shuffle, split):
    if codec not in LIBS:
        pytest.skip("codec library not present")
    for dtype, n in ((np.uint16, 70001), (np.float32, 9000), (np.uint8, 333), (np.uint16, 0), (np.uint16, 1)):
        data = _camera_like(n, dtype, seed=n)
        for blocksize in (0, 4096):
            frame = encode(data, codec, 1, shuffle, blocksize, split)
            assert decode(frame, data.nbytes, threads=3) == data.tobytes()
    data = _camera_like(6000, np.uint16)
    if codec == ZSTD:           # same library, same level, same layout -> the two writers agree byte for byte
        assert encode(data, ZSTD, 1, shuffle, 4096, split) == assemble(data.tobytes(), 2, ZSTD, shuffle, 4096, bool(split))


def test_incompressible_data_becomes_a_stored_frame():
    data = np.random.default_rng(1).integers(0, 2**16, 5000, dtype=np.uint16)
    frame = encode(data, ZSTD, 1, 0)
    assert frame[2] & 0x2 and len(frame) == 16 + data.nbytes and decode(frame, data.nbytes) == data.tobytes()


def test_corrupt_frames_fail_with_a_message():
    data = _camera_like(4096, np.uint16)
    frame = bytearray(encode(data, ZSTD, 1, 1, 2048))
    lib = _cabi.lib()
    out = np.empty(data.nbytes, np.uint8)

    def run(buf, nbytes=data.nbytes):
        src = np.frombuffer(bytes(buf), np.uint8)
        return lib.shrimpy_blosc_decode(src.ctypes.data, src.nbytes, out.ctypes.data, nbytes, 2)

    assert run(frame) == 0
    assert run(frame[:10]) == _cabi.EINVAL and b"header" in lib.shrimpy_last_error()
    assert run(frame[:len(frame) // 2]) == _cabi.EINVAL                         # cbytes beyond the buffer
    assert run(frame, data.nbytes - 2) == _cabi.EINVAL and b"expects" in lib.shrimpy_last_error()
    bad = bytearray(frame); bad[0] = 3
    assert run(bad) == _cabi.EINVAL and b"version" in lib.shrimpy_last_error()
    bad = bytearray(frame); bad[16:20] = struct.pack("<i", len(frame) + 5)      # block start outside the frame
    assert run(bad) == _cabi.EINVAL
    bad = bytearray(frame); bad[40:60] = bytes(20)                              # garbage inside a zstd stream
    assert run(bad) == _cabi.EINVAL and b"did not decode" in lib.shrimpy_last_error()
    bad = bytearray(frame); bad[2] = (bad[2] & 0x1F) | (0 << 5)                 # blosclz stream: named, not guessed
    assert run(bad) == _cabi.EINVAL and b"blosclz" in lib.shrimpy_last_error()
    # header fields are untrusted: sizes near 2^31 must be rejected, not wrapped (int32 arithmetic used to let a
    # blocksize of INT32_MAX through with a success code)
    bad = bytearray(frame); bad[8:12] = struct.pack("<i", 2**31 - 1)
    assert run(bad) == _cabi.EINVAL and b"blocksize" in lib.shrimpy_last_error()
    bad = bytearray(frame); bad[8:12] = struct.pack("<i", data.nbytes + 2)
    assert run(bad) == _cabi.EINVAL and b"blocksize" in lib.shrimpy_last_error()
    bad = bytearray(frame); bad[4:8] = struct.pack("<i", 2**31 - 8); bad[2] |= 0x2   # "stored" frame of ~2 GiB
    assert run(bad, 2**31 - 8) == _cabi.EINVAL


def _acquisition_like(tmp_path, shape, shard, inner, **kw):
    return zarr_io.ZarrArray.create(tmp_path / "acq", shape, shard, np.uint16,
                                    blosc={"cname": "zstd", "clevel": 1, "shuffle": "shuffle"}, shard_inner=inner,
                                    dimension_names=("t", "c", "z", "y", "x"), **kw)


def test_sharded_blosc_zstd_store_round_trips(tmp_path):
    """The acquisition's layout: sharding_indexed -> bytes -> blosc(zstd, shuffle), z-chunked; Z not a multiple of the
    shard depth, so the last shard is cut short in the stack."""
    shape, shard, inner = (2, 2, 21, 12, 40), (1, 1, 8, 12, 40), (1, 1, 4, 12, 40)
    data = _camera_like(int(np.prod(shape)), np.uint16).reshape(shape)
    arr = _acquisition_like(tmp_path, shape, shard, inner)
    for t in range(2):
        for c in range(2):
            arr.write_stack(t, c, data[t, c])
    meta = json.loads((tmp_path / "acq" / "zarr.json").read_text())
    sharding = meta["codecs"][0]
    assert sharding["name"] == "sharding_indexed" and sharding["configuration"]["chunk_shape"] == list(inner)
    blosc = sharding["configuration"]["codecs"][1]
    assert blosc["name"] == "blosc" and blosc["configuration"]["cname"] == "zstd" and blosc["configuration"]["typesize"] == 2
    again = zarr_io.ZarrArray.open(tmp_path / "acq")
    assert again.codec.kind == "blosc" and again.shard_inner == inner and again.grid == (2, 2, 3, 1, 1)
    shard_file = (tmp_path / "acq" / "c" / "0" / "0" / "0" / "0" / "0").read_bytes()
    index = shard_file[-(2 * 16 + 4):]
    assert zarr_io.crc32c(index[:-4]) == int.from_bytes(index[-4:], "little")
    assert len(shard_file) < 8 * 12 * 40 * 2                                   # it did compress
    out = np.empty(shape[2:], np.uint16)
    with ThreadPoolExecutor(4) as pool:
        for t in range(2):
            for c in range(2):
                out[:] = 0
                again.read_stack_into(t, c, out, pool=pool if t else None)
                assert np.array_equal(out, data[t, c])


def test_shard_with_inner_chunks_tiling_y_and_x(tmp_path):
    """Inner chunks smaller than the frame (a writer's default x/y chunking): decoded through scratch, same bytes."""
    shape, shard, inner = (1, 1, 6, 8, 32), (1, 1, 6, 8, 32), (1, 1, 3, 4, 16)
    data = _camera_like(int(np.prod(shape)), np.uint16).reshape(shape)
    arr = _acquisition_like(tmp_path, shape, shard, inner)
    arr.write_stack(0, 0, data[0, 0])
    out = np.empty(shape[2:], np.uint16)
    zarr_io.ZarrArray.open(tmp_path / "acq").read_stack_into(0, 0, out)
    assert np.array_equal(out, data[0, 0])


def test_unsharded_blosc_chunks_and_crc32c_vector(tmp_path):
    arr = zarr_io.ZarrArray.create(tmp_path / "b", (1, 1, 10, 6, 16), (1, 1, 4, 6, 16), np.float32,
                                   blosc={"cname": "lz4", "clevel": 5, "shuffle": "bitshuffle"})
    data = _camera_like(10 * 6 * 16, np.float32).reshape(10, 6, 16)
    arr.write_stack(0, 0, data)
    out = np.empty((10, 6, 16), np.float32)
    zarr_io.ZarrArray.open(tmp_path / "b").read_stack_into(0, 0, out)
    assert np.array_equal(out, data)
    assert zarr_io.crc32c(b"123456789") == 0xE3069283                           # the CRC-32C check value


def test_mutated_frames_never_crash_the_decoder():
    """Byte flips and truncations of valid frames: every call returns (0 or an error code), none takes the process down."""
    lib = _cabi.lib()
    rng = np.random.default_rng(0)
    data = _camera_like(20000, np.uint16)
    frames = [encode(data, c, 1, s, bs, sp) for c in (ZSTD, LZ4, ZLIB) if c in LIBS for s in (0, 1, 2)
              for bs in (0, 2048) for sp in (0, 1)]
    out = np.empty(data.nbytes + 64, np.uint8)
    rejected = 0
    for it in range(3000):
        f = bytearray(frames[it % len(frames)])
        for _ in range(int(rng.integers(1, 6))):
            pos = int(rng.integers(0, min(len(f), 200) if rng.random() < 0.7 else len(f)))
            f[pos] = int(rng.integers(0, 256))
        if rng.random() < 0.2:
            f = f[:int(rng.integers(16, len(f)))]
        src = np.frombuffer(bytes(f), np.uint8)
        nbytes = ctypes.c_int64()
        rc = lib.shrimpy_blosc_info(src.ctypes.data, len(f), ctypes.byref(nbytes), None, None, None, None)
        n = nbytes.value if rc == 0 and 0 <= nbytes.value <= out.nbytes else data.nbytes
        rc = lib.shrimpy_blosc_decode(src.ctypes.data, len(f), out.ctypes.data, n, int(rng.integers(1, 5)))
        assert rc in (_cabi.OK, _cabi.EINVAL, _cabi.ENOMEM)
        rejected += rc != 0
    assert rejected > 1000
