"""Minimal Zarr v3 / OME-NGFF 0.5 reader-writer for the streaming deskew (no ``zarr``/``iohub`` needed).

shrimPy writes and replays OME-Zarr: HCS plates ``row/col/fov/0`` holding ``TCZYX`` arrays, z-chunked
(``shrimpy/mantis/mantis_engine.py:474-481`` ``chunk_size=min(512, nz)``; ``scripts/measure_psf.py:273-287``
writes ``chunks=(1, 1, 50, Y, X)``; ``shrimpy/replay_camera.py:176-204, 293-308`` reads one ``(t, c)``
volume at a time).  This module implements exactly what the loader needs from that format:

* arrays: regular chunk grid, ``default`` chunk-key encoding, codecs ``bytes`` (little endian) optionally
  followed by ``zstd`` (through ``libzstd.so.1`` via ctypes when the library is present) or by ``blosc``
  (the acquisition's own ``blosc-zstd``; framing undone by the native ``shrimpy_blosc_decode`` of the
  C-ABI library, ``csrc/blosc_frame.cu``), bare or inside ``sharding_indexed`` shards (read and write);
* groups: plate / well / image metadata of NGFF 0.5 (``attributes.ome``), enough to enumerate positions,
  channel names and the scale transform, and to write a deskewed plate back.

The reference acquisition compresses with **blosc-zstd** inside shards
(``shrimpy/mantis/mantis_engine.py:474-481``, ``shrimpy/tests/test_mantis_integration.py:182-188``); that
layout is read (shards are memory-mapped, inner chunks decoded concurrently straight into the pinned stack)
and can be written (``ZarrArray.create(..., blosc={...}, shard_inner=...)``).  No blosc library exists
offline, so the frame codec is pinned by hand-assembled frames and round trips only (stated in
``csrc/blosc_frame.cu``); blosclz / snappy streams raise, naming the codec.

A chunk of a ``(1, 1, Zc, Y, X)`` grid is one contiguous z-slab of the ``(Z, Y, X)`` stack, so
``read_stack_into`` lands every chunk file directly at its offset of a caller-provided (pinned) buffer --
no intermediate copy for uncompressed data, one decompress-into-place for zstd.
"""

from __future__ import annotations

import ctypes
import ctypes.util
import json
import mmap
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

__all__ = ["ZarrArray", "Codec", "Position", "open_plate", "create_plate", "zstd_available", "crc32c"]

_DTYPES = {"uint8": np.uint8, "uint16": np.uint16, "int16": np.int16, "uint32": np.uint32, "int32": np.int32,
           "float32": np.float32, "float64": np.float64}


# ---- zstd through the runtime library ------------------------------------------------------------
class _Zstd:
    def __init__(self):
        self.lib = None
        for name in ("libzstd.so.1", ctypes.util.find_library("zstd")):
            if not name:
                continue
            try:
                lib = ctypes.CDLL(name)
            except OSError:
                continue
            lib.ZSTD_decompress.restype = ctypes.c_size_t
            lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
            lib.ZSTD_compress.restype = ctypes.c_size_t
            lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
            lib.ZSTD_compressBound.restype = ctypes.c_size_t
            lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
            lib.ZSTD_isError.restype = ctypes.c_uint
            lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
            self.lib = lib
            break

    def decompress_into(self, src, dst: np.ndarray) -> None:
        if self.lib is None:
            raise RuntimeError("zstd-compressed chunk but libzstd.so.1 is not available")
        src = np.frombuffer(src, dtype=np.uint8)             # bytes, memoryview or a slice of a memory-mapped shard
        n = self.lib.ZSTD_decompress(dst.ctypes.data, dst.nbytes, src.ctypes.data, src.nbytes)
        if self.lib.ZSTD_isError(n) or n != dst.nbytes:
            raise IOError(f"zstd: decoded {n} bytes, expected {dst.nbytes}")

    def compress(self, data: np.ndarray, level: int) -> bytes:
        if self.lib is None:
            raise RuntimeError("zstd codec requested but libzstd.so.1 is not available")
        bound = self.lib.ZSTD_compressBound(data.nbytes)
        buf = ctypes.create_string_buffer(bound)
        n = self.lib.ZSTD_compress(buf, bound, data.ctypes.data, data.nbytes, level)
        if self.lib.ZSTD_isError(n):
            raise IOError("zstd: compression failed")
        return buf.raw[:n]


_zstd = _Zstd()


def zstd_available() -> bool:
    return _zstd.lib is not None


_BLOSC_CODES = {"lz4": 1, "lz4hc": 1, "zlib": 3, "zstd": 4}
_BLOSC_SHUFFLES = {"noshuffle": 0, "shuffle": 1, "bitshuffle": 2}


@dataclass(frozen=True)
class Codec:
    """The bytes->bytes step after the little-endian ``bytes`` codec: none, ``zstd`` or ``blosc``."""

    kind: str = "raw"                  # "raw" | "zstd" | "blosc"
    level: int = 3
    cname: str = "zstd"                # blosc only
    shuffle: str = "shuffle"           # blosc only: noshuffle | shuffle | bitshuffle
    blocksize: int = 0                 # blosc only: 0 = the encoder's default
    split: bool = False                # blosc only, encoder: one stream per byte plane where blosc's rule allows

    def metadata(self, typesize: int) -> List[dict]:
        out = [{"name": "bytes", "configuration": {"endian": "little"}}]
        if self.kind == "zstd":
            out.append({"name": "zstd", "configuration": {"level": int(self.level), "checksum": False}})
        elif self.kind == "blosc":
            out.append({"name": "blosc", "configuration": {"cname": self.cname, "clevel": int(self.level),
                                                           "shuffle": self.shuffle, "typesize": int(typesize),
                                                           "blocksize": int(self.blocksize)}})
        return out

    # -- payload (uint8 array or bytes) -> out, in place ------------------------------------------------
    def decode_into(self, payload, out: np.ndarray, threads: int = 1) -> None:
        if self.kind == "raw":
            if len(payload) != out.nbytes:
                raise IOError(f"chunk holds {len(payload)} bytes, expected {out.nbytes}")
            out.reshape(-1).view(np.uint8)[:] = np.frombuffer(payload, dtype=np.uint8)
        elif self.kind == "zstd":
            _zstd.decompress_into(payload, out)
        else:
            from . import _cabi

            src = np.frombuffer(payload, dtype=np.uint8)
            _cabi.check(_cabi.lib().shrimpy_blosc_decode(src.ctypes.data, src.nbytes, out.ctypes.data, out.nbytes,
                                                         int(threads)))

    def encode(self, data: np.ndarray):
        if self.kind == "raw":
            return memoryview(data.reshape(-1).view(np.uint8))
        if self.kind == "zstd":
            return _zstd.compress(data, self.level)
        from . import _cabi

        lib = _cabi.lib()
        cap = lib.shrimpy_blosc_encode_bound(data.nbytes, self.blocksize, data.dtype.itemsize)
        buf = np.empty(cap, dtype=np.uint8)
        n = ctypes.c_size_t(0)
        _cabi.check(lib.shrimpy_blosc_encode(data.ctypes.data, data.nbytes, data.dtype.itemsize, _BLOSC_CODES[self.cname],
                                             int(self.level), _BLOSC_SHUFFLES[self.shuffle], int(self.blocksize),
                                             int(self.split), buf.ctypes.data, cap, ctypes.byref(n)))
        return memoryview(buf[:n.value])


def _codec_chain(codecs: Sequence[dict]) -> Codec:
    """Validate an array->bytes chain this module can run: bytes (little endian) [+ zstd | blosc]."""
    names = [c.get("name") for c in codecs]
    if not names or names[0] != "bytes":
        raise NotImplementedError(f"unsupported codec chain {names}: expected 'bytes' first")
    endian = (codecs[0].get("configuration") or {}).get("endian", "little")
    if endian != "little":
        raise NotImplementedError("big-endian chunks are not supported")
    rest = names[1:]
    if rest == []:
        return Codec()
    cfg = codecs[1].get("configuration") or {} if len(codecs) > 1 else {}
    if rest == ["zstd"]:
        return Codec("zstd", int(cfg.get("level", 3)))
    if rest == ["blosc"]:
        cname, shuffle = cfg.get("cname", "zstd"), cfg.get("shuffle", "noshuffle")
        if cname not in _BLOSC_CODES:
            raise NotImplementedError(f"blosc cname {cname!r} is not supported (zstd, lz4, lz4hc, zlib are)")
        if shuffle not in _BLOSC_SHUFFLES:
            raise NotImplementedError(f"blosc shuffle {shuffle!r} is not supported")
        return Codec("blosc", int(cfg.get("clevel", 5)), cname, shuffle, int(cfg.get("blocksize", 0) or 0))
    raise NotImplementedError(f"unsupported codec(s) {rest}: expected nothing, 'zstd' or 'blosc' after 'bytes'")


def _as_codec(zstd_level: Optional[int], blosc: Optional[dict]) -> Codec:
    if blosc is not None:
        if zstd_level is not None:
            raise ValueError("give zstd_level or blosc, not both")
        cfg = {"cname": "zstd", "clevel": 1, "shuffle": "shuffle", "blocksize": 0, "split": False, **blosc}
        if cfg["cname"] not in _BLOSC_CODES or cfg["shuffle"] not in _BLOSC_SHUFFLES:
            raise ValueError(f"unsupported blosc configuration {blosc}")
        return Codec("blosc", int(cfg["clevel"]), cfg["cname"], cfg["shuffle"], int(cfg["blocksize"]), bool(cfg["split"]))
    return Codec("zstd", int(zstd_level)) if zstd_level is not None else Codec()


def _crc32c_table():
    table = []
    for n in range(256):
        c = n
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        table.append(c)
    return table


_CRC32C = _crc32c_table()


def crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli) of a shard index (a few hundred bytes; table-driven, pure Python)."""
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC32C[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


@dataclass
class ZarrArray:
    """One Zarr v3 array on a local filesystem."""

    path: Path
    shape: Tuple[int, ...]
    chunks: Tuple[int, ...]
    dtype: np.dtype
    fill_value: float = 0
    codec: Codec = Codec()
    shard_inner: Optional[Tuple[int, ...]] = None      # inner chunk shape when the array is sharded
    shard_index_at_end: bool = True
    shard_index_crc: bool = True
    dimension_names: Tuple[str, ...] = ()
    attributes: dict = field(default_factory=dict)

    @property
    def zstd(self) -> bool:
        return self.codec.kind == "zstd"

    @property
    def zstd_level(self) -> int:
        return self.codec.level

    @property
    def compressed(self) -> bool:
        return self.codec.kind != "raw"

    # ---- construction -----------------------------------------------------------------------------
    @classmethod
    def open(cls, path) -> "ZarrArray":
        path = Path(path)
        meta = json.loads((path / "zarr.json").read_text())
        if meta.get("zarr_format") != 3 or meta.get("node_type") != "array":
            raise ValueError(f"{path} is not a Zarr v3 array")
        if meta["chunk_grid"]["name"] != "regular":
            raise NotImplementedError("only regular chunk grids are supported")
        enc = meta.get("chunk_key_encoding", {"name": "default"})
        if enc.get("name") != "default" or (enc.get("configuration") or {}).get("separator", "/") != "/":
            raise NotImplementedError("only the default chunk key encoding with '/' is supported")
        dtype = meta["data_type"]
        if dtype not in _DTYPES:
            raise NotImplementedError(f"data_type {dtype!r} is not supported")
        codecs = meta["codecs"]
        chunks = tuple(meta["chunk_grid"]["configuration"]["chunk_shape"])
        kw = {}
        if codecs and codecs[0].get("name") == "sharding_indexed":
            if len(codecs) != 1:
                raise NotImplementedError("codecs after sharding_indexed are not supported")
            cfg = codecs[0]["configuration"]
            kw["shard_inner"] = tuple(int(v) for v in cfg["chunk_shape"])
            if len(kw["shard_inner"]) != len(chunks) or any(c % i for c, i in zip(chunks, kw["shard_inner"])):
                raise ValueError(f"inner chunks {kw['shard_inner']} do not tile shards {chunks}")
            kw["shard_index_at_end"] = cfg.get("index_location", "end") == "end"
            index_codecs = [c["name"] for c in cfg.get("index_codecs", [{"name": "bytes"}, {"name": "crc32c"}])]
            if index_codecs not in (["bytes"], ["bytes", "crc32c"]):
                raise NotImplementedError(f"shard index codecs {index_codecs} are not supported")
            kw["shard_index_crc"] = "crc32c" in index_codecs
            codec = _codec_chain(cfg["codecs"])
        else:
            codec = _codec_chain(codecs)
        return cls(path=path, shape=tuple(meta["shape"]), chunks=chunks, dtype=np.dtype(_DTYPES[dtype]),
                   fill_value=meta.get("fill_value", 0) or 0, codec=codec,
                   dimension_names=tuple(meta.get("dimension_names") or ()), attributes=meta.get("attributes", {}), **kw)

    @classmethod
    def create(cls, path, shape, chunks, dtype, *, zstd_level: Optional[int] = None, blosc: Optional[dict] = None,
               shard_inner: Optional[Sequence[int]] = None, dimension_names=(), attributes: Optional[dict] = None,
               fill_value=0) -> "ZarrArray":
        """``blosc={"cname": "zstd", "clevel": 1, "shuffle": "shuffle"}`` with ``shard_inner=(1, 1, zi, Y, X)`` gives the
        layout of the reference acquisition (``chunks`` is then the shard shape, one file per shard)."""
        path = Path(path)
        path.mkdir(parents=True, exist_ok=True)
        dtype = np.dtype(dtype)
        if dtype.name not in _DTYPES:
            raise NotImplementedError(f"dtype {dtype} is not supported")
        codec = _as_codec(zstd_level, blosc)
        codecs = codec.metadata(dtype.itemsize)
        if shard_inner is not None:
            shard_inner = tuple(map(int, shard_inner))
            if len(shard_inner) != len(chunks) or any(int(c) % i for c, i in zip(chunks, shard_inner)):
                raise ValueError(f"inner chunks {shard_inner} do not tile shards {tuple(chunks)}")
            codecs = [{"name": "sharding_indexed", "configuration": {
                "chunk_shape": list(shard_inner), "codecs": codecs,
                "index_codecs": [{"name": "bytes", "configuration": {"endian": "little"}}, {"name": "crc32c"}],
                "index_location": "end"}}]
        meta = {"zarr_format": 3, "node_type": "array", "shape": list(map(int, shape)), "data_type": dtype.name,
                "chunk_grid": {"name": "regular", "configuration": {"chunk_shape": list(map(int, chunks))}},
                "chunk_key_encoding": {"name": "default", "configuration": {"separator": "/"}},
                "fill_value": fill_value, "codecs": codecs, "attributes": attributes or {}}
        if dimension_names:
            meta["dimension_names"] = list(dimension_names)
        (path / "zarr.json").write_text(json.dumps(meta, indent=1))
        return cls(path=path, shape=tuple(map(int, shape)), chunks=tuple(map(int, chunks)), dtype=dtype,
                   fill_value=fill_value, codec=codec, shard_inner=shard_inner,
                   dimension_names=tuple(dimension_names), attributes=attributes or {})

    # ---- chunk access -----------------------------------------------------------------------------
    @property
    def grid(self) -> Tuple[int, ...]:
        return tuple(-(-s // c) for s, c in zip(self.shape, self.chunks))

    def chunk_path(self, index: Sequence[int]) -> Path:
        return self.path / "c" / "/".join(str(int(i)) for i in index)

    def read_chunk_into(self, index: Sequence[int], out: np.ndarray, threads: int = 1) -> int:
        """Decode one (outer) chunk into ``out`` (C-contiguous, full chunk shape). Returns bytes read from disk."""
        if tuple(out.shape) != self.chunks or out.dtype != self.dtype or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous {self.dtype} array of shape {self.chunks}")
        path = self.chunk_path(index)
        if not path.exists():
            out[...] = self.fill_value
            return 0
        if self.shard_inner is not None:
            jobs, nbytes = self._shard_jobs(path, out, threads)
            for fn, args in jobs:
                fn(*args)
            return nbytes
        if not self.compressed:
            with open(path, "rb", buffering=0) as fh:       # straight into the (pinned) destination
                n = fh.readinto(memoryview(out.reshape(-1).view(np.uint8)))
            if n != out.nbytes:
                raise IOError(f"{path}: read {n} bytes, expected {out.nbytes}")
            return n
        payload = _map_file(path)
        self.codec.decode_into(payload, out, threads)
        return len(payload)

    def _shard_jobs(self, path: Path, dest: np.ndarray, threads: int = 1):
        """Independent decode jobs ``[(fn, args)]`` for the inner chunks of one shard file, landing in ``dest`` -- the
        shard's region of the caller's buffer, possibly cut short along any axis at the array edge.  The file is
        memory-mapped: payloads are decoded from the page cache without an intermediate copy, and an inner chunk
        whose destination is contiguous (a z-slab of a stack) is decoded in place.  Returns (jobs, shard bytes)."""
        inner = self.shard_inner
        per_axis = tuple(c // i for c, i in zip(self.chunks, inner))
        n_inner = int(np.prod(per_axis))
        blob = _map_file(path)
        index_bytes = n_inner * 16 + (4 if self.shard_index_crc else 0)
        if len(blob) < index_bytes:
            raise IOError(f"{path}: {len(blob)} bytes cannot hold a shard index of {index_bytes}")
        raw_index = blob[len(blob) - index_bytes:] if self.shard_index_at_end else blob[:index_bytes]
        index_payload = raw_index[:n_inner * 16].tobytes()
        if self.shard_index_crc:
            stored = int.from_bytes(raw_index[n_inner * 16:n_inner * 16 + 4].tobytes(), "little")
            if crc32c(index_payload) != stored:
                raise IOError(f"{path}: shard index checksum mismatch (crc32c)")
        table = np.frombuffer(index_payload, dtype="<u8").reshape(n_inner, 2)
        jobs = []
        for flat, (offset, nbytes) in enumerate(table):
            sub = np.unravel_index(flat, per_axis)
            lo = tuple(int(s) * i for s, i in zip(sub, inner))
            hi = tuple(min(l + i, d) for l, i, d in zip(lo, inner, dest.shape))
            if any(h <= l for l, h in zip(lo, hi)):
                continue                                    # wholly beyond the array edge
            target = dest[tuple(slice(l, h) for l, h in zip(lo, hi))]
            if offset == 2**64 - 1 and nbytes == 2**64 - 1:
                jobs.append((_fill, (target, self.fill_value)))
                continue
            if int(offset) + int(nbytes) > len(blob):
                raise IOError(f"{path}: inner chunk {flat} lies outside the shard")
            payload = blob[int(offset):int(offset) + int(nbytes)]
            whole = tuple(target.shape) == inner and target.flags.c_contiguous
            jobs.append((self._decode_inner, (payload, target, whole, threads)))
        return jobs, len(blob)

    def _decode_inner(self, payload, target: np.ndarray, whole: bool, threads: int) -> None:
        if whole:
            self.codec.decode_into(payload, target, threads)
        else:
            tmp = np.empty(self.shard_inner, dtype=self.dtype)
            self.codec.decode_into(payload, tmp, threads)
            target[...] = tmp[tuple(slice(0, n) for n in target.shape)]

    def _full_chunk(self, data: np.ndarray, shape: Tuple[int, ...]) -> np.ndarray:
        data = np.ascontiguousarray(data, dtype=self.dtype)
        if tuple(data.shape) != tuple(shape):
            full = np.full(shape, self.fill_value, dtype=self.dtype)             # edge chunks are stored full-size
            full[tuple(slice(0, s) for s in data.shape)] = data
            data = full
        return data

    def write_chunk(self, index: Sequence[int], data: np.ndarray, pool=None) -> int:
        """Write one (outer) chunk.  For a sharded array that is one shard file -- inner chunks in C order, the
        ``(offset, nbytes)`` index and its CRC-32C at the end; with ``pool`` the inner chunks are encoded concurrently
        (call it from outside that pool)."""
        data = self._full_chunk(data, self.chunks)
        path = self.chunk_path(index)
        path.parent.mkdir(parents=True, exist_ok=True)
        if self.shard_inner is None:
            payload = self.codec.encode(data)
            with open(path, "wb") as fh:
                fh.write(payload)
            return len(payload)
        inner = self.shard_inner
        per_axis = tuple(c // i for c, i in zip(self.chunks, inner))
        blocks = []
        for flat in range(int(np.prod(per_axis))):
            sub = np.unravel_index(flat, per_axis)
            blocks.append(data[tuple(slice(int(s) * i, (int(s) + 1) * i) for s, i in zip(sub, inner))])
        encode = lambda block: self.codec.encode(np.ascontiguousarray(block))       # noqa: E731
        payloads = list(pool.map(encode, blocks)) if pool is not None else [encode(b) for b in blocks]
        table = np.empty((len(blocks), 2), dtype="<u8")
        pos = 0
        with open(path, "wb") as fh:
            for flat, payload in enumerate(payloads):
                fh.write(payload)
                table[flat] = (pos, len(payload))
                pos += len(payload)
            raw_index = table.tobytes()
            fh.write(raw_index + crc32c(raw_index).to_bytes(4, "little"))
        return pos + len(raw_index) + 4

    # ---- (t, c) stacks of a TCZYX array -------------------------------------------------------------
    def _check_tczyx(self) -> None:
        if len(self.shape) != 5:
            raise ValueError(f"expected a 5-D TCZYX array, got shape {self.shape}")

    def stack_chunks(self, t: int, c: int) -> List[Tuple[Tuple[int, ...], slice]]:
        """Chunks making up stack ``(t, c)`` when chunks span full Y and X: [(chunk index, z slice)]."""
        self._check_tczyx()
        if self.chunks[0] != 1 or self.chunks[1] != 1 or self.chunks[3:] != self.shape[3:]:
            raise NotImplementedError(f"stack streaming needs chunks (1, 1, Zc, Y, X), got {self.chunks}")
        Z, zc = self.shape[2], self.chunks[2]
        return [((t, c, k, 0, 0), slice(k * zc, min((k + 1) * zc, Z))) for k in range(-(-Z // zc))]

    def read_stack_into(self, t: int, c: int, out: np.ndarray, pool=None, piece_bytes: int = 8 << 20) -> int:
        """Read the ``(Z, Y, X)`` stack of ``(t, c)`` into ``out``; z-chunks land in place. Returns disk bytes.

        With ``pool`` (a ``ThreadPoolExecutor``) the work is cut into independent pieces -- byte ranges of
        ``piece_bytes`` for uncompressed chunks (``os.preadv`` straight into the destination, the GIL is
        released), whole (inner) chunks for compressed ones, decoded in place (the native decoders release the
        GIL too) -- and run concurrently.  Only the part of a partial last chunk that holds data is read.
        Chunks (or shards) that also tile Y and X are decoded through a scratch block and copied into their window
        of ``out`` -- any regular chunking of a TCZYX array can be read, the z-slab layouts are the fast ones.
        """
        self._check_tczyx()
        Z, Y, X = self.shape[2:]
        if tuple(out.shape) != (Z, Y, X) or out.dtype != self.dtype or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous {self.dtype} array of shape {(Z, Y, X)}")
        if self.chunks[0] != 1 or self.chunks[1] != 1:
            raise NotImplementedError(f"stack streaming needs chunks of one timepoint and channel, got {self.chunks}")
        zc, yc, xc = self.chunks[2:]
        slabs = (yc, xc) == (Y, X)                          # every chunk is a contiguous z-slab of the stack
        cells = [(kz, ky, kx) for kz in range(-(-Z // zc)) for ky in range(-(-Y // yc)) for kx in range(-(-X // xc))]
        tasks, listed = [], 0
        workers = getattr(pool, "_max_workers", 1) if pool is not None else 1
        for kz, ky, kx in cells:
            index = (t, c, kz, ky, kx)
            window = out[kz * zc:(kz + 1) * zc, ky * yc:(ky + 1) * yc, kx * xc:(kx + 1) * xc]   # clipped at the edges
            nz = window.shape[0]
            path = self.chunk_path(index)
            if not path.exists():
                window[...] = self.fill_value
            elif self.shard_inner is not None:
                inner_count = int(np.prod([n // i for n, i in zip(self.chunks, self.shard_inner)]))
                threads = max(1, workers // max(1, inner_count * len(cells)))
                jobs, nbytes = self._shard_jobs(path, window[None, None], threads)
                tasks += jobs
                listed += nbytes
            elif not slabs:
                tasks.append((self._read_window, (index, window)))
            elif not self.compressed:
                view = memoryview(window.reshape(-1).view(np.uint8))          # the data-bearing prefix of the chunk
                for off in range(0, len(view), piece_bytes):
                    tasks.append((_pread_piece, (path, off, view[off:off + piece_bytes])))
            elif nz == zc:
                tasks.append((self.read_chunk_into, (index, window.reshape(self.chunks), max(1, workers // len(cells)))))
            else:                                           # compressed partial chunk: decode, then copy the prefix
                tasks.append((self._read_window, (index, window)))
        if pool is None or len(tasks) <= 1:
            return listed + sum(fn(*a) or 0 for fn, a in tasks)
        return listed + sum(f.result() or 0 for f in [pool.submit(fn, *a) for fn, a in tasks])

    def _read_window(self, index, window: np.ndarray) -> int:
        """Decode a whole chunk into scratch and copy the part that lies inside the array into ``window``."""
        scratch = np.empty(self.chunks, dtype=self.dtype)
        n = self.read_chunk_into(index, scratch)
        window[...] = scratch[0, 0][tuple(slice(0, k) for k in window.shape)]
        return n

    def write_stack(self, t: int, c: int, data: np.ndarray, pool=None) -> int:
        """Write stack ``(t, c)``; with ``pool`` the chunk files are written concurrently (one writer per file:
        concurrent writers of ONE file serialise on its inode lock, so parallelism comes from the chunk count)."""
        self._check_tczyx()
        jobs = [(index, data[zs][None, None]) for index, zs in self.stack_chunks(t, c)]
        if self.shard_inner is not None:                    # shard files one after the other, inner chunks in parallel
            return sum(self.write_chunk(index, block, pool) for index, block in jobs)
        if pool is None or len(jobs) <= 1:
            return sum(self.write_chunk(index, block) for index, block in jobs)
        return sum(f.result() for f in [pool.submit(self.write_chunk, index, block) for index, block in jobs])


def _fill(target: np.ndarray, value) -> None:
    target[...] = value


def _map_file(path) -> np.ndarray:
    """The file as a read-only uint8 array over a private memory map (the map lives as long as any slice of it)."""
    with open(path, "rb") as fh:
        size = os.fstat(fh.fileno()).st_size
        if size == 0:
            return np.empty(0, dtype=np.uint8)
        return np.frombuffer(mmap.mmap(fh.fileno(), size, access=mmap.ACCESS_READ), dtype=np.uint8)


def _pread_piece(path, offset: int, dest: memoryview) -> int:
    # (Copying out of a memory map instead -- 6 GB/s per thread against 3-4 GB/s here for one thread -- was measured on
    # a B200 box with 14 loader threads: 8.0 -> 4.6 GVoxel/s for the plate; the threads serialise on the address
    # space's lock while mapping.  preadv stays.)
    fd = os.open(path, os.O_RDONLY)
    try:
        done = 0
        while done < len(dest):
            n = os.preadv(fd, [dest[done:]], offset + done)
            if n <= 0:
                raise IOError(f"{path}: short read at {offset + done}")
            done += n
        return done
    finally:
        os.close(fd)


# ---- OME-NGFF 0.5 HCS plate ----------------------------------------------------------------------
@dataclass
class Position:
    name: str                      # "row/col/fov"
    array: ZarrArray               # resolution level 0, TCZYX
    channel_names: Tuple[str, ...]
    scale: Tuple[float, ...]       # TCZYX scale of level 0


def _write_group(path: Path, ome: Optional[dict] = None) -> None:
    path.mkdir(parents=True, exist_ok=True)
    attrs = {"ome": ome} if ome is not None else {}
    (path / "zarr.json").write_text(json.dumps({"zarr_format": 3, "node_type": "group", "attributes": attrs}, indent=1))


def _read_ome(path: Path) -> dict:
    meta = json.loads((path / "zarr.json").read_text())
    attrs = meta.get("attributes", {})
    return attrs.get("ome", attrs)


def create_plate(path, position_names: Sequence[str], shape_tczyx, chunks, dtype=np.uint16, *,
                 channel_names: Optional[Sequence[str]] = None, scale=(1.0, 1.0, 1.0, 1.0, 1.0),
                 zstd_level: Optional[int] = None, blosc: Optional[dict] = None,
                 shard_inner: Optional[Sequence[int]] = None) -> List[Position]:
    """Create an HCS plate ``row/col/fov/0`` (NGFF 0.5 metadata) with empty arrays; returns its positions."""
    path = Path(path)
    names = [tuple(n.split("/")) for n in position_names]
    if any(len(n) != 3 for n in names):
        raise ValueError("position names must look like 'A/1/fov0'")
    rows = sorted({n[0] for n in names})
    cols = sorted({n[1] for n in names}, key=lambda s: (len(s), s))
    wells = sorted({(n[0], n[1]) for n in names})
    _write_group(path, {"version": "0.5", "plate": {
        "rows": [{"name": r} for r in rows], "columns": [{"name": c} for c in cols],
        "wells": [{"path": f"{r}/{c}", "rowIndex": rows.index(r), "columnIndex": cols.index(c)} for r, c in wells],
        "version": "0.5"}})
    channel_names = list(channel_names or [f"ch{i}" for i in range(shape_tczyx[1])])
    axes = [{"name": "t", "type": "time"}, {"name": "c", "type": "channel"},
            {"name": "z", "type": "space", "unit": "micrometer"}, {"name": "y", "type": "space", "unit": "micrometer"},
            {"name": "x", "type": "space", "unit": "micrometer"}]
    out = []
    for r in rows:
        _write_group(path / r)
    for r, c in wells:
        fovs = [n[2] for n in names if (n[0], n[1]) == (r, c)]
        _write_group(path / r / c, {"version": "0.5", "well": {"images": [{"path": f} for f in fovs], "version": "0.5"}})
    for r, c, f in names:
        _write_group(path / r / c / f, {"version": "0.5", "multiscales": [{
            "version": "0.5", "name": "0", "axes": axes,
            "datasets": [{"path": "0", "coordinateTransformations": [{"type": "scale", "scale": list(map(float, scale))}]}]}],
            "omero": {"channels": [{"label": ch, "active": True} for ch in channel_names]}})
        arr = ZarrArray.create(path / r / c / f / "0", shape_tczyx, chunks, dtype, zstd_level=zstd_level, blosc=blosc,
                               shard_inner=shard_inner, dimension_names=("t", "c", "z", "y", "x"))
        out.append(Position(f"{r}/{c}/{f}", arr, tuple(channel_names), tuple(map(float, scale))))
    return out


def open_plate(path) -> List[Position]:
    """Enumerate the positions of an HCS plate (or of a single FOV group) in plate order."""
    path = Path(path)
    ome = _read_ome(path)

    def image(p: Path, name: str) -> Position:
        meta = _read_ome(p)
        ms = meta["multiscales"][0]
        ds = ms["datasets"][0]
        scale = next((tuple(t["scale"]) for t in ds.get("coordinateTransformations", []) if t.get("type") == "scale"),
                     (1.0,) * 5)
        channels = tuple(ch.get("label", str(i)) for i, ch in enumerate(meta.get("omero", {}).get("channels", [])))
        return Position(name, ZarrArray.open(p / ds["path"]), channels, scale)

    if "plate" in ome:
        out = []
        for well in ome["plate"]["wells"]:
            wmeta = _read_ome(path / well["path"])
            for img in wmeta["well"]["images"]:
                out.append(image(path / well["path"] / img["path"], f"{well['path']}/{img['path']}"))
        return out
    if "multiscales" in ome:
        return [image(path, path.name)]
    raise ValueError(f"{path}: neither an HCS plate nor an image group")
