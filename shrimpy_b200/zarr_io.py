"""Minimal Zarr v3 / OME-NGFF 0.5 reader-writer for the streaming deskew (no ``zarr``/``iohub`` needed).

shrimPy writes and replays OME-Zarr: HCS plates ``row/col/fov/0`` holding ``TCZYX`` arrays, z-chunked
(``shrimpy/mantis/mantis_engine.py:474-481`` ``chunk_size=min(512, nz)``; ``scripts/measure_psf.py:273-287``
writes ``chunks=(1, 1, 50, Y, X)``; ``shrimpy/replay_camera.py:176-204, 293-308`` reads one ``(t, c)``
volume at a time).  This module implements exactly what the loader needs from that format:

* arrays: regular chunk grid, ``default`` chunk-key encoding, codecs ``bytes`` (little endian) optionally
  followed by ``zstd`` (through ``libzstd.so.1`` via ctypes when the library is present), and READING of
  ``sharding_indexed`` shards whose inner codecs are ``bytes``[+``zstd``];
* groups: plate / well / image metadata of NGFF 0.5 (``attributes.ome``), enough to enumerate positions,
  channel names and the scale transform, and to write a deskewed plate back.

Deviation, stated: the reference acquisition compresses with **blosc-zstd** inside shards
(``shrimpy/tests/test_mantis_integration.py:152-188``).  Blosc framing is not implemented (no blosc library
offline); such stores raise ``NotImplementedError`` naming the codec.  Synthetic benchmark plates are written
uncompressed or zstd-compressed.

A chunk of a ``(1, 1, Zc, Y, X)`` grid is one contiguous z-slab of the ``(Z, Y, X)`` stack, so
``read_stack_into`` lands every chunk file directly at its offset of a caller-provided (pinned) buffer --
no intermediate copy for uncompressed data, one decompress-into-place for zstd.
"""

from __future__ import annotations

import ctypes
import ctypes.util
import json
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

__all__ = ["ZarrArray", "Position", "open_plate", "create_plate", "zstd_available"]

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

    def decompress_into(self, src: bytes, dst: np.ndarray) -> None:
        if self.lib is None:
            raise RuntimeError("zstd-compressed chunk but libzstd.so.1 is not available")
        n = self.lib.ZSTD_decompress(dst.ctypes.data, dst.nbytes, src, len(src))
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


def _codec_chain(codecs: Sequence[dict]) -> Tuple[bool, Optional[int]]:
    """Validate an array->bytes chain this module can run: bytes (little endian) [+ zstd]. Returns (zstd?, level)."""
    names = [c.get("name") for c in codecs]
    if not names or names[0] != "bytes":
        raise NotImplementedError(f"unsupported codec chain {names}: expected 'bytes' first")
    endian = (codecs[0].get("configuration") or {}).get("endian", "little")
    if endian != "little":
        raise NotImplementedError("big-endian chunks are not supported")
    rest = names[1:]
    if rest == []:
        return False, None
    if rest == ["zstd"]:
        return True, int((codecs[1].get("configuration") or {}).get("level", 3))
    raise NotImplementedError(
        f"unsupported codec(s) {rest} (blosc framing is not implemented offline; use bytes or bytes+zstd)")


@dataclass
class ZarrArray:
    """One Zarr v3 array on a local filesystem."""

    path: Path
    shape: Tuple[int, ...]
    chunks: Tuple[int, ...]
    dtype: np.dtype
    fill_value: float = 0
    zstd: bool = False
    zstd_level: int = 3
    shard_inner: Optional[Tuple[int, ...]] = None      # inner chunk shape when the array is sharded
    shard_index_at_end: bool = True
    shard_index_crc: bool = True
    dimension_names: Tuple[str, ...] = ()
    attributes: dict = field(default_factory=dict)

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
        kw = {}
        if codecs and codecs[0].get("name") == "sharding_indexed":
            if len(codecs) != 1:
                raise NotImplementedError("codecs after sharding_indexed are not supported")
            cfg = codecs[0]["configuration"]
            kw["shard_inner"] = tuple(int(v) for v in cfg["chunk_shape"])
            kw["shard_index_at_end"] = cfg.get("index_location", "end") == "end"
            index_codecs = [c["name"] for c in cfg.get("index_codecs", [{"name": "bytes"}, {"name": "crc32c"}])]
            if index_codecs not in (["bytes"], ["bytes", "crc32c"]):
                raise NotImplementedError(f"shard index codecs {index_codecs} are not supported")
            kw["shard_index_crc"] = "crc32c" in index_codecs
            zstd, level = _codec_chain(cfg["codecs"])
        else:
            zstd, level = _codec_chain(codecs)
        return cls(path=path, shape=tuple(meta["shape"]),
                   chunks=tuple(meta["chunk_grid"]["configuration"]["chunk_shape"]), dtype=np.dtype(_DTYPES[dtype]),
                   fill_value=meta.get("fill_value", 0) or 0, zstd=zstd, zstd_level=level or 3,
                   dimension_names=tuple(meta.get("dimension_names") or ()), attributes=meta.get("attributes", {}), **kw)

    @classmethod
    def create(cls, path, shape, chunks, dtype, *, zstd_level: Optional[int] = None, dimension_names=(),
               attributes: Optional[dict] = None, fill_value=0) -> "ZarrArray":
        path = Path(path)
        path.mkdir(parents=True, exist_ok=True)
        dtype = np.dtype(dtype)
        if dtype.name not in _DTYPES:
            raise NotImplementedError(f"dtype {dtype} is not supported")
        codecs = [{"name": "bytes", "configuration": {"endian": "little"}}]
        if zstd_level is not None:
            codecs.append({"name": "zstd", "configuration": {"level": int(zstd_level), "checksum": False}})
        meta = {"zarr_format": 3, "node_type": "array", "shape": list(map(int, shape)), "data_type": dtype.name,
                "chunk_grid": {"name": "regular", "configuration": {"chunk_shape": list(map(int, chunks))}},
                "chunk_key_encoding": {"name": "default", "configuration": {"separator": "/"}},
                "fill_value": fill_value, "codecs": codecs, "attributes": attributes or {}}
        if dimension_names:
            meta["dimension_names"] = list(dimension_names)
        (path / "zarr.json").write_text(json.dumps(meta, indent=1))
        return cls(path=path, shape=tuple(map(int, shape)), chunks=tuple(map(int, chunks)), dtype=dtype,
                   fill_value=fill_value, zstd=zstd_level is not None, zstd_level=zstd_level or 3,
                   dimension_names=tuple(dimension_names), attributes=attributes or {})

    # ---- chunk access -----------------------------------------------------------------------------
    @property
    def grid(self) -> Tuple[int, ...]:
        return tuple(-(-s // c) for s, c in zip(self.shape, self.chunks))

    def chunk_path(self, index: Sequence[int]) -> Path:
        return self.path / "c" / "/".join(str(int(i)) for i in index)

    def _decode_into(self, payload: bytes, out: np.ndarray) -> None:
        if self.zstd:
            _zstd.decompress_into(payload, out)
        else:
            if len(payload) != out.nbytes:
                raise IOError(f"chunk holds {len(payload)} bytes, expected {out.nbytes}")
            out.reshape(-1).view(np.uint8)[:] = np.frombuffer(payload, dtype=np.uint8)

    def read_chunk_into(self, index: Sequence[int], out: np.ndarray) -> int:
        """Decode one (outer) chunk into ``out`` (C-contiguous, full chunk shape). Returns bytes read from disk."""
        if tuple(out.shape) != self.chunks or out.dtype != self.dtype or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous {self.dtype} array of shape {self.chunks}")
        path = self.chunk_path(index)
        if not path.exists():
            out[...] = self.fill_value
            return 0
        if self.shard_inner is not None:
            return self._read_shard_into(path, out)
        if not self.zstd:
            with open(path, "rb", buffering=0) as fh:       # straight into the (pinned) destination
                n = fh.readinto(memoryview(out.reshape(-1).view(np.uint8)))
            if n != out.nbytes:
                raise IOError(f"{path}: read {n} bytes, expected {out.nbytes}")
            return n
        payload = path.read_bytes()
        self._decode_into(payload, out)
        return len(payload)

    def _read_shard_into(self, path: Path, out: np.ndarray) -> int:
        inner = self.shard_inner
        per_axis = tuple(c // i for c, i in zip(self.chunks, inner))
        n_inner = int(np.prod(per_axis))
        blob = path.read_bytes()
        index_bytes = n_inner * 16 + (4 if self.shard_index_crc else 0)
        raw_index = blob[-index_bytes:] if self.shard_index_at_end else blob[:index_bytes]
        table = np.frombuffer(raw_index[:n_inner * 16], dtype="<u8").reshape(n_inner, 2)
        tmp = np.empty(inner, dtype=self.dtype)
        for flat, (offset, nbytes) in enumerate(table):
            sub = np.unravel_index(flat, per_axis)
            sl = tuple(slice(s * i, (s + 1) * i) for s, i in zip(sub, inner))
            if offset == 2**64 - 1 and nbytes == 2**64 - 1:
                out[sl] = self.fill_value
                continue
            self._decode_into(blob[int(offset):int(offset + nbytes)], tmp)
            out[sl] = tmp
        return len(blob)

    def write_chunk(self, index: Sequence[int], data: np.ndarray) -> int:
        if self.shard_inner is not None:
            raise NotImplementedError("writing sharded arrays is not supported")
        data = np.ascontiguousarray(data, dtype=self.dtype)
        if tuple(data.shape) != self.chunks:
            full = np.full(self.chunks, self.fill_value, dtype=self.dtype)       # edge chunks are stored full-size
            full[tuple(slice(0, s) for s in data.shape)] = data
            data = full
        path = self.chunk_path(index)
        path.parent.mkdir(parents=True, exist_ok=True)
        payload = _zstd.compress(data, self.zstd_level) if self.zstd else memoryview(data.reshape(-1).view(np.uint8))
        with open(path, "wb") as fh:
            fh.write(payload)
        return len(payload)

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

    def read_stack_into(self, t: int, c: int, out: np.ndarray, pool=None, piece_bytes: int = 32 << 20) -> int:
        """Read the ``(Z, Y, X)`` stack of ``(t, c)`` into ``out``; z-chunks land in place. Returns disk bytes.

        With ``pool`` (a ``ThreadPoolExecutor``) the work is cut into independent pieces -- byte ranges of
        ``piece_bytes`` for uncompressed chunks (``os.preadv`` straight into the destination, the GIL is
        released), whole chunks for compressed ones -- and read concurrently.  Only the part of a partial
        last chunk that holds data is read.
        """
        self._check_tczyx()
        Z, Y, X = self.shape[2:]
        if tuple(out.shape) != (Z, Y, X) or out.dtype != self.dtype or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous {self.dtype} array of shape {(Z, Y, X)}")
        zc = self.chunks[2]
        tasks = []
        for index, zs in self.stack_chunks(t, c):
            nz = zs.stop - zs.start
            path = self.chunk_path(index)
            if not self.zstd and self.shard_inner is None:
                if not path.exists():
                    out[zs] = self.fill_value
                    continue
                view = memoryview(out[zs].reshape(-1).view(np.uint8))       # the data-bearing prefix of the chunk
                for off in range(0, len(view), piece_bytes):
                    tasks.append((_pread_piece, (path, off, view[off:off + piece_bytes])))
            elif nz == zc:
                tasks.append((self.read_chunk_into, (index, out[zs].reshape(self.chunks))))
            else:                                           # compressed partial chunk: decode, then copy the prefix
                tasks.append((self._read_partial, (index, out[zs])))
        if pool is None or len(tasks) <= 1:
            return sum(fn(*a) for fn, a in tasks)
        return sum(f.result() for f in [pool.submit(fn, *a) for fn, a in tasks])

    def _read_partial(self, index, dest: np.ndarray) -> int:
        scratch = np.empty(self.chunks, dtype=self.dtype)
        n = self.read_chunk_into(index, scratch)
        dest[...] = scratch[0, 0, :dest.shape[0]]
        return n

    def write_stack(self, t: int, c: int, data: np.ndarray, pool=None) -> int:
        """Write stack ``(t, c)``; with ``pool`` the chunk files are written concurrently (one writer per file:
        concurrent writers of ONE file serialise on its inode lock, so parallelism comes from the chunk count)."""
        self._check_tczyx()
        jobs = [(index, data[zs][None, None]) for index, zs in self.stack_chunks(t, c)]
        if pool is None or len(jobs) <= 1:
            return sum(self.write_chunk(index, block) for index, block in jobs)
        return sum(f.result() for f in [pool.submit(self.write_chunk, index, block) for index, block in jobs])


def _pread_piece(path, offset: int, dest: memoryview) -> int:
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
                 zstd_level: Optional[int] = None) -> List[Position]:
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
        arr = ZarrArray.create(path / r / c / f / "0", shape_tczyx, chunks, dtype, zstd_level=zstd_level,
                               dimension_names=("t", "c", "z", "y", "x"))
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
