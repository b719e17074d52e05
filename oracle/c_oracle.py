"""ctypes front-end of the plain-C oracle (``oracle/deskew_oracle.c``).

TEST INFRASTRUCTURE ONLY; parity unpinned (see ``deskew_oracle.py``).  The C
entry points take a ``[begin, end)`` slab over the outermost output axis, and
this module fans slabs out over Python threads (ctypes drops the GIL), which
is how the multi-core CPU baseline of ``bench.py`` is produced.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

from . import deskew_oracle as _py

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile the C oracle with the committed Makefile (gcc only)."""
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < (_HERE / "deskew_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(str(_LIB_PATH))
        lib.oracle_deskew.restype = ctypes.c_int
        lib.oracle_deskew.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_float,
            ctypes.c_int, ctypes.c_int,
        ]
        lib.oracle_affine.restype = ctypes.c_int
        lib.oracle_affine.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_int,
        ]
        _lib = lib
    return _lib


def _fan_out(fn, total: int, threads: int) -> None:
    threads = max(1, min(int(threads), total)) if total else 1
    if threads == 1:
        rc = fn(0, total)
        if rc:
            raise RuntimeError(f"C oracle returned {rc}")
        return
    edges = np.linspace(0, total, threads + 1).astype(int)
    with ThreadPoolExecutor(threads) as pool:
        for rc in pool.map(lambda i: fn(int(edges[i]), int(edges[i + 1])), range(threads)):
            if rc:
                raise RuntimeError(f"C oracle returned {rc}")


def deskew_data(raw_data: np.ndarray, ls_angle_deg: float, px_to_scan_ratio: float, keep_overhang: bool,
                average_n_slices: int = 1, cval: float | None = 0.0, threads: int | None = None) -> np.ndarray:
    """C restatement of ``deskew_oracle.deskew_data`` (uint16 or float32 input, float32 output)."""
    lib = _load()
    raw = np.ascontiguousarray(raw_data)
    if raw.dtype != np.uint16:
        raw = raw.astype(np.float32, copy=False)
    if cval is None:
        cval = float(raw.min())
    Z, Y, X = raw.shape
    M = _py.deskew_affine_matrix(raw.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang)
    (Yn, _, Xp), _ = _py.get_deskewed_data_shape(raw.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang,
                                                 average_n_slices)
    out = np.empty((Yn, X, Xp), dtype=np.float32)
    if out.size == 0:
        return out

    def run(b, e):
        return lib.oracle_deskew(raw.ctypes.data, int(raw.dtype == np.float32), out.ctypes.data, Z, Y, X, Xp,
                                 int(average_n_slices), float(M[0, 0]), float(M[0, 2]), float(M[0, 3]),
                                 float(cval), b, e)

    _fan_out(run, Yn, threads or (os.cpu_count() or 1))
    return out


def apply_affine_transform(zyx_data: np.ndarray, matrix: np.ndarray, output_shape_zyx, cval: float = 0.0,
                           threads: int | None = None) -> np.ndarray:
    """C restatement of ``deskew_oracle.apply_affine_transform``."""
    lib = _load()
    vol = np.ascontiguousarray(zyx_data, dtype=np.float32)
    M = np.ascontiguousarray(np.asarray(matrix, dtype=np.float64)[:3, :4])
    oz, oy, ox = (int(s) for s in output_shape_zyx)
    out = np.empty((oz, oy, ox), dtype=np.float32)
    if out.size == 0:
        return out

    def run(b, e):
        return lib.oracle_affine(vol.ctypes.data, out.ctypes.data, *vol.shape, oz, oy, ox, M.ctypes.data,
                                 float(cval), b, e)

    _fan_out(run, oz, threads or (os.cpu_count() or 1))
    return out
