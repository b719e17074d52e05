"""Light-sheet deskew: the host side of the drop-in boundary.

Mirrors the ``biahub`` names shrimPy imports:

* ``get_deskewed_data_shape``  -- ``shrimpy/preprocessing.py:226-231``, ``scripts/measure_psf.py:230-234``
* ``fast_deskew_zyx``          -- ``shrimpy/preprocessing.py:408-413`` (torch tensor in, tensor out, same device)
* ``deskew_data``              -- ``scripts/measure_psf.py:239-246`` (numpy in, numpy out)

Parameter NAMES are part of the contract: shrimPy selects keyword arguments by
``inspect.signature`` (``shrimpy/preprocessing.py:44-56``), so a parameter with
another name would be dropped silently.

All arithmetic on voxels runs in the CUDA library behind the C-ABI
(``include/shrimpy_b200.h``); this module only computes the float64 geometry,
allocates through torch and passes raw pointers plus the current stream.
There is no CPU or PyTorch fallback.
"""

from __future__ import annotations

import ctypes
import math
import os
import threading
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple, Union

import numpy as np

from . import _cabi

__all__ = [
    "DeskewGeometry",
    "deskew_geometry",
    "get_deskewed_data_shape",
    "fast_deskew_zyx",
    "deskew_zyx",
    "empty_deskewed",
    "deskew_data",
    "deskew_window",
    "window_needs",
    "HostPipeline",
]


@dataclass(frozen=True)
class DeskewGeometry:
    """Float64 geometry of one deskew: shapes plus the affine row mapping (o0, o2) to scan index."""

    raw_shape: Tuple[int, int, int]
    out_shape: Tuple[int, int, int]       # (ceil(Y/n), X, Xp)
    voxel_size: Tuple[float, float, float]
    n_avg: int
    m00: float                            # -r*cos(theta)
    m02: float                            # r
    shift: float                          # Z_shift (0 with keep_overhang)

    @property
    def unaveraged_shape(self) -> Tuple[int, int, int]:
        return (self.raw_shape[1], self.raw_shape[2], self.out_shape[2])

    def matrix(self) -> np.ndarray:
        """4x4 output-index -> input-index matrix (scipy convention) of the un-averaged resample."""
        _, Y, X = self.raw_shape
        return np.array([[self.m00, 0.0, self.m02, self.shift],
                         [-1.0, 0.0, 0.0, Y - 1.0],
                         [0.0, -1.0, 0.0, X - 1.0],
                         [0.0, 0.0, 0.0, 1.0]])

    @property
    def algorithmic_bytes(self) -> Tuple[int, int]:
        """(input voxels, output voxels) -- every input read once, every output written once."""
        return (int(np.prod(self.raw_shape)), int(np.prod(self.out_shape)))


def deskew_geometry(raw_data_shape: Sequence[int], ls_angle_deg: float, px_to_scan_ratio: float,
                    keep_overhang: bool, average_n_slices: int = 1, pixel_size_um: float = 1) -> DeskewGeometry:
    """All scalar geometry in float64, with numpy's ``cos``/``sin`` like the upstream Python.

    Xp = ceil(Z/r + Y cos(theta)) with the overhang kept, ceil(Z/r - Y cos(theta)) without;
    Z_shift = 0 resp. floor(Y cos(theta) r)  (SURVEY.md section 8 a2/a4).
    """
    if len(raw_data_shape) != 3:
        raise ValueError(f"raw_data_shape must be (Z, Y, X), got {tuple(raw_data_shape)}")
    Z, Y, X = (int(s) for s in raw_data_shape)
    n = int(average_n_slices)
    if min(Z, Y, X) <= 0 or n <= 0:
        raise ValueError(f"non-positive size in raw_data_shape={tuple(raw_data_shape)} / average_n_slices={n}")
    if not px_to_scan_ratio > 0:
        raise ValueError("px_to_scan_ratio must be positive")
    # numpy's cos / sin, as the upstream Python evaluates them; everything after that is the C-ABI's one geometry
    # routine, so a C host that passes the same two numbers gets this geometry bit for bit (include/shrimpy_b200.h)
    theta = ls_angle_deg * np.pi / 180
    shape = (ctypes.c_int64 * 3)()
    vox = (ctypes.c_double * 3)()
    row = (ctypes.c_double * 3)()
    _cabi.check(_cabi.lib().shrimpy_deskew_geometry_trig(
        Z, Y, X, float(np.cos(theta)), float(np.sin(theta)), float(px_to_scan_ratio), int(bool(keep_overhang)), n,
        float(pixel_size_um), shape, vox, row))
    voxel_size = (float(vox[0]), pixel_size_um, pixel_size_um)     # the caller's own objects for the two plain copies
    return DeskewGeometry(raw_shape=(Z, Y, X), out_shape=(int(shape[0]), int(shape[1]), int(shape[2])),
                          voxel_size=voxel_size, n_avg=n, m00=float(row[0]), m02=float(row[1]), shift=float(row[2]))


def get_deskewed_data_shape(raw_data_shape: Sequence[int], ls_angle_deg: float, px_to_scan_ratio: float,
                            keep_overhang: bool, average_n_slices: int = 1, pixel_size_um: float = 1):
    """``(deskewed_shape_zyx, voxel_size_zyx)`` -- see :func:`deskew_geometry`."""
    g = deskew_geometry(raw_data_shape, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices,
                        pixel_size_um)
    return g.out_shape, g.voxel_size


# ---------------------------------------------------------------------------------------------
# device path (torch tensors)
# ---------------------------------------------------------------------------------------------

def _torch():
    import torch

    return torch


def _require_cuda(torch, device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("shrimpy_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda" if device is None else device)
    if dev.type != "cuda":
        raise RuntimeError(f"shrimpy_b200 computes on CUDA devices only, got device={device!r}; "
                           "there is no CPU fallback")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _device_dtype(torch, t):
    """Map a tensor dtype to (C-ABI dtype code, tensor in a supported dtype)."""
    if t.dtype == torch.uint16:
        return _cabi.U16, t
    if t.dtype == torch.float32:
        return _cabi.F32, t
    return _cabi.F32, t.to(torch.float32)   # same cast shrimpy/preprocessing.py:316 applies


def _resolve_cval(torch, raw, code, cval, stream) -> float:
    if cval is not None:
        return float(cval)
    # scipy-generation default: pad with min(raw); reduced on the device.  The reduction walks `numel` consecutive
    # elements, so a strided view (an X chunk, a row-padded stack) is compacted first: its minimum is over the view's
    # own voxels, like `raw.min()` in deskew_data.
    dense = raw if raw.is_contiguous() else raw.contiguous()
    slot = torch.empty(1, dtype=torch.float32, device=raw.device)
    _cabi.check(_cabi.lib().shrimpy_min_device(dense.data_ptr(), code, dense.numel(), slot.data_ptr(), stream))
    return float(slot.item())


def _kernel_view(raw):
    """``raw`` in a layout the kernels address -- unit stride along x, rows and slices that do not overlap -- without a
    copy when it already is (C-contiguous stacks, X-chunk views, padded rows), compacted otherwise (transposed views,
    and broadcast views such as ``plane[None].expand(Z, Y, X)`` whose zero strides the C-ABI would read as "use the
    contiguous default").  Returns ``(tensor, stride_z, stride_y)`` in elements, never 0."""
    Z, Y, X = raw.shape
    sz, sy, sx = raw.stride()
    ok = raw.numel() > 0 and (sx == 1 or X == 1) and (Y == 1 or sy >= X) and (Z == 1 or sz >= (Y - 1) * (sy if Y > 1 else 0) + X)
    if not ok:
        raw = raw.contiguous()
        return raw, Y * X, X
    sy = sy if Y > 1 else X
    sz = sz if Z > 1 else Y * sy
    return raw, sz, sy


def _check_out(torch, out, shape, device, what="out"):
    P, X, C = shape
    if (tuple(out.shape) != tuple(shape) or out.dtype != torch.float32 or out.device != device
            or (out.numel() and ((C > 1 and out.stride(2) != 1) or (X > 1 and out.stride(1) < C)
                                 or (P > 1 and out.stride(0) < (X - 1) * (out.stride(1) if X > 1 else 0) + C)))):
        raise ValueError(f"{what} must be a float32 tensor of shape {tuple(shape)} on {device} with unit stride along "
                         "the last axis and non-overlapping rows and planes (see empty_deskewed)")
    s1 = out.stride(1) if X > 1 else C
    s0 = out.stride(0) if P > 1 else X * s1
    return s0, s1


def deskew_zyx(raw_data, ls_angle_deg: float, px_to_scan_ratio: float, keep_overhang: bool,
               average_n_slices: int = 1, cval: Optional[float] = 0.0, out=None, kernel: str = "auto",
               value_range=None, scale=None):
    """Deskew a CUDA tensor ``(Z, Y, X)`` (uint16 or float32) into float32 ``(ceil(Y/n), X, Xp)``.

    Runs asynchronously on torch's current stream; ``out`` may be a preallocated float32 tensor of the
    deskewed shape, contiguous or with padded rows (``empty_deskewed``).
    ``value_range``: an optional float32 CUDA tensor of two elements that receives ``(min, max)`` of the deskewed volume, reduced inside the deskew
    kernel (the range the tracking step needs next, ``shrimpy/dynatrack/tracking.py:583-584``; pass it
    on to ``reductions.percentile(..., value_range=...)``).  ``scale``: optional flat-field scale field
    ``(Y, X)`` applied in the same pass (see ``flatfield.deskew_flat_field_zyx``); only with ``value_range``.
    """
    torch = _torch()
    if raw_data.dim() != 3:
        raise ValueError(f"raw_data must be (Z, Y, X), got shape {tuple(raw_data.shape)}")
    if raw_data.device.type != "cuda":
        raise RuntimeError("deskew_zyx expects a CUDA tensor")
    g = deskew_geometry(tuple(raw_data.shape), ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices)
    code, raw = _device_dtype(torch, raw_data)
    raw, raw_sz, raw_sy = _kernel_view(raw)
    with torch.cuda.device(raw.device):
        stream = torch.cuda.current_stream().cuda_stream
        if out is None:
            out = torch.empty(g.out_shape, dtype=torch.float32, device=raw.device)
        out_s0, out_s1 = _check_out(torch, out, g.out_shape, raw.device)
        if out.numel() == 0:
            return out
        if not out.is_contiguous() and (value_range is not None or scale is not None):
            raise ValueError("value_range / scale need a contiguous out")
        fill = _resolve_cval(torch, raw, code, cval, stream)
        Z, Y, X = g.raw_shape
        if value_range is not None:
            if (value_range.dtype != torch.float32 or value_range.numel() != 2 or not value_range.is_contiguous()
                    or value_range.device != raw.device):
                raise ValueError(f"value_range must be a contiguous float32 tensor of 2 elements on {raw.device}")
            if scale is not None and (tuple(scale.shape) != (Y, X) or scale.dtype != torch.float32
                                      or not scale.is_contiguous()):
                raise ValueError(f"scale must be a contiguous float32 tensor of shape {(Y, X)}")
            _cabi.check(_cabi.lib().shrimpy_deskew_range_device(
                raw.data_ptr(), code, scale.data_ptr() if scale is not None else None, out.data_ptr(),
                value_range.data_ptr(), Z, Y, X, g.out_shape[2], g.n_avg, g.m00, g.m02, g.shift, fill,
                raw_sz, raw_sy, _cabi.KERNELS[kernel], stream))
            return out
        if scale is not None:
            raise ValueError("scale is only taken together with value_range; use flatfield.deskew_flat_field_zyx")
        _cabi.check(_cabi.lib().shrimpy_deskew_device(
            raw.data_ptr(), code, out.data_ptr(), Z, Y, X, g.out_shape[2], g.n_avg, g.m00, g.m02, g.shift, fill,
            raw_sz, raw_sy, out_s0, out_s1, _cabi.KERNELS[kernel], stream))
    return out


def empty_deskewed(g: DeskewGeometry, device, row_align: int = 8):
    """Uninitialised float32 tensor of the deskewed shape whose rows start on ``row_align``-float boundaries (default
    8 floats = one 32-byte DRAM sector): a view ``buf[:, :, :Xp]`` of a buffer with a padded last axis, for ``out=``.

    Why: the deskewed rows are ``Xp`` floats long and ``Xp`` is usually odd (1279 for the mantis FOV), so in a
    contiguous result every row starts at a different offset inside a sector and the kernel's 128-byte row stores
    leave partial sectors behind.  Measured on B200 (``tools/probe/padded_out_probe.py``): with padded rows the
    write-dominated ``average_n_slices=1`` deskews run 0.82 -> 0.64 ms and 1.18 -> 0.88 ms (keep_overhang), the
    ``n=3`` case 0.301 -> 0.291 ms.  The values are identical; only the strides differ, so a consumer that needs a
    contiguous tensor should not use this (the copy costs more than it saves).
    """
    torch = _torch()
    P, X, Xp = g.out_shape
    pitch = -(-Xp // row_align) * row_align
    return torch.empty((P, X, pitch), dtype=torch.float32, device=device)[:, :, :Xp]


def fast_deskew_zyx(raw_data, ls_angle_deg: float, px_to_scan_ratio: float, keep_overhang: bool,
                    average_n_slices: int = 1, cval: Optional[float] = 0.0):
    """Drop-in for ``biahub.deskew.fast_deskew_zyx`` (``shrimpy/preprocessing.py:408-413``).

    The result lives on the input's device.  A CUDA tensor is processed in place
    on the current stream; a CPU tensor is streamed through the GPU with the host
    pipeline and comes back as a CPU tensor (the arithmetic never runs on the CPU).
    """
    torch = _torch()
    if not isinstance(raw_data, torch.Tensor):
        raise TypeError("fast_deskew_zyx expects a torch.Tensor; use deskew_data for numpy arrays")
    if raw_data.device.type == "cuda":
        return deskew_zyx(raw_data, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices, cval)
    if raw_data.device.type != "cpu":
        raise RuntimeError(f"unsupported device {raw_data.device}")
    host = raw_data.detach()
    if host.dtype not in (torch.uint16, torch.float32):
        host = host.to(torch.float32)
    result = deskew_data(host.contiguous().numpy(), ls_angle_deg, px_to_scan_ratio, keep_overhang,
                         average_n_slices, cval=cval)
    return torch.from_numpy(result)


def window_needs(g: DeskewGeometry, p_begin: int, p_count: int, c_begin: int, c_count: int):
    """Raw rows ``[y0, y1)`` and scan slices ``[z0, z1)`` that an output window reads."""
    yr = (ctypes.c_int32 * 2)()
    zr = (ctypes.c_int32 * 2)()
    Z, Y, _ = g.raw_shape
    _cabi.check(_cabi.lib().shrimpy_deskew_window_needs(Z, Y, g.n_avg, g.m00, g.m02, g.shift, p_begin, p_count,
                                                        c_begin, c_count, yr, zr))
    return (int(yr[0]), int(yr[1])), (int(zr[0]), int(zr[1]))


def deskew_window(raw_slab, g: DeskewGeometry, *, p_begin: int, p_count: int, c_begin: int, c_count: int,
                  y_origin: int, z_origin: int, cval: float = 0.0, out=None, kernel: str = "auto"):
    """Deskew one output window from a raw slab (CUDA tensor holding rows/slices from the origins on).

    Returns ``out[p_begin:p_begin+p_count, :, c_begin:c_begin+c_count]`` as a compact tensor;
    geometry ``g`` is that of the FULL stack so the voxels equal the un-windowed result bit for bit.
    """
    torch = _torch()
    is_tensor = isinstance(raw_slab, torch.Tensor)     # paged_stack.DeviceSlab: an address and a shape, always dense
    code, raw = _device_dtype(torch, raw_slab)
    if is_tensor:
        raw, raw_sz, raw_sy = _kernel_view(raw)
    else:
        raw_sz, raw_sy = raw.stride(0), raw.stride(1)
    Z, Y, X = g.raw_shape
    if raw.shape[2] != X:
        raise ValueError("slab must span the full raw X axis")
    with torch.cuda.device(raw.device):
        stream = torch.cuda.current_stream().cuda_stream
        if out is None:
            out = torch.empty((p_count, X, c_count), dtype=torch.float32, device=raw.device)
        out_s0, out_s1 = _check_out(torch, out, (p_count, X, c_count), raw.device)
        if out.numel() == 0:
            return out
        win = _cabi.Window(p_begin, p_count, c_begin, c_count, y_origin, raw.shape[1], z_origin, raw.shape[0])
        _cabi.check(_cabi.lib().shrimpy_deskew_window_device(
            raw.data_ptr(), code, out.data_ptr(), Z, Y, X, g.out_shape[2], g.n_avg, g.m00, g.m02, g.shift,
            float(cval), raw_sz, raw_sy, out_s0, out_s1, ctypes.byref(win),
            _cabi.KERNELS[kernel], stream))
    return out


# ---------------------------------------------------------------------------------------------
# host path (numpy arrays)
# ---------------------------------------------------------------------------------------------

class HostPipeline:
    """Owner of one ``shrimpy_pipeline`` (three streams + device slab buffers on one GPU)."""

    def __init__(self, device: int = 0, device_bytes_budget: int = 0):
        handle = ctypes.c_void_p()
        _cabi.check(_cabi.lib().shrimpy_pipeline_create(int(device), int(device_bytes_budget), ctypes.byref(handle)))
        self._handle = handle
        self.device = int(device)

    def close(self) -> None:
        if self._handle:
            _cabi.lib().shrimpy_pipeline_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def stats(self) -> dict:
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        _cabi.check(_cabi.lib().shrimpy_pipeline_stats(self._handle, ctypes.byref(a), ctypes.byref(b),
                                                       ctypes.byref(c)))
        return {"launches": a.value, "h2d_bytes": b.value, "d2h_bytes": c.value}

    def deskew(self, raw: np.ndarray, g: DeskewGeometry, cval: float, out: Optional[np.ndarray] = None) -> np.ndarray:
        """``raw`` C-contiguous uint16/float32 ``(Z,Y,X)``; ``out`` C-contiguous float32 of ``g.out_shape``."""
        if raw.dtype == np.uint16:
            code = _cabi.U16
        elif raw.dtype == np.float32:
            code = _cabi.F32
        else:
            raise TypeError(f"host deskew takes uint16 or float32, got {raw.dtype}")
        if not raw.flags.c_contiguous or tuple(raw.shape) != g.raw_shape:
            raise ValueError("raw must be C-contiguous with the geometry's shape")
        if out is None:
            out = np.empty(g.out_shape, dtype=np.float32)
        elif out.dtype != np.float32 or not out.flags.c_contiguous or tuple(out.shape) != g.out_shape:
            raise ValueError(f"out must be C-contiguous float32 of shape {g.out_shape}")
        Z, Y, X = g.raw_shape
        _cabi.check(_cabi.lib().shrimpy_deskew_host(
            self._handle, raw.ctypes.data, code, out.ctypes.data, Z, Y, X, g.out_shape[2], g.n_avg,
            g.m00, g.m02, g.shift, float(cval)))
        return out


_pipelines: dict = {}
_pipelines_lock = threading.Lock()


def _pipeline_for(index: int) -> HostPipeline:
    """One pipeline per (device, calling thread): two threads that call ``deskew_data`` on the same GPU (an IO pool
    over positions) each stream through their own slots and streams and overlap; a pipeline that is nevertheless
    shared is serialised by the mutex inside it (``csrc/pipeline.cu``)."""
    key = (index, threading.get_ident())
    with _pipelines_lock:
        pipe = _pipelines.get(key)
        if pipe is None:
            alive = {t.ident for t in threading.enumerate()}
            for k in [k for k in _pipelines if k[1] not in alive]:     # threads that have ended: free their device slabs
                _pipelines.pop(k).close()
            pipe = _pipelines[key] = HostPipeline(index)
    return pipe


def deskew_data(raw_data: np.ndarray, ls_angle_deg: float, px_to_scan_ratio: float, keep_overhang: bool,
                average_n_slices: int = 1, device: Union[str, int, None] = "cuda", cval: Optional[float] = 0.0,
                out: Optional[np.ndarray] = None) -> np.ndarray:
    """Drop-in for ``biahub.analysis.deskew.deskew_data`` (``scripts/measure_psf.py:239-246``).

    numpy ``(Z, Y, X)`` in, float32 numpy ``(ceil(Y/n), X, Xp)`` out.  uint16 and
    float32 stacks are streamed as they are; any other dtype is cast to float32
    first (what ``shrimpy/preprocessing.py:316`` does).  ``cval=None`` pads with
    ``min(raw)`` (the scipy-generation default), otherwise with ``cval``.
    """
    torch = _torch()
    dev = _require_cuda(torch, device if not isinstance(device, int) else f"cuda:{device}")
    raw = np.asarray(raw_data)
    if raw.ndim != 3:
        raise ValueError(f"raw_data must be (Z, Y, X), got shape {raw.shape}")
    if raw.dtype not in (np.uint16, np.float32):
        raw = raw.astype(np.float32)
    raw = np.ascontiguousarray(raw)
    g = deskew_geometry(raw.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices)
    if math.prod(g.out_shape) == 0:
        return np.empty(g.out_shape, dtype=np.float32)
    fill = float(raw.min()) if cval is None else float(cval)
    if out is None:
        out = _empty_pinned_result(torch, g.out_shape)
    return _pipeline_for(dev.index).deskew(raw, g, fill, out)


class _PinnedBlock:
    """One page-locked block handed out as the base object of a result array; returns itself to the pool when the
    caller drops the array."""

    def __init__(self, ptr: int, nbytes: int, shape):
        self.ptr, self.nbytes = ptr, nbytes
        self.__array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": "<f4", "data": (ptr, False),
                                    "version": 3}

    def __del__(self):
        try:
            _pinned_pool.give_back(self.ptr, self.nbytes)
        except Exception:       # interpreter shutdown: the library or the pool may be gone already
            pass


class _PinnedPool:
    """Page-locked result blocks of exactly the sizes asked for, cached by size.  Live and cached blocks together
    never exceed ``SHRIMPY_PINNED_RESULT_BYTES`` (default 4 GiB): page-locking new memory is slow (hundreds of ms per
    GB), so a loop that consumes each result pins once and reuses the block, and a caller that keeps every result
    (``scripts/measure_psf.py:239-249`` collects the chunks in a list) gets ordinary arrays, filled through the
    pipeline's staging ring, for whatever does not fit."""

    def __init__(self):
        self.lock = threading.Lock()
        self.free: dict = {}            # nbytes -> [ptr, ...]
        self.total = 0                  # bytes page-locked through this pool, live + cached

    @staticmethod
    def budget() -> int:
        return int(os.environ.get("SHRIMPY_PINNED_RESULT_BYTES", 4 << 30))

    def take(self, shape) -> Optional[np.ndarray]:
        nbytes = 4 * int(np.prod(shape))
        with self.lock:
            cached = self.free.get(nbytes)
            ptr = cached.pop() if cached else None
            if ptr is None:
                # make room by releasing cached blocks of other sizes before giving up
                while self.total + nbytes > self.budget() and any(self.free.values()):
                    size = next(k for k, v in self.free.items() if v)
                    _cabi.lib().shrimpy_host_free(ctypes.c_void_p(self.free[size].pop()))
                    self.total -= size
                if self.total + nbytes > self.budget():
                    return None
                self.total += nbytes
        if ptr is None:
            out = ctypes.c_void_p()
            if _cabi.lib().shrimpy_host_alloc(nbytes, ctypes.byref(out)) != _cabi.OK or not out.value:
                with self.lock:
                    self.total -= nbytes
                return None                 # the host refuses to lock that much: an ordinary array will do
            ptr = int(out.value)
        return np.asarray(_PinnedBlock(ptr, nbytes, shape))

    def give_back(self, ptr: int, nbytes: int) -> None:
        with self.lock:
            self.free.setdefault(nbytes, []).append(ptr)


_pinned_pool = _PinnedPool()


def _empty_pinned_result(torch, shape) -> np.ndarray:
    """Result array of a numpy-in/numpy-out call: an ordinary float32 ndarray over a page-locked block of exactly
    its size (``shrimpy_host_alloc``), or a pageable one when the pool's budget is spent.

    The device-to-host copy is the stage that bounds the host path and it runs at the full PCIe
    rate straight into page-locked memory (B200: 29 ms per mantis channel against 47 ms through the pipeline's staging
    ring into a pageable array).  When the caller drops the array the block returns to the pool."""
    if math.prod(shape) == 0:
        return np.empty(shape, dtype=np.float32)
    out = _pinned_pool.take(shape)
    return out if out is not None else np.empty(shape, dtype=np.float32)
