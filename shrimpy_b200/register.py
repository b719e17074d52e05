"""Affine registration resample (label-free volume onto fluorescence space).

The operation is named by BASELINE.json (``configs[2]``) but is absent from the reference tree
(only prose at ``README.md:8``); upstream it is ``biahub``'s ``apply_affine_transform`` whose scipy
method is ``scipy.ndimage.affine_transform(zyx, matrix, output_shape, order=1, mode="constant",
cval=0)`` after ``nan_to_num``.  ``matrix`` is 4x4 (or 3x4),
output index -> input index, in ZYX voxel units.  The trilinear gather runs in the CUDA library
(``shrimpy_affine_device``); there is no CPU fallback.
"""

from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np

from . import _cabi

__all__ = ["apply_affine_transform", "affine_transform_zyx"]

# below this size the dense-only kernels are as fast as padding the rows for the streaming ones
_PAD_MIN_VOXELS = 1 << 22


def _matrix12(matrix) -> np.ndarray:
    M = np.asarray(matrix, dtype=np.float64)
    if M.shape == (4, 4):
        if not np.array_equal(M[3], [0.0, 0.0, 0.0, 1.0]):
            raise ValueError("the last row of a 4x4 affine must be (0, 0, 0, 1)")
        M = M[:3]
    if M.shape != (3, 4):
        raise ValueError(f"matrix must be 4x4 or 3x4, got {M.shape}")
    if not np.all(np.isfinite(M)):
        raise ValueError("matrix has non-finite entries")
    return np.ascontiguousarray(M)


def affine_transform_zyx(zyx_data, matrix, output_shape_zyx: Sequence[int], cval: float = 0.0,
                         nan_to_zero: bool = True, out=None):
    """Resample a CUDA float32 tensor ``(Z, Y, X)`` onto ``output_shape_zyx`` (asynchronous, current stream)."""
    import torch

    if zyx_data.device.type != "cuda":
        raise RuntimeError("affine_transform_zyx expects a CUDA tensor; there is no CPU fallback")
    if zyx_data.dim() != 3:
        raise ValueError(f"zyx_data must be (Z, Y, X), got {tuple(zyx_data.shape)}")
    vol = zyx_data if zyx_data.dtype == torch.float32 else zyx_data.to(torch.float32)
    vol = vol.contiguous()
    shape = tuple(int(s) for s in output_shape_zyx)
    if len(shape) != 3 or min(shape) < 0:
        raise ValueError(f"bad output shape {shape}")
    M = _matrix12(matrix)
    with torch.cuda.device(vol.device):
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=vol.device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 tensor of shape {shape}")
        if out.numel() == 0:
            return out
        if vol.numel() == 0:
            raise ValueError("empty input volume")
        Mp = M.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        stream = torch.cuda.current_stream().cuda_stream
        iz, iy, ix = vol.shape
        if ix % 4 != 0 and vol.numel() >= _PAD_MIN_VOXELS:
            # The plane-streaming kernels stage planes with TMA, which wants rows on 16-byte boundaries: pad the rows
            # once (one read + one write of the volume) and hand the strides over; TMA zero-fills beyond the logical X.
            ixp = (ix + 3) // 4 * 4
            padded = torch.empty((iz, iy, ixp), dtype=torch.float32, device=vol.device)
            padded[:, :, :ix].copy_(vol)
            code = _cabi.lib().shrimpy_affine_strided_device(
                padded.data_ptr(), out.data_ptr(), iz, iy, ix, iy * ixp, ixp, *shape, Mp, float(cval),
                int(bool(nan_to_zero)), stream)
            if code == _cabi.OK:
                return out
            if "not eligible" not in _cabi.lib().shrimpy_last_error().decode("utf-8", "replace"):
                _cabi.check(code)
        _cabi.check(_cabi.lib().shrimpy_affine_device(
            vol.data_ptr(), out.data_ptr(), iz, iy, ix, *shape, Mp, float(cval), int(bool(nan_to_zero)), stream))
    return out


def apply_affine_transform(zyx_data, matrix, output_shape_zyx: Sequence[int], cval: float = 0.0, device="cuda"):
    """numpy in -> numpy out, torch in -> torch out (same device); trilinear, constant ``cval`` outside."""
    import torch

    if isinstance(zyx_data, torch.Tensor):
        if zyx_data.device.type == "cuda":
            return affine_transform_zyx(zyx_data, matrix, output_shape_zyx, cval)
        return affine_transform_zyx(zyx_data.to(device), matrix, output_shape_zyx, cval).to(zyx_data.device)
    if not torch.cuda.is_available():
        raise RuntimeError("shrimpy_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    vol = torch.from_numpy(np.ascontiguousarray(zyx_data, dtype=np.float32)).to(device)
    return affine_transform_zyx(vol, matrix, output_shape_zyx, cval).cpu().numpy()
