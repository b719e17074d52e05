"""Bright-field flat-field correction, the step in front of the deskew.

Stands in for ``_LabelfreePreprocessor._flat_field_BF`` (``shrimpy/preprocessing.py:385-404``)::

    static_pattern = volume.quantile(0.5, dim=0)        # per-pixel median over Z (numpy.median semantics)
    return volume / static_pattern * static_pattern.mean()

The correction is a per-pixel scale ``mean(pattern) / pattern[y, x]``.  ``flat_field_BF`` applies it as a
stand-alone pass; ``deskew_flat_field_zyx`` hands the scale field to the deskew kernel instead, which applies
it to the interpolated value of each tilt row (the deskew interpolates along z only, so the scale commutes
with it) -- the corrected float32 volume (2x the raw bytes) is never written.  Bright-field only, as upstream.
All arithmetic on voxels runs in the CUDA library; there is no CPU fallback.
"""

from __future__ import annotations

import ctypes
from typing import Optional

from . import _cabi
from .deskew import _device_dtype, _kernel_view, deskew_geometry

__all__ = ["flat_field_pattern", "flat_field_scale", "flat_field_BF", "deskew_flat_field_zyx"]


def _prep(volume):
    import torch

    if not isinstance(volume, torch.Tensor) or volume.device.type != "cuda":
        raise RuntimeError("flat-field functions expect a CUDA tensor; there is no CPU fallback")
    if volume.dim() != 3:
        raise ValueError(f"volume must be (Z, Y, X), got {tuple(volume.shape)}")
    code, vol = _device_dtype(torch, volume)
    vol, sz, sy = _kernel_view(vol)          # never a zero or overlapping stride (broadcast views are compacted)
    return torch, code, vol, sz, sy


def flat_field_pattern(volume):
    """Per-pixel median over the scan axis, float32 ``(Y, X)`` (``numpy.median`` / ``quantile(0.5)`` semantics)."""
    torch, code, vol, sz, sy = _prep(volume)
    Z, Y, X = vol.shape
    with torch.cuda.device(vol.device):
        pattern = torch.empty((Y, X), dtype=torch.float32, device=vol.device)
        _cabi.check(_cabi.lib().shrimpy_flatfield_pattern_device(
            vol.data_ptr(), code, pattern.data_ptr(), Z, Y, X, sz, sy,
            torch.cuda.current_stream().cuda_stream))
    return pattern


def flat_field_scale(volume):
    """Scale field ``mean(pattern) / pattern`` of a stack, float32 ``(Y, X)``."""
    torch = _prep(volume)[0]
    pattern = flat_field_pattern(volume)
    with torch.cuda.device(pattern.device):
        scale = torch.empty_like(pattern)
        scratch = torch.empty(1, dtype=torch.float64, device=pattern.device)
        _cabi.check(_cabi.lib().shrimpy_flatfield_scale_device(
            pattern.data_ptr(), pattern.numel(), scale.data_ptr(), scratch.data_ptr(),
            torch.cuda.current_stream().cuda_stream))
    return scale


def flat_field_BF(volume):
    """Drop-in for ``_LabelfreePreprocessor._flat_field_BF``: corrected float32 volume on the same device."""
    torch, code, vol, _, _ = _prep(volume)
    vol = vol.contiguous()
    scale = flat_field_scale(vol)
    Z, Y, X = vol.shape
    with torch.cuda.device(vol.device):
        out = torch.empty((Z, Y, X), dtype=torch.float32, device=vol.device)
        _cabi.check(_cabi.lib().shrimpy_flatfield_apply_device(
            vol.data_ptr(), code, scale.data_ptr(), out.data_ptr(), Z, Y, X, torch.cuda.current_stream().cuda_stream))
    return out


def deskew_flat_field_zyx(raw_data, ls_angle_deg: float, px_to_scan_ratio: float, keep_overhang: bool,
                          average_n_slices: int = 1, cval: float = 0.0, scale=None, out=None, kernel: str = "auto"):
    """Flat-field (bright-field) + deskew in one pass over the raw stack.

    ``scale`` may carry a precomputed scale field (e.g. from a previous timepoint of the same position);
    by default it is computed from ``raw_data`` itself, as the reference does per stack.
    """
    torch, code, raw, raw_sz, raw_sy = _prep(raw_data)
    g = deskew_geometry(tuple(raw.shape), ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices)
    Z, Y, X = g.raw_shape
    if scale is None:
        scale = flat_field_scale(raw)
    elif tuple(scale.shape) != (Y, X) or scale.dtype != torch.float32 or not scale.is_contiguous():
        raise ValueError(f"scale must be a contiguous float32 tensor of shape {(Y, X)}")
    with torch.cuda.device(raw.device):
        if out is None:
            out = torch.empty(g.out_shape, dtype=torch.float32, device=raw.device)
        elif tuple(out.shape) != g.out_shape or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 tensor of shape {g.out_shape}")
        if out.numel() == 0:
            return out
        _cabi.check(_cabi.lib().shrimpy_deskew_flatfield_device(
            raw.data_ptr(), code, scale.data_ptr(), out.data_ptr(), Z, Y, X, g.out_shape[2], g.n_avg,
            g.m00, g.m02, g.shift, float(cval), raw_sz, raw_sy,
            ctypes.cast(None, ctypes.POINTER(_cabi.Window)), _cabi.KERNELS[kernel],
            torch.cuda.current_stream().cuda_stream))
    return out
