"""Reductions over the deskewed volume that the tracking step performs right after the deskew
(SURVEY.md section 8f, rank 4).  Mirrors two pure-torch helpers of the reference:

* ``_percentile(img, percentile, nbins=256)``              -- ``shrimpy/dynatrack/tracking.py:572-595``
* ``_intensity_center_of_mass(img, background=0.0)``       -- ``shrimpy/dynatrack/tracking.py:598-649``

and the projection the tracker saves beside the centroid, ``(img - background).clamp_min(0).amax(dim=0)``
(``tracking.py:1447-1455``) -- ``max_projection``.

The reference makes ~10 full passes over the volume for the pair (min, max, histc, a float32 copy, subtract,
clamp, sum and three marginal sums); here it is three streaming passes (range, histogram, weighted sums),
each an HBM-bound kernel behind the C-ABI.  No CPU fallback.
"""

from __future__ import annotations

import numpy as np

from . import _cabi

__all__ = ["value_range", "percentile", "intensity_center_of_mass", "max_projection"]


def _f32_cuda(img):
    import torch

    if not isinstance(img, torch.Tensor) or img.device.type != "cuda":
        raise RuntimeError("reductions expect a CUDA tensor; there is no CPU fallback")
    if img.numel() == 0:
        raise ValueError("empty tensor")
    return torch, img.to(torch.float32).contiguous()


def value_range(img):
    """``(min, max)`` of a CUDA tensor as Python floats (one pass)."""
    torch, v = _f32_cuda(img)
    with torch.cuda.device(v.device):
        out = torch.empty(2, dtype=torch.float32, device=v.device)
        _cabi.check(_cabi.lib().shrimpy_minmax_device(v.data_ptr(), v.numel(), out.data_ptr(),
                                                      torch.cuda.current_stream().cuda_stream))
    lo, hi = out.tolist()
    return lo, hi


_range_pass = value_range     # (the keyword of the same name below shadows the function)


def percentile(img, percentile: float, nbins: int = 256, value_range=None) -> float:
    """Histogram estimate of a percentile (0-100): the upper edge of the bin where the CDF reaches it.

    ``value_range``: ``(min, max)`` of ``img`` when already known -- a pair of floats or the two-element tensor
    ``deskew_zyx(..., value_range=...)`` filled inside the deskew kernel; saves the range pass."""
    if nbins != 256:
        raise NotImplementedError("the histogram kernel has 256 bins, like the reference's default")
    torch, v = _f32_cuda(img)
    if value_range is None:
        vmin, vmax = _range_pass(v)
    else:
        vmin, vmax = (float(x) for x in (value_range.tolist() if hasattr(value_range, "tolist") else value_range))
    if vmax <= vmin:
        return vmin
    with torch.cuda.device(v.device):
        hist = torch.empty(256, dtype=torch.int64, device=v.device)
        _cabi.check(_cabi.lib().shrimpy_hist256_device(v.data_ptr(), v.numel(), vmin, vmax, hist.data_ptr(),
                                                       torch.cuda.current_stream().cuda_stream))
    counts = hist.cpu().numpy().astype(np.float64)
    cdf = (np.cumsum(counts) / counts.sum()).astype(np.float32)
    idx = min(int(np.searchsorted(cdf, np.float32(percentile / 100.0), side="left")), nbins - 1)
    return vmin + (idx + 1) * (vmax - vmin) / nbins


def intensity_center_of_mass(img, background: float = 0.0):
    """Intensity-weighted centre of mass ``(z, y, x)`` with weights ``max(v - background, 0)``; the geometric
    centre when there is no positive mass (as the reference does)."""
    torch, v = _f32_cuda(img)
    if v.dim() != 3:
        raise ValueError(f"expected a (Z, Y, X) volume, got {tuple(v.shape)}")
    Z, Y, X = v.shape
    with torch.cuda.device(v.device):
        sums = torch.empty(4, dtype=torch.float64, device=v.device)
        _cabi.check(_cabi.lib().shrimpy_center_of_mass_device(v.data_ptr(), Z, Y, X, float(background), sums.data_ptr(),
                                                              torch.cuda.current_stream().cuda_stream))
    total, sz, sy, sx = sums.tolist()
    if total <= 0:
        return torch.tensor([(s - 1) / 2.0 for s in v.shape], device=v.device, dtype=torch.float32)
    return torch.tensor([sz / total, sy / total, sx / total], device=v.device, dtype=torch.float32)


def max_projection(img, background: float = 0.0):
    """``(img - background).clamp_min(0).amax(dim=0)`` of a ``(Z, Y, X)`` CUDA volume in one pass (bit-identical to the
    torch expression for finite data: the subtraction is the same float32 operation and the maximum does not depend
    on order; NaN voxels are skipped rather than propagated)."""
    torch, v = _f32_cuda(img)
    if v.dim() != 3:
        raise ValueError(f"expected a (Z, Y, X) volume, got {tuple(v.shape)}")
    Z, Y, X = v.shape
    with torch.cuda.device(v.device):
        out = torch.empty((Y, X), dtype=torch.float32, device=v.device)
        _cabi.check(_cabi.lib().shrimpy_zmax_projection_device(v.data_ptr(), Z, Y, X, float(background), out.data_ptr(),
                                                               torch.cuda.current_stream().cuda_stream))
    return out
