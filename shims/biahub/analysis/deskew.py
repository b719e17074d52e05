"""Shim for ``biahub.analysis.deskew`` (imported at ``scripts/measure_psf.py:15-17``)."""

from shrimpy_b200.deskew import deskew_data, get_deskewed_data_shape

__all__ = ["deskew_data", "get_deskewed_data_shape"]
