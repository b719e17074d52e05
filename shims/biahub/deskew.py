"""Shim for ``biahub.deskew`` (imported at ``shrimpy/preprocessing.py:226,408``)."""

from shrimpy_b200.deskew import deskew_data, fast_deskew_zyx, get_deskewed_data_shape

__all__ = ["deskew_data", "fast_deskew_zyx", "get_deskewed_data_shape"]
