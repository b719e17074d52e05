"""Shim for ``biahub.settings`` (imported at ``shrimpy/preprocessing.py:138``)."""

from shrimpy_b200.settings import DeskewSettings

__all__ = ["DeskewSettings"]
