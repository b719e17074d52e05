"""shrimpy_b200 -- B200-native (sm_100a) light-sheet deskew and affine registration resample.

A from-scratch replacement for the one data-parallel hot path behind
czbiohub-sf/shrimPy (``shrimpy/preprocessing.py:406-417`` and the ``biahub``
deskew helpers it calls).  Public names mirror the ones shrimPy imports; see
``INTEGRATION.md`` for how to switch the reference over.
"""

from .deskew import (DeskewGeometry, HostPipeline, deskew_data, deskew_geometry, deskew_window, deskew_zyx,
                     empty_deskewed, fast_deskew_zyx, get_deskewed_data_shape, window_needs)
from .settings import DeskewSettings
from . import flatfield, reductions, register  # noqa: E402,F401  (light: torch is imported lazily inside the calls)

__version__ = "0.1.0"

__all__ = [
    "DeskewGeometry",
    "DeskewSettings",
    "HostPipeline",
    "deskew_data",
    "deskew_geometry",
    "deskew_window",
    "deskew_zyx",
    "empty_deskewed",
    "fast_deskew_zyx",
    "get_deskewed_data_shape",
    "window_needs",
    "install_biahub_shim",
    "flatfield",
    "reductions",
    "register",
]


def install_biahub_shim() -> None:
    """Make ``import biahub.deskew`` / ``biahub.settings`` / ``biahub.analysis.deskew`` resolve to this package.

    Equivalent to putting ``shims/`` on ``PYTHONPATH``; refuses to shadow a real ``biahub``.
    """
    import importlib.util
    import sys
    from pathlib import Path

    if "biahub" in sys.modules and not getattr(sys.modules["biahub"], "__shrimpy_b200_shim__", False):
        raise RuntimeError("a real biahub is already imported; not shadowing it")
    shim_root = str(Path(__file__).resolve().parent.parent / "shims")
    if shim_root not in sys.path:
        sys.path.insert(0, shim_root)
    spec = importlib.util.find_spec("biahub")
    if spec is None or not str(spec.origin).startswith(shim_root):
        raise RuntimeError("could not place the biahub shim first on sys.path")
