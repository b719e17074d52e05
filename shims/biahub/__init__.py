"""``biahub``-named shim: routes the three imports shrimPy performs to ``shrimpy_b200``.

shrimPy imports (``shrimpy/preprocessing.py:138,226,408``; ``scripts/measure_psf.py:15-17``):
``biahub.settings.DeskewSettings``, ``biahub.deskew.{fast_deskew_zyx,get_deskewed_data_shape}``,
``biahub.analysis.deskew.{deskew_data,get_deskewed_data_shape}``.  Put this
directory's parent (``shims/``) on ``PYTHONPATH`` or call
``shrimpy_b200.install_biahub_shim()``.
"""

__shrimpy_b200_shim__ = True
