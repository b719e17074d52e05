"""CPU oracle of the bright-field flat-field correction.  TEST INFRASTRUCTURE ONLY.

Restates ``_LabelfreePreprocessor._flat_field_BF`` (``/root/reference/shrimpy/preprocessing.py:385-404``)
in numpy: per-pixel median over Z (``numpy.median`` == ``Tensor.quantile(0.5)`` with linear interpolation),
then ``volume / pattern * pattern.mean()`` in float32.

PINNED: unlike the deskew, this function lives in the reference tree itself and is pure torch, so the
committed golden vector ``tests/golden/flatfield.npz`` was produced by running the UNMODIFIED reference
method on CPU (``tests/golden/make_golden.py``); ``tests/test_oracle.py`` checks this restatement against it.
"""

from __future__ import annotations

import numpy as np


def flat_field_pattern(volume: np.ndarray) -> np.ndarray:
    return np.median(np.asarray(volume, dtype=np.float32), axis=0).astype(np.float32)


def flat_field_BF(volume: np.ndarray) -> np.ndarray:
    vol = np.asarray(volume, dtype=np.float32)
    pattern = flat_field_pattern(vol)
    return (vol / pattern * pattern.mean(dtype=np.float32)).astype(np.float32)
