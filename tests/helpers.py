"""Shared helpers for the test-suite (synthetic stacks, tolerances)."""

from __future__ import annotations

import numpy as np

# north_star: float outputs within 1e-3 of the dynamic range for linear interpolation.  Both sides
# evaluate one lerp per tap in >= float32, so the observed error is ~1e-7; the suite holds the
# kernels to 2e-6 of range and records the contract value next to it.
CONTRACT_TOL = 1e-3
TIGHT_TOL = 2e-6


def synthetic_stack(shape, seed, dtype=np.uint16):
    """Seeded uint16 stack uniform in [100, 60000] or a float32 variant."""
    rng = np.random.default_rng(seed)
    raw = rng.integers(100, 60000, size=shape, dtype=np.uint16)
    if dtype == np.uint16:
        return raw
    return (raw.astype(np.float32) + rng.random(shape, dtype=np.float32)).astype(dtype)


def assert_close_range(got, want, tol=TIGHT_TOL, what=""):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    assert got.dtype == want.dtype == np.float32, f"{what}: dtypes {got.dtype} {want.dtype}"
    rng = float(want.max() - want.min()) if want.size else 0.0
    rng = rng if rng > 0 else 1.0
    err = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)))) if want.size else 0.0
    assert err <= tol * rng, f"{what}: max|err| {err:.3e} > {tol:.1e} x range {rng:.3e}"
    return err / rng
