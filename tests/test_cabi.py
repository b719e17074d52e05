"""The C-ABI library loads on a CPU-only box and exports everything include/shrimpy_b200.h declares."""

import ctypes
import math
import re
from pathlib import Path

import numpy as np
import pytest

import shrimpy_b200 as sb
from shrimpy_b200 import _cabi

ROOT = Path(__file__).resolve().parent.parent


def _declared_functions():
    text = (ROOT / "include" / "shrimpy_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(shrimpy_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = _cabi.lib()
    declared = _declared_functions()
    assert len(declared) >= 12
    assert set(declared) == set(_cabi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_no_fallback_message():
    assert _cabi.lib().shrimpy_abi_version() == 1
    assert _cabi.launch_count() >= 0


def test_c_geometry_matches_python():
    """One code path: Python's geometry IS the C-ABI's trig entry point fed numpy's cos / sin, so they agree with ==;
    the angle entry point (libm cos / sin) agrees with == whenever libm and numpy return the same two numbers."""
    lib = _cabi.lib()
    rng = np.random.default_rng(1)
    libm_differs = 0
    for _ in range(300):
        Z, Y, X = (int(v) for v in rng.integers(1, 900, 3))
        th, r = round(float(rng.uniform(1, 45)), 2), round(float(rng.uniform(0.2, 1.2)), 3)
        keep, n = int(rng.integers(0, 2)), int(rng.integers(1, 5))
        shape = (ctypes.c_int64 * 3)()
        vox = (ctypes.c_double * 3)()
        row = (ctypes.c_double * 3)()
        theta = th * np.pi / 180
        ct, st = float(np.cos(theta)), float(np.sin(theta))
        assert lib.shrimpy_deskew_geometry_trig(Z, Y, X, ct, st, r, keep, n, 0.116, shape, vox, row) == 0
        g = sb.deskew_geometry((Z, Y, X), th, r, bool(keep), n, 0.116)
        assert tuple(shape) == g.out_shape
        assert (row[0], row[1], row[2]) == (g.m00, g.m02, g.shift)
        assert tuple(vox) == g.voxel_size
        # the same numbers as the upstream numpy expressions, bit for bit
        assert g.m00 == -r * np.cos(theta) and g.voxel_size[0] == n * np.sin(theta) * 0.116
        assert g.shift == (0 if keep else int(np.floor(Y * np.cos(theta) * r)))
        assert g.out_shape[2] == max(0, int(np.ceil(Z / r + Y * np.cos(theta)) if keep else np.ceil(Z / r - Y * np.cos(theta))))
        assert lib.shrimpy_deskew_geometry(Z, Y, X, th, r, keep, n, 0.116, shape, vox, row) == 0
        if math.cos(theta) == ct and math.sin(theta) == st:
            assert tuple(shape) == g.out_shape and (row[0], row[1], row[2]) == (g.m00, g.m02, g.shift)
            assert tuple(vox) == g.voxel_size
        else:
            libm_differs += 1
            assert row[0] == pytest.approx(g.m00, rel=2e-16) and tuple(vox) == pytest.approx(g.voxel_size, rel=2e-16)
    assert libm_differs <= 300    # informational: how often libm and numpy disagree in the last bit on this machine


def test_argument_errors_surface_without_a_gpu():
    lib = _cabi.lib()
    rc = lib.shrimpy_deskew_device(None, 7, None, 4, 4, 4, 4, 1, -0.3, 0.39, 0.0, 0.0, 0, 0, 0, 0, 0, None)
    assert rc == _cabi.EINVAL and b"raw_dtype" in lib.shrimpy_last_error()
    rc = lib.shrimpy_deskew_device(None, 0, None, 0, 4, 4, 4, 1, -0.3, 0.39, 0.0, 0.0, 0, 0, 0, 0, 0, None)
    assert rc == _cabi.EINVAL
    with pytest.raises(_cabi.ShrimpyB200Error):
        _cabi.check(rc)
    yr = (ctypes.c_int32 * 2)()
    zr = (ctypes.c_int32 * 2)()
    assert lib.shrimpy_deskew_window_needs(600, 300, 3, -0.3377, 0.39, 101.0, 10, 5, 128, 256, yr, zr) == 0
    assert (yr[0], yr[1]) == (300 - 45, 300 - 30)
    assert 0 <= zr[0] < zr[1] <= 600


def test_window_needs_cover_every_tap():
    """Host mirror of the kernel arithmetic: the reported slab really contains every tap."""
    g = sb.deskew_geometry((120, 31, 8), 30.0, 0.39, False, 3)
    Yn, _, Xp = g.out_shape
    for p0, pc, c0, cc in [(0, Yn, 0, Xp), (2, 3, 17, 40), (Yn - 1, 1, Xp - 5, 5), (0, 1, 0, 1)]:
        (y0, y1), (z0, z1) = sb.window_needs(g, p0, pc, c0, cc)
        o0 = np.arange(3 * p0, min(3 * (p0 + pc), 31))
        assert y0 == 31 - 1 - o0.max() and y1 == 31 - o0.min()
        z_in = (g.shift + o0[:, None] * g.m00) + np.arange(c0, c0 + cc)[None, :] * g.m02
        inside = (z_in >= 0) & (z_in <= 119)
        if inside.any():
            lo = np.floor(z_in[inside]).min()
            hi = np.minimum(np.floor(z_in[inside]) + 1, 119).max()
            assert z0 <= lo and hi < z1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setenv("SHRIMPY_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        _cabi.lib()


def test_no_cpu_path():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this test is about the CPU-only box")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sb.deskew_data(np.zeros((4, 4, 8), np.uint16), 30.0, 0.39, True)
    with pytest.raises(RuntimeError):
        sb.fast_deskew_zyx(torch.zeros((4, 4, 8)), 30.0, 0.39, True)


def test_host_result_allocation_falls_back_to_pageable_memory():
    """``deskew_data`` takes its result from torch's pinned-host cache; a host that cannot lock the pages (here: no
    CUDA runtime at all) gets an ordinary array of the same shape and dtype."""
    import torch

    from shrimpy_b200.deskew import _empty_pinned_result

    out = _empty_pinned_result(torch, (3, 4, 5))
    assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == (3, 4, 5)
    assert out.flags.c_contiguous and out.flags.writeable
