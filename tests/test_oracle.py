"""The oracle checked against itself (scipy form vs closed form vs C), analytic known answers and the
committed golden vectors.  CPU only.  Parity status with respect to biahub: UNPINNED (see oracle/)."""

from pathlib import Path

import numpy as np
import pytest

from helpers import synthetic_stack
from oracle import c_oracle, deskew_oracle as o

GOLDEN = Path(__file__).parent / "golden"


def test_shapes_known_answers():
    # SURVEY.md section 8 a2 / 8c
    assert o.get_deskewed_data_shape((101, 256, 256), 30, 0.39, False)[0] == (256, 256, 38)
    assert o.get_deskewed_data_shape((101, 256, 256), 30, 0.39, True)[0] == (256, 256, 481)
    assert o.get_deskewed_data_shape((600, 300, 2048), 30, 0.39, False, 3)[0] == (100, 2048, 1279)
    assert o.get_deskewed_data_shape((4000, 300, 2048), 30, 0.39, True)[0] == (300, 2048, 10517)
    assert o.get_deskewed_data_shape((2, 3, 4), 36, 0.386, True)[0] == (3, 4, 8)
    _, vox = o.get_deskewed_data_shape((600, 300, 2048), 30, 0.39, False, 3, pixel_size_um=0.116)
    assert vox == pytest.approx((3 * 0.5 * 0.116, 0.116, 0.116))


def test_settings_rounding():
    assert o.round_settings(30.004, 0.1133, 0.174) == (30.0, 0.651)
    assert o.round_settings(30, px_to_scan_ratio=0.38961) == (30.0, 0.39)
    with pytest.raises(ValueError):
        o.round_settings(30, pixel_size_um=0.1)


GRID = [(shape, th, r, keep, n)
        for shape in [(23, 10, 7), (40, 11, 5)] for th in (30, 36) for r in (0.39, 0.651)
        for keep in (True, False) for n in (1, 2, 3)]


@pytest.mark.parametrize("shape,th,r,keep,n", GRID)
def test_three_statements_agree(shape, th, r, keep, n):
    """scipy form == numpy closed form == C restatement, bit for bit, incl. Y % n != 0."""
    raw = synthetic_stack(shape, seed=sum(shape) + n)
    a = o.deskew_data(raw, th, r, keep, n, cval=3.0)
    b = o.deskew_data_closed_form(raw, th, r, keep, n, cval=3.0)
    c = c_oracle.deskew_data(raw, th, r, keep, n, cval=3.0, threads=2)
    assert a.dtype == b.dtype == c.dtype == np.float32
    assert a.shape == o.get_deskewed_data_shape(shape, th, r, keep, n)[0]
    assert np.array_equal(a, b)
    assert np.array_equal(a, c)


def test_float32_input_c_oracle():
    raw = synthetic_stack((30, 9, 12), seed=2, dtype=np.float32)
    assert np.array_equal(o.deskew_data(raw, 30, 0.39, True, 2), c_oracle.deskew_data(raw, 30, 0.39, True, 2))


def test_constant_volume_and_inside_mask():
    Z, Y, X = 37, 6, 4
    raw = np.full((Z, Y, X), 500, dtype=np.uint16)
    for keep in (True, False):
        M = o.deskew_affine_matrix(raw.shape, 30, 0.39, keep)
        d = o.deskew_data(raw, 30, 0.39, keep, 1, cval=-9.0)
        o0 = np.arange(Y, dtype=np.float64)[:, None]
        o2 = np.arange(d.shape[2], dtype=np.float64)[None, :]
        z_in = (M[0, 3] + o0 * M[0, 0]) + o2 * M[0, 2]
        inside = (z_in >= 0) & (z_in <= Z - 1)
        assert np.array_equal(d[:, 0, :] == 500.0, inside)
        assert np.array_equal(d[:, 0, :] == -9.0, ~inside)


def test_strict_boundary_rule():
    """z_in = 0 and z_in = Z-1 are inside; anything beyond is cval, with no blending."""
    Z = 5
    raw = np.arange(1, Z + 1, dtype=np.float32)[:, None, None] * np.ones((1, 1, 1), np.float32)
    # r = 0.5, keep_overhang: z_in = 0.5*o2 - 0.5*ct*o0, single tilt row o0 = 0 -> z_in = 0.5*o2
    d = o.deskew_data(raw, 30, 0.5, True, 1, cval=-1.0)
    want = [1, 1.5, 2, 2.5, 3, 3.5, 4, 4.5, 5] + [-1.0] * (d.shape[2] - 9)
    assert d[0, 0].tolist() == want


def test_ramps_check_the_flips():
    Z, Y, X = 20, 5, 6
    zz, yy, xx = np.meshgrid(np.arange(Z), np.arange(Y), np.arange(X), indexing="ij")
    dy = o.deskew_data(yy.astype(np.float32), 30, 0.5, True, 1, cval=-1)
    dx = o.deskew_data(xx.astype(np.float32), 30, 0.5, True, 1, cval=-1)
    for p in range(Y):
        vals = set(np.unique(dy[p])) - {-1.0}
        assert vals == {float(Y - 1 - p)}          # output axis 0 = flipped tilt axis
    for q in range(X):
        vals = set(np.unique(dx[:, q])) - {-1.0}
        assert vals == {float(X - 1 - q)}          # output axis 1 = flipped coverslip axis


def test_x_chunk_reverse_concat_identity():
    """scripts/measure_psf.py:218-249."""
    raw = synthetic_stack((40, 9, 16), seed=4)
    whole = o.deskew_data(raw, 30, 0.39, True, 3)
    parts = [o.deskew_data(c, 30, 0.39, True, 3) for c in np.split(raw, 4, axis=-1)]
    assert np.array_equal(np.concatenate(parts[::-1], axis=-2), whole)


def test_average_n_slices():
    d = np.arange(10 * 2 * 3, dtype=np.float32).reshape(10, 2, 3)
    assert np.array_equal(o.average_n_slices(d, 1), d)
    a = o.average_n_slices(d, 3)
    assert a.shape == (4, 2, 3) and a.dtype == np.float32
    assert np.array_equal(a[0], d[:3].mean(axis=0))
    assert np.array_equal(a[3], d[9])              # 10 % 3 = 1 -> last plane replicated twice


def test_affine_statements_agree():
    rng = np.random.default_rng(0)
    vol = rng.standard_normal((7, 12, 13)).astype(np.float32)
    th = np.deg2rad(5.0)
    M = np.array([[1.02, 0.03, 0.0, 0.5], [0.0, np.cos(th), -np.sin(th), 1.0], [0.01, np.sin(th), np.cos(th), -0.7],
                  [0, 0, 0, 1]])
    a = o.apply_affine_transform(vol, M, (8, 12, 14))
    b = o.apply_affine_transform_closed_form(vol, M, (8, 12, 14))
    c = c_oracle.apply_affine_transform(vol, M, (8, 12, 14), threads=2)
    assert np.array_equal(a == 0.0, b == 0.0) and np.array_equal(a == 0.0, c == 0.0)
    assert np.max(np.abs(a - b)) <= 1e-6 and np.max(np.abs(a - c)) <= 1e-6
    vol[3, 4, 5] = np.nan
    assert np.isfinite(o.apply_affine_transform(vol, M, (8, 12, 14))).all()
    assert np.isfinite(c_oracle.apply_affine_transform(vol, M, (8, 12, 14))).all()


def test_golden_vectors():
    """The committed scipy outputs (tests/golden/make_golden.py) are reproduced by all three statements."""
    data = np.load(GOLDEN / "deskew_small.npz")
    names = sorted({k.split("__")[0] for k in data.files if k.endswith("__params")})
    assert len(names) == 6
    for name in names:
        raw = data[f"{name}__raw"]
        ang, r, keep, n, cval = data[f"{name}__params"]
        want = data[f"{name}__out"]
        assert np.array_equal(o.deskew_data(raw, ang, r, bool(keep), int(n), cval=cval), want), name
        assert np.array_equal(o.deskew_data_closed_form(raw, ang, r, bool(keep), int(n), cval=cval), want), name
        assert np.array_equal(c_oracle.deskew_data(raw, ang, r, bool(keep), int(n), cval=cval), want), name
    vol = data["affine_general__vol"]
    for tag in ("general", "rot90"):
        M, want = data[f"affine_{tag}__matrix"], data[f"affine_{tag}__out"]
        assert np.array_equal(o.apply_affine_transform(vol, M, want.shape), want)
        assert np.max(np.abs(c_oracle.apply_affine_transform(vol, M, want.shape) - want)) <= 1e-6


def test_flatfield_oracle_matches_the_reference_golden():
    """tests/golden/flatfield.npz was produced by the UNMODIFIED reference method
    _LabelfreePreprocessor._flat_field_BF (shrimpy/preprocessing.py:385-404) -- a reference-pinned result."""
    from oracle import flatfield_oracle as ff

    data = np.load(GOLDEN / "flatfield.npz")
    for name in ("even_z", "odd_z", "tall"):
        vol, want = data[f"{name}__vol"], data[f"{name}__out"]
        got = ff.flat_field_BF(vol)
        assert got.shape == want.shape and got.dtype == np.float32
        # the reference test's own tolerance is atol=1e-2 (tests/test_preprocessing.py:162); we hold 1e-3
        assert np.max(np.abs(got - want)) <= 1e-3


def test_non_finite_voxels_are_outside_the_parity_contract():
    """Camera stacks are uint16, so every raw voxel is finite; for float32 stacks that are not, the three statements
    stop agreeing and the documented scope says so (oracle header).  scipy multiplies every one of the
    eight trilinear taps, zero weights included, so one NaN voxel poisons the outputs whose *zero-weight* y/x taps
    touch it; the closed form and the C restatement (and the CUDA kernels) interpolate along scan only."""
    raw = np.random.default_rng(0).normal(size=(20, 6, 8)).astype(np.float32)
    raw[9, 4, 1] = np.nan
    a = o.deskew_data(raw, 30.0, 0.39, True, 1)
    b = o.deskew_data_closed_form(raw, 30.0, 0.39, True, 1)
    c = c_oracle.deskew_data(raw, 30.0, 0.39, True, 1)
    assert np.array_equal(np.isnan(b), np.isnan(c)) and 0 < np.isnan(b).sum() < np.isnan(a).sum()
    assert not (np.isnan(b) & ~np.isnan(a)).any()                   # scipy's NaN set contains the closed form's
    finite = ~np.isnan(a)
    assert np.max(np.abs(a[finite] - b[finite])) <= 1e-6 * float(np.nanmax(b) - np.nanmin(b))
