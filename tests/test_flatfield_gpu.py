"""Flat-field (bright-field) on the GPU: median-over-Z pattern, scale field, stand-alone correction and the
version fused into the deskew kernel, against the reference-generated golden and the numpy oracle."""

from pathlib import Path

import numpy as np
import pytest

from helpers import assert_close_range, synthetic_stack

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def env():
    import torch

    from oracle import deskew_oracle, flatfield_oracle
    from shrimpy_b200 import flatfield

    return torch, flatfield, flatfield_oracle, deskew_oracle


def _bright_field(shape, seed):
    """uint16 stack with a smooth illumination pattern (so the correction is not the identity)."""
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    yy, xx = np.meshgrid(np.linspace(0.6, 1.4, Y), np.linspace(0.75, 1.25, X), indexing="ij")
    vol = rng.integers(2000, 9000, shape).astype(np.float64) * (yy * xx)[None]
    return np.round(vol).astype(np.uint16)


def test_reference_golden(env):
    """Outputs of the unmodified reference _flat_field_BF (tests/golden/make_golden.py); its own test uses atol=1e-2."""
    torch, ff, _, _ = env
    data = np.load(GOLDEN / "flatfield.npz")
    for name in ("even_z", "odd_z", "tall"):
        vol, want = data[f"{name}__vol"], data[f"{name}__out"]
        for t in (torch.from_numpy(vol).cuda(), torch.from_numpy(vol.astype(np.float32)).cuda()):
            got = ff.flat_field_BF(t).cpu().numpy()
            assert got.dtype == np.float32 and got.shape == want.shape
            assert np.max(np.abs(got - want)) <= 2e-3, name


@pytest.mark.parametrize("shape", [(8, 6, 10), (9, 5, 12), (33, 7, 64), (64, 3, 130), (600, 4, 192), (1201, 2, 70), (2500, 2, 33)])
@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_median_pattern_is_exact(env, shape, dtype):
    """The radix select is exact: the pattern equals numpy.median bit for bit (even and odd Z, ties included)."""
    torch, ff, ffo, _ = env
    vol = _bright_field(shape, seed=sum(shape))
    if dtype == np.float32:
        vol = (vol.astype(np.float32) - 4000.25) * np.float32(0.37)        # negative values and fractions
    vol[:, 0, 0] = vol[0, 0, 0]                                              # a column of ties
    got = ff.flat_field_pattern(torch.from_numpy(vol).cuda()).cpu().numpy()
    assert np.array_equal(got, ffo.flat_field_pattern(vol))


@pytest.mark.parametrize("shape", [(128, 3, 64), (600, 2, 128), (130, 2, 66), (64, 2, 31)])
def test_median_full_range_even_z(env, shape):
    """Keys spread over the whole uint16 range with an even Z: the two middle values usually sit in different high
    bytes, so the upper one comes from the minimum the last counting pass tracks (packed 16-bit DPX form in the paired
    path, 32-bit form in the scalar one); columns at the extremes (0, 65535) exercise the wrap-around limits."""
    torch, ff, ffo, _ = env
    rng = np.random.default_rng(sum(shape))
    vol = rng.integers(0, 65536, shape, dtype=np.uint16)
    vol[:, 0, 1] = 65535
    vol[: shape[0] // 2, 0, 2] = 65535                       # lower middle 65535 is impossible, upper middle is
    vol[shape[0] // 2:, 0, 2] = 0
    vol[: shape[0] // 2, 0, 3] = 65280                       # 0xff00 / 0xffff: both in the last high byte
    vol[shape[0] // 2:, 0, 3] = 65535
    vol[:, 0, 4] = np.where(np.arange(shape[0]) % 2 == 0, 255, 256).astype(np.uint16)   # neighbours across a high byte
    got = ff.flat_field_pattern(torch.from_numpy(vol).cuda()).cpu().numpy()
    assert np.array_equal(got, ffo.flat_field_pattern(vol))
    f32 = (vol.astype(np.float32) - 32768.0) * np.float32(1.7)
    got = ff.flat_field_pattern(torch.from_numpy(f32).cuda()).cpu().numpy()
    assert np.array_equal(got, ffo.flat_field_pattern(f32))


def test_scale_and_standalone_correction(env):
    torch, ff, ffo, _ = env
    vol = _bright_field((40, 12, 96), seed=5)
    t = torch.from_numpy(vol).cuda()
    pattern = ffo.flat_field_pattern(vol)
    scale = ff.flat_field_scale(t).cpu().numpy()
    assert np.allclose(scale, pattern.mean(dtype=np.float64) / pattern, rtol=3e-7)
    got = ff.flat_field_BF(t).cpu().numpy()
    assert_close_range(got, ffo.flat_field_BF(vol), 1e-6, "stand-alone flat-field")


@pytest.mark.parametrize("n,keep", [(1, True), (3, False), (2, True), (4, False), (5, True)])
@pytest.mark.parametrize("kernel", ["auto", "direct"])
def test_fused_flatfield_deskew_matches_two_steps(env, n, keep, kernel):
    """deskew(flat_field(raw)) as the reference runs it (preprocessing.py:320-327) == the fused kernel."""
    torch, ff, ffo, o = env
    raw = _bright_field((90, 14, 128), seed=n)
    want = o.deskew_data(ffo.flat_field_BF(raw), 30.0, 0.39, keep, n, cval=-5.0)
    got = ff.deskew_flat_field_zyx(torch.from_numpy(raw).cuda(), 30.0, 0.39, keep, n, cval=-5.0, kernel=kernel)
    got = got.cpu().numpy()
    assert_close_range(got, want, 2e-6, f"fused n={n} keep={keep} {kernel}")
    assert np.array_equal(got == -5.0, want == -5.0)


def test_fused_kernels_agree_and_float_input(env):
    torch, ff, _, _ = env
    raw = _bright_field((120, 10, 192), seed=9)
    t = torch.from_numpy(raw).cuda()
    a = ff.deskew_flat_field_zyx(t, 30.0, 0.39, True, 3, kernel="tma")
    b = ff.deskew_flat_field_zyx(t, 30.0, 0.39, True, 3, kernel="direct")
    c = ff.deskew_flat_field_zyx(t.to(torch.float32), 30.0, 0.39, True, 3, kernel="tma")
    assert torch.equal(a, b) and torch.equal(a, c)
    scale = ff.flat_field_scale(t)
    d = ff.deskew_flat_field_zyx(t, 30.0, 0.39, True, 3, scale=scale)
    assert torch.equal(a, d)
    with pytest.raises(ValueError):
        ff.deskew_flat_field_zyx(t, 30.0, 0.39, True, 3, scale=scale[:, :10])
