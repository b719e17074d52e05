"""Tracking reductions over a deskewed volume against goldens produced by the UNMODIFIED reference helpers
(shrimpy/dynatrack/tracking.py:572-649, see tests/golden/make_golden.py)."""

import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


def _volume(seed, shape):
    rng = np.random.default_rng(seed)
    vol = rng.gamma(2.0, 300.0, size=shape).astype(np.float32)
    vol[shape[0] // 2, shape[1] // 3, shape[2] // 4] += 50000.0
    return vol


def test_percentile_and_center_of_mass_match_the_reference_golden():
    import torch

    from shrimpy_b200 import reductions as red

    for case in json.loads((GOLDEN / "reductions.json").read_text()):
        vol = _volume(case["seed"], tuple(case["shape"]))
        t = torch.from_numpy(vol).cuda()
        lo, hi = red.value_range(t)
        assert lo == float(vol.min()) and hi == float(vol.max())
        width = (hi - lo) / 256
        for p, want in case["percentiles"].items():
            got = red.percentile(t, float(p))
            # same bin as the reference (its float32 histogram and ours agree exactly at these sizes)
            assert abs(got - want) <= 1e-3 * width + 1e-6 * abs(want), (p, got, want)
        for bg, want in case["com"].items():
            got = red.intensity_center_of_mass(t, background=float(bg)).cpu().numpy()
            assert got.dtype == np.float32
            # the reference sums in float32; allow its rounding (centroids are O(10-100))
            assert np.allclose(got, np.array(want), atol=2e-3), (bg, got, want)


def test_constant_and_shifted_volumes():
    import torch

    from shrimpy_b200 import reductions as red

    const = torch.full((5, 6, 7), 3.25, device="cuda")
    assert red.percentile(const, 50.0) == 3.25                      # vmax <= vmin -> vmin (tracking.py:585-586)
    assert torch.equal(red.intensity_center_of_mass(const, background=10.0).cpu(), torch.tensor([2.0, 2.5, 3.0]))
    one = torch.zeros((9, 10, 11), device="cuda")
    one[4, 7, 2] = 5.0
    assert torch.equal(red.intensity_center_of_mass(one).cpu(), torch.tensor([4.0, 7.0, 2.0]))
    neg = -torch.rand((4, 5, 6), device="cuda") - 1.0                # negative "mass" never pulls the centroid
    assert torch.equal(red.intensity_center_of_mass(neg).cpu(), torch.tensor([1.5, 2.0, 2.5]))
    with pytest.raises(RuntimeError):
        red.percentile(torch.zeros(4), 50.0)


def test_on_a_deskewed_volume_end_to_end():
    """deskew -> background percentile -> centre of mass, the chain the tracker runs, vs numpy on the same output."""
    import torch

    import shrimpy_b200 as sb
    from helpers import synthetic_stack
    from shrimpy_b200 import reductions as red

    raw = synthetic_stack((120, 30, 128), seed=2)
    raw[50:60, 10:14, 40:60] += 5000
    vol = sb.deskew_zyx(torch.from_numpy(raw).cuda(), 30.0, 0.39, False, 3)
    host = vol.cpu().numpy().astype(np.float64)
    bg = red.percentile(vol, 50.0)
    w = np.clip(host - bg, 0, None)
    want = [(w.sum(axis=tuple(a for a in range(3) if a != ax)) * np.arange(w.shape[ax])).sum() / w.sum() for ax in range(3)]
    got = red.intensity_center_of_mass(vol, background=bg).cpu().numpy()
    assert np.allclose(got, want, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("n,keep,dtype", [(3, False, "u16"), (1, True, "u16"), (2, True, "f32"), (5, False, "u16")])
def test_value_range_fused_into_the_deskew(n, keep, dtype):
    """deskew_zyx(..., value_range=t) fills (min, max) of the result inside the deskew kernel: equal to torch.min /
    torch.max of the volume (what tracking.py:583-584 computes next), for the TMA and the direct kernel (n = 5),
    with padding included; the percentile that follows is the same with and without the saved pass."""
    import torch

    import shrimpy_b200 as sb
    from shrimpy_b200 import flatfield, reductions

    gen = torch.Generator(device="cuda").manual_seed(n)
    raw = torch.randint(300, 50000, (150, 23, 192), dtype=torch.int32, device="cuda", generator=gen)
    raw = raw.to(torch.uint16) if dtype == "u16" else raw.to(torch.float32) * 0.5 - 7000.0
    rng = torch.empty(2, dtype=torch.float32, device="cuda")
    out = sb.deskew_zyx(raw, 30.0, 0.39, keep, n, cval=-5.0, value_range=rng)
    assert torch.equal(out, sb.deskew_zyx(raw, 30.0, 0.39, keep, n, cval=-5.0))
    assert rng.tolist() == [float(out.min()), float(out.max())]
    assert reductions.percentile(out, 50.0, value_range=rng) == reductions.percentile(out, 50.0)
    if dtype == "u16" and n <= 4:
        scale = flatfield.flat_field_scale(raw)
        fused = sb.deskew_zyx(raw, 30.0, 0.39, keep, n, cval=-5.0, value_range=rng, scale=scale)
        assert torch.equal(fused, flatfield.deskew_flat_field_zyx(raw, 30.0, 0.39, keep, n, cval=-5.0, scale=scale))
        assert rng.tolist() == [float(fused.min()), float(fused.max())]


@pytest.mark.parametrize("shape", [(100, 64, 1279), (37, 50, 128), (1, 3, 5), (9, 1, 4)])
def test_max_projection_equals_the_torch_expression(shape):
    """tracking.py:1447-1455: ``(img - background).clamp_min(0).amax(dim=0)``; vector path (Y*X % 4 == 0) and scalar."""
    import torch

    from shrimpy_b200 import reductions as red

    vol = torch.from_numpy(_volume(11, shape)).cuda()
    for bg in (0.0, 412.5, 1e9):
        want = (vol - bg).clamp_min(0).amax(dim=0)
        got = red.max_projection(vol, bg)
        assert got.shape == want.shape and got.dtype == torch.float32
        assert torch.equal(got, want), (shape, bg)
    view = vol[:, :, 1:]                                             # a non-contiguous view is made contiguous first
    assert torch.equal(red.max_projection(view, 100.0), (view - 100.0).clamp_min(0).amax(dim=0))
    with pytest.raises(ValueError):
        red.max_projection(vol[0])


def test_histogram_bins_equal_the_ieee_division_form():
    """The kernel bins with a multiply and falls back to the division only near a bin edge; the counts must equal
    trunc((x - min) / (max - min) * 256) in float32 for every voxel -- including data that sits exactly on edges."""
    import torch

    from shrimpy_b200 import _cabi

    rng = np.random.default_rng(5)
    cases = {
        "integers_on_edges": rng.integers(0, 513, 300001).astype(np.float32),            # range 512: every value on an edge
        "uniform": rng.uniform(-3.0, 1000.0, 400003).astype(np.float32),
        "thirds": (rng.integers(0, 769, 200000) / np.float32(3.0)).astype(np.float32),  # edges of a range that is not 2^k
        "gamma": rng.gamma(2.0, 300.0, 500000).astype(np.float32),
        "large": rng.normal(300.0, 40.0, 12_000_001).astype(np.float32),                 # several pipelined batches per thread
    }
    for name, x in cases.items():
        lo, hi = np.float32(x.min()), np.float32(x.max())
        want = np.minimum(((x - lo) / (hi - lo) * np.float32(256.0)).astype(np.int64), 255)
        want = np.bincount(want, minlength=256)
        t = torch.from_numpy(x).cuda()
        for view in (t, t[1:]):                                                          # aligned and unaligned start
            hist = torch.empty(256, dtype=torch.int64, device="cuda")
            _cabi.check(_cabi.lib().shrimpy_hist256_device(view.data_ptr(), view.numel(), float(lo), float(hi),
                                                           hist.data_ptr(), torch.cuda.current_stream().cuda_stream))
            ref = want if view is t else np.bincount(
                np.minimum(((x[1:] - lo) / (hi - lo) * np.float32(256.0)).astype(np.int64), 255), minlength=256)
            assert np.array_equal(hist.cpu().numpy(), ref), name


@pytest.mark.parametrize("shape", [(5, 7, 1279), (3, 4, 2101), (2, 3, 64), (4, 5, 67)])
def test_center_of_mass_vector_path_against_float64(shape):
    """Rows whose 16-byte phase changes from row to row (X % 4 != 0), rows longer than one chunk of vectors (X > 1536),
    the shortest vectorised row (X = 64), and an unaligned base pointer (scalar kernel): all against a float64 sum."""
    import torch

    from shrimpy_b200 import reductions as red

    vol = _volume(21, shape)
    for bg in (0.0, 350.0):
        w = np.clip(vol.astype(np.float64) - bg, 0, None)
        want = [(w.sum(axis=tuple(a for a in range(3) if a != ax)) * np.arange(shape[ax])).sum() / w.sum() for ax in range(3)]
        t = torch.from_numpy(vol).cuda()
        got = red.intensity_center_of_mass(t, background=bg).cpu().numpy()
        assert np.allclose(got, want, rtol=2e-6, atol=1e-4), (shape, bg, got, want)
        flat = torch.empty(vol.size + 1, dtype=torch.float32, device="cuda")
        view = flat[1:].view(shape)                                   # 4-byte aligned only
        view.copy_(t)
        got = red.intensity_center_of_mass(view, background=bg).cpu().numpy()
        assert np.allclose(got, want, rtol=2e-6, atol=1e-4), (shape, bg, "unaligned")
