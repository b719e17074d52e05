"""GPU parity of the affine registration resample against the scipy oracle."""

from pathlib import Path

import numpy as np
import pytest

from helpers import CONTRACT_TOL, assert_close_range

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
AFFINE_TOL = 3e-6   # three nested float32 lerps vs scipy's float64 weights


@pytest.fixture(scope="module")
def env():
    import torch

    from oracle import c_oracle, deskew_oracle
    from shrimpy_b200 import register

    return torch, register, deskew_oracle, c_oracle


def _rot(a, b, c):
    a, b, c = np.deg2rad([a, b, c])
    Rz = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rx = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _run(env, vol, M, shape, **kw):
    torch, register, _, _ = env
    out = register.affine_transform_zyx(torch.from_numpy(vol).cuda(), M, shape, **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.fixture(params=["auto", "tile", "gather"])
def kernel_choice(request, monkeypatch):
    """Run a test through the default dispatch (planar kernel for block-diagonal matrices, else the tiled one),
    through the tiled kernel alone, and through the plain gather kernel."""
    if request.param == "auto":
        monkeypatch.delenv("SHRIMPY_AFFINE_KERNEL", raising=False)
    else:
        monkeypatch.setenv("SHRIMPY_AFFINE_KERNEL", request.param)
    return request.param


def test_golden_vectors(env, kernel_choice):
    data = np.load(GOLDEN / "deskew_small.npz")
    vol = data["affine_general__vol"]
    for tag in ("general", "rot90"):
        M, want = data[f"affine_{tag}__matrix"], data[f"affine_{tag}__out"]
        got = _run(env, vol, M, want.shape)
        rel = assert_close_range(got, want, AFFINE_TOL, tag)
        assert rel <= CONTRACT_TOL
        assert np.array_equal(got == 0.0, want == 0.0)


@pytest.mark.parametrize("shape_in,shape_out", [((20, 64, 80), (20, 64, 80)), ((11, 33, 47), (13, 50, 129)),
                                                  ((1, 5, 7), (1, 5, 7)), ((6, 40, 40), (0, 3, 3))])
def test_general_matrix_matches_scipy(env, shape_in, shape_out, kernel_choice):
    _, _, o, _ = env
    rng = np.random.default_rng(sum(shape_in))
    vol = rng.standard_normal(shape_in).astype(np.float32)
    M = np.eye(4)
    M[:3, :3] = _rot(2.0, 1.0, 3.0) @ np.diag([1.03, 0.97, 1.1])
    M[:3, 3] = [0.4, -1.2, 2.3]
    want = o.apply_affine_transform(vol, M, shape_out, cval=-2.0)
    got = _run(env, vol, M, shape_out, cval=-2.0)
    assert got.shape == want.shape
    if want.size:
        assert_close_range(got, want, AFFINE_TOL, "general")
        assert np.array_equal(got == -2.0, want == -2.0)


def test_identity_and_integer_shift_are_exact(env):
    rng = np.random.default_rng(5)
    vol = rng.standard_normal((9, 31, 37)).astype(np.float32)
    assert np.array_equal(_run(env, vol, np.eye(4), vol.shape), vol)
    M = np.eye(4)
    M[:3, 3] = [1, -2, 3]
    got = _run(env, vol, M, vol.shape, cval=9.0)
    want = np.full_like(vol, 9.0)
    want[:8, 2:, :34] = vol[1:, :29, 3:]
    assert np.array_equal(got, want)


def test_rot90_scale_onto_deskewed_grid(env, kernel_choice):
    """Mantis-like label-free -> fluorescence registration: in-plane 90 deg rotation x 1.288 + shift."""
    _, _, o, c = env
    rng = np.random.default_rng(6)
    vol = rng.standard_normal((12, 96, 128)).astype(np.float32)
    M = np.array([[1.0, 0, 0, 0.5], [0, 0, -1.288, 120.0], [0, 1.288, 0, -3.0], [0, 0, 0, 1]])
    want = o.apply_affine_transform(vol, M, (10, 90, 70))
    got = _run(env, vol, M, (10, 90, 70))
    assert_close_range(got, want, AFFINE_TOL, "rot90")
    assert np.array_equal(got == 0.0, want == 0.0)


def test_nan_to_num_and_numpy_front_end(env, kernel_choice):
    torch, register, o, _ = env
    rng = np.random.default_rng(7)
    vol = rng.standard_normal((5, 20, 24)).astype(np.float32)
    vol[2, 3, 4] = np.nan
    vol[1, 7, 9] = np.inf
    M = np.eye(4)
    M[:3, :3] = _rot(0.0, 0.0, 4.0)
    want = o.apply_affine_transform(vol, M, vol.shape)
    got = register.apply_affine_transform(vol, M, vol.shape)          # numpy in -> numpy out
    assert isinstance(got, np.ndarray) and np.isfinite(got).all()
    far = np.abs(want) < 1e30                                            # away from the FLT_MAX taps
    assert np.max(np.abs(got[far] - want[far])) <= 1e-5
    with pytest.raises(ValueError):
        register.apply_affine_transform(vol, np.ones((4, 4)), vol.shape)


def test_tile_and_gather_kernels_agree(env, monkeypatch):
    rng = np.random.default_rng(11)
    vol = rng.standard_normal((16, 150, 170)).astype(np.float32)
    for M in (np.array([[1.0, 0, 0, 0.5], [0, 0, -1.288, 160.0], [0, 1.288, 0, -3.0], [0, 0, 0, 1]]),
              np.array([[0.9, 0.05, 0.02, 0.3], [0.01, 1.1, -0.07, 2.0], [0.03, 0.06, 0.95, -1.5], [0, 0, 0, 1]]),
              np.diag([2.5, 0.4, 3.0, 1.0])):
        monkeypatch.delenv("SHRIMPY_AFFINE_KERNEL", raising=False)
        a = _run(env, vol, M, (14, 120, 200), cval=1.5)
        monkeypatch.setenv("SHRIMPY_AFFINE_KERNEL", "gather")
        b = _run(env, vol, M, (14, 120, 200), cval=1.5)
        # same geometry bit for bit; the tiled kernel truncates the lerp weights to 23 bits
        assert np.array_equal(a == 1.5, b == 1.5)
        assert_close_range(a, b, AFFINE_TOL, "tile vs gather")


@pytest.mark.parametrize("zrow", [(0.6, 0.0, 0.0, 0.3), (1.0, 0.0, 0.0, -2.0), (2.3, 0.0, 0.0, 0.7), (1.1, 0.02, -0.015, 1.5),
                                  (-0.9, 0.0, 0.0, 30.2)])
def test_z_separable_matrices(env, zrow, monkeypatch):
    """Input (y, x) independent of o0 (in-plane registration + z shift/scale/tilt): the kernel reuses the bilinear
    value of an input plane across consecutive o0; the scan index may advance by 0, 1, 2 or go backwards."""
    _, _, o, _ = env
    rng = np.random.default_rng(21)
    vol = rng.standard_normal((33, 70, 90)).astype(np.float32)
    th = np.deg2rad(7.0)
    M = np.array([list(zrow), [0.0, 1.05 * np.cos(th), -np.sin(th), 4.0], [0.0, np.sin(th), 0.95 * np.cos(th), -3.0],
                  [0, 0, 0, 1.0]])
    want = o.apply_affine_transform(vol, M, (40, 64, 100), cval=0.5)
    auto = _run(env, vol, M, (40, 64, 100), cval=0.5)            # planar kernel when the z row has no tilt
    assert_close_range(auto, want, AFFINE_TOL, f"auto {zrow}")
    assert np.array_equal(auto == 0.5, want == 0.5)
    monkeypatch.setenv("SHRIMPY_AFFINE_KERNEL", "tile")
    got = _run(env, vol, M, (40, 64, 100), cval=0.5)
    assert_close_range(got, want, AFFINE_TOL, f"zsep {zrow}")
    assert np.array_equal(got == 0.5, want == 0.5)
    monkeypatch.setenv("SHRIMPY_AFFINE_NO_ZSEP", "1")
    general = _run(env, vol, M, (40, 64, 100), cval=0.5)
    assert np.array_equal(got, general)          # same lerp sequence -> bit-identical to the general path


@pytest.mark.parametrize("angle", [0.0, 7.0, 45.0, 90.0, 93.0, 180.0, 270.0])
@pytest.mark.parametrize("zscale", [1.0, 0.55, 1.7])
def test_planar_kernel_in_plane_rotations(env, angle, zscale):
    """In-plane rotation x scale + z shift/scale (the mantis registration family): planar kernel, lanes along o2
    for small angles and along o1 (with the shared-memory output transpose) near 90/270 degrees."""
    _, _, o, _ = env
    rng = np.random.default_rng(int(angle) + 7)
    vol = rng.standard_normal((21, 150, 131)).astype(np.float32)
    vol = np.ascontiguousarray(np.pad(vol, ((0, 0), (0, 0), (0, 1))))          # X = 132, a multiple of 4
    th = np.deg2rad(angle)
    cy, cx = 75.0, 66.0
    R = 1.288 * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    M = np.eye(4)
    M[0, 0], M[0, 3] = zscale, 0.4
    M[1:3, 1:3] = R
    M[1:3, 3] = np.array([cy, cx]) - R @ np.array([50.0, 70.0])
    shape = (25, 100, 140)
    want = o.apply_affine_transform(vol, M, shape, cval=-7.0)
    got = _run(env, vol, M, shape, cval=-7.0)
    assert_close_range(got, want, AFFINE_TOL, f"planar {angle} {zscale}")
    assert np.array_equal(got == -7.0, want == -7.0)


def test_planar_wide_tiles_identity_and_shift(env):
    """Wide outputs select the 64-lane tiles of the planar kernel; identity must reproduce the input exactly."""
    rng = np.random.default_rng(31)
    vol = rng.standard_normal((20, 96, 512)).astype(np.float32)
    assert np.array_equal(_run(env, vol, np.eye(4), vol.shape), vol)
    M = np.eye(4)
    M[:3, 3] = [2, -3, 5]
    got = _run(env, vol, M, vol.shape, cval=9.0)
    want = np.full_like(vol, 9.0)
    want[:18, 3:, :507] = vol[2:, :93, 5:]
    assert np.array_equal(got, want)


STREAM_CASES = {
    # name: (z row (scale, shift), in-plane angle, in-plane scale, input shape, output shape)
    "long_march_up": ((0.13, 0.4), 7.0, 1.05, (40, 72, 96), (300, 70, 100)),
    "long_march_down": ((-0.21, 38.7), 7.0, 0.95, (40, 72, 96), (200, 64, 90)),
    "rot90_long": ((0.3, 0.2), 90.0, 1.288, (50, 96, 128), (170, 90, 70)),
    "rot270_skip_planes": ((2.6, -1.0), 270.0, 0.8, (60, 96, 128), (30, 100, 110)),
    "z_constant": ((0.0, 3.25), 3.0, 1.0, (8, 64, 64), (5, 64, 64)),
}


@pytest.mark.parametrize("case", sorted(STREAM_CASES))
@pytest.mark.parametrize("cfg", ["", "4,2,8", "2,4,4", "2,2,8", "4,1,4", "2,1,8", "1,8,4", "1,4,8"])
def test_stream_kernel_marches(env, case, cfg, monkeypatch):
    """Block-diagonal matrices through the z-streaming kernel (forced): marches longer than one launch (> 128 output
    planes), planes visited downwards, skipped planes (|z scale| > 2), a constant z, every tile shape and ring depth."""
    _, _, o, _ = env
    (zs, zt), angle, scale, shape_in, shape_out = STREAM_CASES[case]
    th = np.deg2rad(angle)
    R = scale * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    M = np.eye(4)
    M[0, 0], M[0, 3] = zs, zt
    M[1:3, 1:3] = R
    M[1:3, 3] = np.array([shape_in[1] / 2, shape_in[2] / 2]) - R @ np.array([shape_out[1] / 2, shape_out[2] / 2])
    rng = np.random.default_rng(len(case) + 3)
    vol = rng.standard_normal(shape_in).astype(np.float32)
    want = o.apply_affine_transform(vol, M, shape_out, cval=2.5)
    monkeypatch.setenv("SHRIMPY_AFFINE_KERNEL", "s!")
    if cfg:
        monkeypatch.setenv("SHRIMPY_STREAM_CFG", cfg)
    try:
        got = _run(env, vol, M, shape_out, cval=2.5)
    except Exception as exc:   # a forced tile shape exists for one lane orientation only
        if cfg and "not eligible" in str(exc):
            pytest.skip(f"tile {cfg} does not apply to {case}")
        raise
    assert_close_range(got, want, AFFINE_TOL, f"stream {case} {cfg}")
    assert np.array_equal(got == 2.5, want == 2.5)


TILT_CASES = {
    "small_rotations": ([[1.0284, -0.0508, 0.0192, 0.4], [0.0545, 0.968, -0.0384, -1.2], [-0.0161, 0.0347, 1.0992, 2.3]], (40, 96, 132), (44, 90, 150)),
    "downwards": ([[-0.93, 0.03, -0.02, 37.6], [0.02, 1.04, 0.05, -2.0], [0.01, -0.04, 0.91, 6.0]], (40, 96, 132), (44, 100, 140)),
    "z_stretch": ([[0.47, 0.011, 0.017, 3.2], [-0.03, 0.9, 0.02, 4.0], [0.02, 0.01, 1.2, -7.5]], (30, 80, 200), (70, 97, 170)),
    "z_squeeze": ([[2.3, -0.02, 0.03, -4.0], [0.1, 1.0, 0.0, 1.0], [0.0, 0.1, 1.0, -3.0]], (90, 64, 96), (45, 70, 101)),
    "steep_tilt": ([[1.0, 0.21, -0.13, 2.0], [-0.2, 0.98, 0.0, 9.0], [0.12, 0.0, 0.99, 1.0]], (48, 120, 136), (50, 110, 130)),
    "two_launches": ([[0.25, 0.004, -0.006, 1.1], [0.003, 1.0, 0.02, 0.2], [-0.004, -0.02, 1.0, 2.6]], (40, 40, 64), (150, 44, 70)),
    "rot90_tilt": ([[1.0, 0.03, -0.02, 1.5], [0.02, 0.01, -1.288, 170.0], [-0.015, 1.288, 0.02, -3.0]], (30, 140, 136), (28, 100, 100)),
    "rot270_tilt_down": ([[-0.8, -0.02, 0.04, 25.0], [0.03, -0.02, 0.9, 3.0], [0.01, -1.1, 0.03, 130.0]], (30, 120, 132), (33, 110, 125)),
    "no_z_motion": ([[0.0, 0.05, 0.03, 3.3], [0.3, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0]], (12, 64, 64), (20, 60, 66)),
}


@pytest.mark.parametrize("case", sorted(TILT_CASES))
@pytest.mark.parametrize("cfg", ["", "2,2,0,0", "4,1,0,32", "2,4,0,0", "4,2,0,16", "2,1,0,64", "1,4,0,0", "1,8,0,32"])
def test_tilt_kernel_general_matrices(env, case, cfg, monkeypatch):
    """General 3-D matrices through the marching tilt kernel (forced: an ineligible case is an error, not a silent
    fallback): planes visited up- and downwards, z stretch / squeeze, steep tilt, marches split over several launches,
    every tile shape; geometry bit-exact, values within the float32 lerp tolerance of scipy."""
    _, _, o, _ = env
    rows, shape_in, shape_out = TILT_CASES[case]
    M = np.array(rows + [[0, 0, 0, 1.0]], dtype=np.float64)
    rng = np.random.default_rng(len(case))
    vol = rng.standard_normal(shape_in).astype(np.float32)
    want = o.apply_affine_transform(vol, M, shape_out, cval=-3.0)
    monkeypatch.setenv("SHRIMPY_AFFINE_KERNEL", "m!")
    if cfg:
        monkeypatch.setenv("SHRIMPY_TILT_CFG", cfg)
    try:
        got = _run(env, vol, M, shape_out, cval=-3.0)
    except Exception as exc:   # a forced tile shape may not fit this matrix; the automatic choice must
        if cfg and "not eligible" in str(exc):
            pytest.skip(f"tile {cfg} does not fit {case}")
        raise
    assert_close_range(got, want, AFFINE_TOL, f"tilt {case} {cfg}")
    assert np.array_equal(got == -3.0, want == -3.0)


def test_tilt_kernel_nan_to_num(env, monkeypatch):
    torch, register, o, _ = env
    rng = np.random.default_rng(17)
    vol = rng.standard_normal((16, 48, 64)).astype(np.float32)
    vol[5, 20, 30] = np.nan
    vol[9, 11, 40] = -np.inf
    M = np.array(TILT_CASES["small_rotations"][0] + [[0, 0, 0, 1.0]])
    want = o.apply_affine_transform(vol, M, vol.shape)
    monkeypatch.setenv("SHRIMPY_AFFINE_KERNEL", "m!")
    got = _run(env, vol, M, vol.shape)
    assert np.isfinite(got).all()
    far = np.abs(want) < 1e30
    assert np.max(np.abs(got[far] - want[far])) <= 1e-5
    assert np.array_equal(got == 0.0, want == 0.0)


@pytest.mark.parametrize("which", ["in_plane_90", "small_rotations"])
def test_odd_row_length_goes_through_padded_rows(env, which, monkeypatch):
    """X not a multiple of 4 (e.g. a deskewed (.., 1279) volume as the INPUT of a registration): the Python layer pads
    the rows once and the plane-streaming kernels read through the strided tensor map (TMA zero-fills beyond the logical
    X).  Same voxels as scipy, same padded set, and the same as the dense-only kernels within the lerp tolerance."""
    torch, register, _, c = env
    rng = np.random.default_rng(41)
    vol = rng.standard_normal((40, 300, 353)).astype(np.float32)          # 4.2 M voxels, X = 353
    if which == "in_plane_90":
        M = np.array([[0.9, 0, 0, 1.5], [0, 0, -1.1, 298.0], [0, 1.1, 0, 2.0], [0, 0, 0, 1.0]])
        shape = (38, 300, 280)
    else:
        M = np.array(TILT_CASES["small_rotations"][0] + [[0, 0, 0, 1.0]])
        shape = (40, 300, 353)
    want = c.apply_affine_transform(vol, M, shape, cval=-4.0)
    monkeypatch.setenv("SHRIMPY_AFFINE_KERNEL", "s!" if which == "in_plane_90" else "m!")   # a fallback would be an error
    got = _run(env, vol, M, shape, cval=-4.0)
    assert_close_range(got, want, AFFINE_TOL, f"padded rows {which}")
    assert np.array_equal(got == -4.0, want == -4.0)
    monkeypatch.delenv("SHRIMPY_AFFINE_KERNEL")
    monkeypatch.setattr(register, "_PAD_MIN_VOXELS", 1 << 62)          # dense-only kernels on the unpadded array
    dense = _run(env, vol, M, shape, cval=-4.0)
    assert np.array_equal(dense == -4.0, got == -4.0)
    assert_close_range(dense, got, AFFINE_TOL, "dense vs padded")


def test_medium_volume_vs_c_oracle(env):
    _, _, _, c = env
    rng = np.random.default_rng(8)
    vol = rng.standard_normal((24, 200, 256)).astype(np.float32)
    M = np.eye(4)
    M[:3, :3] = _rot(2.0, 1.0, 3.0) @ np.diag([1.03, 0.97, 1.1])
    M[:3, 3] = [0.4, -1.2, 2.3]
    want = c.apply_affine_transform(vol, M, vol.shape)
    got = _run(env, vol, M, vol.shape)
    assert_close_range(got, want, AFFINE_TOL, "medium")
    assert np.array_equal(got == 0.0, want == 0.0)
