"""Parity at BASELINE.json's FULL sizes, through properties that do not need the CPU oracle to finish the whole volume:

* a seeded random SAMPLE of output voxels is recomputed on the GPU with plain float64 torch ops that restate the
  closed form of SURVEY.md appendix C / oracle/deskew_oracle.py (scipy's coordinate order, strict outside rule);
  the padded-voxel set must match bit for bit and the values within the float32 lerp tolerance;
* an X slab of the result equals the C oracle run on that slab alone (the deskew never mixes x columns:
  scripts/measure_psf.py:218-249), and X chunks concatenated in reverse equal the whole, exactly;
* the TMA and the direct kernel agree bit for bit; windows reassemble bit for bit;
* the affine resample reproduces the input under the identity and under integer shifts, exactly.

Sizes: config 2 (600,300,2048) uint16 n=3; config 3 (107,2048,2048) float32; config 5 (4000,300,2048) uint16,
keep_overhang=True, n=1 -> (300,2048,10517) float32 = 25.9 GB on one GPU.
"""

import numpy as np
import pytest

from helpers import TIGHT_TOL

pytestmark = pytest.mark.gpu
AFFINE_TOL = 3e-6


@pytest.fixture(scope="module")
def env():
    import torch

    import shrimpy_b200 as sb
    from oracle import c_oracle
    from shrimpy_b200 import register

    assert torch.cuda.is_available()
    return torch, sb, register, c_oracle


def _random_u16(torch, shape, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(100, 60000, shape, dtype=torch.int32, device="cuda", generator=gen).to(torch.uint16)


def _sample_deskew(torch, raw, g, cval, count, seed):
    """Expected values of `count` random output voxels (float64 torch restatement of the closed form)."""
    Z, Y, X = g.raw_shape
    P, _, Xp = g.out_shape
    n = g.n_avg
    gen = torch.Generator(device="cuda").manual_seed(seed)
    p = torch.randint(0, P, (count,), device="cuda", generator=gen)
    o1 = torch.randint(0, X, (count,), device="cuda", generator=gen)
    o2 = torch.randint(0, Xp, (count,), device="cuda", generator=gen)
    # a quarter of the samples sit on the first / last columns, where the padding starts
    edge = count // 4
    o2[:edge] = torch.randint(0, min(Xp, 300), (edge,), device="cuda", generator=gen)
    o2[edge:2 * edge] = Xp - 1 - torch.randint(0, min(Xp, 300), (edge,), device="cuda", generator=gen)
    acc = torch.zeros(count, dtype=torch.float64, device="cuda")
    n_in = torch.zeros(count, dtype=torch.int64, device="cuda")
    rawf = raw.reshape(-1).view(torch.int16)          # torch has no CUDA gather for uint16: fetch the bits, undo the sign

    def taps(index):
        return (rawf[index].to(torch.int32) & 0xFFFF).to(torch.float64)

    for k in range(n):
        o0 = torch.clamp(n * p + k, max=Y - 1)
        z = (g.shift + o0.to(torch.float64) * g.m00) + o2.to(torch.float64) * g.m02     # products and sums rounded separately
        inside = (z >= 0.0) & (z <= Z - 1.0)
        z0 = torch.clamp(torch.floor(z), 0, Z - 1).to(torch.int64)
        z1 = torch.clamp(z0 + 1, max=Z - 1)
        w = z - z0.to(torch.float64)
        col = (Y - 1 - o0) * X + (X - 1 - o1)
        a = taps(z0 * (Y * X) + col)
        b = taps(z1 * (Y * X) + col)
        v = torch.where(inside, a + w * (b - a), torch.full_like(a, cval))
        acc += v
        n_in += inside.to(torch.int64)
    return (p, o1, o2), acc / n, n_in


def _check_samples(torch, out, idx, want, n_in, cval, tol, what):
    got = out[idx[0], idx[1], idx[2]].to(torch.float64)
    span = float(want.max() - want.min()) or 1.0
    err = float((got - want).abs().max())
    assert err <= tol * span, f"{what}: max|err| {err:.3e} > {tol:.1e} x range {span:.3e}"
    # voxels whose every contributing row is outside are exactly cval, and only those (values stay >= 100 inside)
    assert torch.equal(got == cval, n_in == 0), f"{what}: padded-voxel set differs"


def test_config2_full_size(env):
    torch, sb, _, c_oracle = env
    shape, ang, r, keep, n = (600, 300, 2048), 30.0, 0.39, False, 3
    g = sb.deskew_geometry(shape, ang, r, keep, n)
    assert g.out_shape == (100, 2048, 1279)
    raw = _random_u16(torch, shape, seed=2)
    out = sb.deskew_zyx(raw, ang, r, keep, n, cval=-1.0)
    assert tuple(out.shape) == g.out_shape and out.dtype == torch.float32

    idx, want, n_in = _sample_deskew(torch, raw, g, -1.0, 400_000, seed=11)
    _check_samples(torch, out, idx, want, n_in, -1.0, TIGHT_TOL, "config 2 samples")

    # an X slab against the C oracle (bit-exact geometry, values within tolerance); output axis 1 is raw x reversed
    xs = slice(1000, 1064)
    slab = c_oracle.deskew_data(raw[:, :, xs].cpu().numpy(), ang, r, keep, n, cval=-1.0)
    mine = out[:, 2048 - 1064:2048 - 1000, :].cpu().numpy()
    assert np.array_equal(mine == -1.0, slab == -1.0)
    assert np.max(np.abs(mine - slab)) <= TIGHT_TOL * float(slab.max() - slab.min())

    # kernels agree bit for bit; X chunks in reverse order reproduce the whole (scripts/measure_psf.py:248-249)
    assert torch.equal(out, sb.deskew_zyx(raw, ang, r, keep, n, cval=-1.0, kernel="direct"))
    chunks = [sb.deskew_zyx(raw[:, :, i:i + 512], ang, r, keep, n, cval=-1.0) for i in range(0, 2048, 512)]
    assert torch.equal(out, torch.cat(chunks[::-1], dim=1))

    # float32 input (what shrimpy/preprocessing.py:316 hands over) gives the same voxels
    assert torch.equal(out, sb.deskew_zyx(raw.to(torch.float32), ang, r, keep, n, cval=-1.0))


def test_config5_full_size_single_gpu(env):
    torch, sb, _, _ = env
    shape, ang, r, keep, n = (4000, 300, 2048), 30.0, 0.39, True, 1
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * 2**30:
        pytest.skip("needs 40 GB of free device memory")
    g = sb.deskew_geometry(shape, ang, r, keep, n)
    assert g.out_shape == (300, 2048, 10517)
    raw = _random_u16(torch, shape, seed=5)
    out = sb.deskew_zyx(raw, ang, r, keep, n, cval=7.0)
    idx, want, n_in = _sample_deskew(torch, raw, g, 7.0, 600_000, seed=55)
    _check_samples(torch, out, idx, want, n_in, 7.0, TIGHT_TOL, "config 5 samples")
    # the overhang (keep_overhang=True) really is padded: first and last columns of the far tilt rows
    assert float(out[0, 0, -1]) == 7.0 or float(out[-1, 0, 0]) == 7.0
    # a column window of the output recomputed from only the scan slices it needs equals the whole, bit for bit
    c0, c1 = 5000, 5600
    (y0, y1), (z0, z1) = sb.window_needs(g, 0, 300, c0, c1 - c0)
    assert (y0, y1) == (0, 300) and 0 < z0 < z1 < 4000 and z1 - z0 < 400      # ~ r * (600 + cos(theta) * 299) slices
    win = sb.deskew_window(raw[z0:z1], g, p_begin=0, p_count=300, c_begin=c0, c_count=c1 - c0, y_origin=0, z_origin=z0,
                           cval=7.0)
    assert torch.equal(win, out[:, :, c0:c1])
    del out, raw
    torch.cuda.empty_cache()


def _sample_affine(torch, vol, M, oshape, cval, count, seed):
    iz, iy, ix = vol.shape
    gen = torch.Generator(device="cuda").manual_seed(seed)
    o = [torch.randint(0, s, (count,), device="cuda", generator=gen) for s in oshape]
    c = []
    for a in range(3):
        t = (M[a][3] + o[0].to(torch.float64) * M[a][0]) + o[1].to(torch.float64) * M[a][1]
        c.append(t + o[2].to(torch.float64) * M[a][2])
    dims = (iz, iy, ix)
    inside = torch.ones(count, dtype=torch.bool, device="cuda")
    f, w = [], []
    for a in range(3):
        inside &= (c[a] >= 0.0) & (c[a] <= dims[a] - 1.0)
        fa = torch.clamp(torch.floor(c[a]), 0, dims[a] - 1).to(torch.int64)
        f.append(fa)
        w.append(c[a] - fa.to(torch.float64))
    g = [torch.clamp(f[a] + 1, max=dims[a] - 1) for a in range(3)]
    v = vol.reshape(-1).to(torch.float64)

    def tap(z, y, x):
        return v[(z * iy + y) * ix + x]

    def lerp(a, b, t):
        return a + t * (b - a)

    lo = lerp(lerp(tap(f[0], f[1], f[2]), tap(f[0], f[1], g[2]), w[2]), lerp(tap(f[0], g[1], f[2]), tap(f[0], g[1], g[2]), w[2]), w[1])
    hi = lerp(lerp(tap(g[0], f[1], f[2]), tap(g[0], f[1], g[2]), w[2]), lerp(tap(g[0], g[1], f[2]), tap(g[0], g[1], g[2]), w[2]), w[1])
    want = torch.where(inside, lerp(lo, hi, w[0]), torch.full_like(lo, cval))
    return o, want, inside


def _rot(a, b, c):
    a, b, c = np.deg2rad([a, b, c])
    Rz = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rx = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def test_config3_full_size(env):
    torch, _, register, _ = env
    shape = (107, 2048, 2048)
    gen = torch.Generator(device="cuda").manual_seed(3)
    vol = torch.randn(shape, device="cuda", generator=gen)
    Mg = np.eye(4)
    Mg[:3, :3] = _rot(2.0, 1.0, 3.0) @ np.diag([1.03, 0.97, 1.1])
    Mg[:3, 3] = [0.4, -1.2, 2.3]
    M90 = np.array([[1.0, 0, 0, 3.5], [0, 0, -1.288, 2040.0], [0, 1.288, 0, -20.0], [0, 0, 0, 1]])
    M90t = M90.copy()
    M90t[0, 1:3] = [0.02, -0.015]
    M90t[1, 0], M90t[2, 0] = 0.03, -0.02
    Mz = np.diag([0.75, 1.0, 1.0, 1.0])
    Mz[0, 3] = 0.3
    for name, M, oshape in (("general", Mg, shape), ("rot90", M90, (100, 2048, 1279)), ("rot90+tilt", M90t, (100, 2048, 1279)),
                            ("z scale", Mz, (140, 2048, 2048))):
        out = register.affine_transform_zyx(vol, M, oshape, cval=-9.0)
        o, want, inside = _sample_affine(torch, vol, M, oshape, -9.0, 400_000, seed=len(name))
        got = out[o[0], o[1], o[2]].to(torch.float64)
        span = float(want.max() - want.min())
        err = float((got - want).abs().max())
        assert err <= AFFINE_TOL * span, f"{name}: max|err| {err:.3e} > {AFFINE_TOL:.1e} x range {span:.3e}"
        assert torch.equal(got == -9.0, ~inside), f"{name}: padded-voxel set differs"
        del out

    # the reverse direction: a deskewed-grid volume (X = 1279, not a multiple of 4) onto the label-free grid
    back = torch.randn((100, 2048, 1279), device="cuda", generator=gen)
    Minv = np.linalg.inv(M90t)
    out = register.affine_transform_zyx(back, Minv, (107, 2048, 2048), cval=-9.0)
    o, want, inside = _sample_affine(torch, back, Minv, (107, 2048, 2048), -9.0, 400_000, seed=77)
    got = out[o[0], o[1], o[2]].to(torch.float64)
    err = float((got - want).abs().max())
    assert err <= AFFINE_TOL * float(want.max() - want.min()), f"odd X: max|err| {err:.3e}"
    assert torch.equal(got == -9.0, ~inside), "odd X: padded-voxel set differs"
    del out, back

    # identity and integer shifts are exact copies
    assert torch.equal(register.affine_transform_zyx(vol, np.eye(4), shape), vol)
    Ms = np.eye(4)
    Ms[:3, 3] = [2, -3, 5]
    got = register.affine_transform_zyx(vol, Ms, shape, cval=9.0)
    assert torch.equal(got[:105, 3:, :2043], vol[2:, :2045, 5:])
    assert bool((got[105:] == 9.0).all()) and bool((got[:, :3] == 9.0).all()) and bool((got[:, :, 2043:] == 9.0).all())
