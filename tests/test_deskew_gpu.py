"""GPU parity: CUDA deskew (through the C-ABI) against the CPU oracle on identical seeded inputs."""

import numpy as np
import pytest

from helpers import CONTRACT_TOL, TIGHT_TOL, assert_close_range, synthetic_stack

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch

    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def sb():
    import shrimpy_b200

    return shrimpy_b200


@pytest.fixture(scope="module")
def oracle():
    from oracle import c_oracle, deskew_oracle

    return deskew_oracle, c_oracle


def _gpu(torch, sb, raw, *args, kernel="auto", **kw):
    t = torch.from_numpy(raw).cuda()
    out = sb.deskew_zyx(t, *args, kernel=kernel, **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy()


SMALL = [
    # shape, angle, ratio, keep_overhang, n
    ((23, 10, 7), 30.0, 0.39, True, 1),
    ((23, 10, 7), 30.0, 0.39, False, 3),
    ((40, 11, 5), 36.0, 0.651, True, 2),
    ((40, 11, 5), 36.0, 0.651, False, 3),
    ((64, 9, 64), 30.0, 0.39, False, 3),
    ((50, 12, 72), 30.0, 0.374, True, 4),
    ((31, 7, 130), 45.0, 0.77, False, 2),
    ((17, 5, 16), 12.5, 1.3, True, 1),
    ((2, 3, 4), 36.0, 0.386, True, 1),
    ((1, 1, 1), 30.0, 0.39, True, 1),
    ((30, 13, 24), 30.0, 0.39, False, 5),   # n > 4: direct kernel only
]


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
@pytest.mark.parametrize("kernel", ["direct", "auto"])
@pytest.mark.parametrize("case", SMALL, ids=lambda c: "x".join(map(str, c[0])) + f"-{c[1]}-{c[2]}-{c[3]}-{c[4]}")
def test_small_cases_match_scipy(torch, sb, oracle, case, kernel, dtype):
    shape, ang, r, keep, n = case
    raw = synthetic_stack(shape, seed=sum(shape), dtype=dtype)
    want = oracle[0].deskew_data(raw, ang, r, keep, n, cval=7.0)
    got = _gpu(torch, sb, raw, ang, r, keep, n, cval=7.0, kernel=kernel)
    assert got.shape == sb.get_deskewed_data_shape(shape, ang, r, keep, n)[0]
    assert_close_range(got, want, TIGHT_TOL, f"{case} {kernel}")
    # geometry is bit-exact: the set of padded voxels is identical
    assert np.array_equal(got == 7.0, want == 7.0)


@pytest.mark.parametrize("keep", [False, True])
@pytest.mark.parametrize("kernel", ["direct", "tma"])
def test_config1_matches_oracle(torch, sb, oracle, keep, kernel):
    """BASELINE configs[0]: uint16 (101,256,256), 30 deg, r=0.39, n=1 -- checked against scipy itself."""
    raw = synthetic_stack((101, 256, 256), seed=0)
    want = oracle[0].deskew_data(raw, 30.0, 0.39, keep, 1)
    got = _gpu(torch, sb, raw, 30.0, 0.39, keep, 1, kernel=kernel)
    assert got.shape == ((256, 256, 481) if keep else (256, 256, 38))
    rel = assert_close_range(got, want, TIGHT_TOL, "config 1")
    assert rel <= CONTRACT_TOL
    assert np.array_equal(got == 0.0, want == 0.0)


@pytest.mark.parametrize("n", [1, 2, 3, 4])
@pytest.mark.parametrize("keep", [False, True])
def test_mantis_slab_tma_vs_c_oracle(torch, sb, oracle, n, keep):
    """A 1/16-width mantis FOV (600, 300, 128) with the config-2 parameters, TMA kernel vs the C oracle."""
    raw = synthetic_stack((600, 300, 128), seed=1)
    want = oracle[1].deskew_data(raw, 30.0, 0.39, keep, n)
    got = _gpu(torch, sb, raw, 30.0, 0.39, keep, n, kernel="tma")
    assert_close_range(got, want, TIGHT_TOL, f"mantis slab n={n} keep={keep}")
    assert np.array_equal(got == 0.0, want == 0.0)


def test_float32_input_equals_uint16_input(torch, sb):
    """The fused uint16 path is the same function as convert-then-deskew (preprocessing.py:316)."""
    raw = synthetic_stack((120, 30, 192), seed=3)
    a = _gpu(torch, sb, raw, 30.0, 0.39, False, 3, kernel="tma")
    b = _gpu(torch, sb, raw.astype(np.float32), 30.0, 0.39, False, 3, kernel="tma")
    assert np.array_equal(a, b)


def test_tma_equals_direct_bitwise(torch, sb):
    raw = synthetic_stack((200, 47, 256), seed=4)
    for n in (1, 2, 3, 4):
        a = _gpu(torch, sb, raw, 30.0, 0.39, True, n, kernel="tma", cval=-3.5)
        b = _gpu(torch, sb, raw, 30.0, 0.39, True, n, kernel="direct", cval=-3.5)
        assert np.array_equal(a, b), n


def test_x_chunks_reverse_concat_identity(torch, sb):
    """scripts/measure_psf.py:218-249: X-chunks deskewed alone, concatenated in reverse along axis -2."""
    raw = synthetic_stack((90, 20, 256), seed=5)
    whole = _gpu(torch, sb, raw, 30.0, 0.39, True, 3)
    parts = [_gpu(torch, sb, np.ascontiguousarray(c), 30.0, 0.39, True, 3) for c in np.split(raw, 4, axis=-1)]
    assert np.array_equal(np.concatenate(parts[::-1], axis=-2), whole)


def test_noncontiguous_x_chunk_view(torch, sb, oracle):
    """A strided view (raw[:, :, 64:128]) goes through without a copy and matches."""
    raw = synthetic_stack((70, 12, 256), seed=6)
    t = torch.from_numpy(raw).cuda()[:, :, 64:128]
    got = sb.deskew_zyx(t, 30.0, 0.39, False, 3).cpu().numpy()
    want = oracle[1].deskew_data(raw[:, :, 64:128], 30.0, 0.39, False, 3)
    assert_close_range(got, want, TIGHT_TOL, "strided view")


def test_analytic_ramp_and_constant(torch, sb):
    """raw[z,y,x] = z  ->  D = z_in exactly inside; constant volume -> constant inside, cval outside."""
    Z, Y, X = 64, 8, 16
    g = sb.deskew_geometry((Z, Y, X), 30.0, 0.5, True, 1)
    ramp = np.broadcast_to(np.arange(Z, dtype=np.float32)[:, None, None], (Z, Y, X)).copy()
    got = _gpu(torch, sb, ramp, 30.0, 0.5, True, 1, cval=-1.0)
    o0 = np.arange(Y, dtype=np.float64)[:, None]
    o2 = np.arange(g.out_shape[2], dtype=np.float64)[None, :]
    z_in = (g.shift + o0 * g.m00) + o2 * g.m02
    inside = (z_in >= 0) & (z_in <= Z - 1)
    want = np.where(inside, z_in, -1.0).astype(np.float32)
    assert np.array_equal(got[:, 0, :] == -1.0, ~inside)
    assert np.max(np.abs(got[:, 3, :] - want)) <= 1e-5
    const = np.full((Z, Y, X), 1234, dtype=np.uint16)
    got = _gpu(torch, sb, const, 30.0, 0.5, True, 1, cval=-1.0)
    assert np.array_equal(got[:, 5, :], np.where(inside, 1234.0, -1.0).astype(np.float32))


def test_flips_and_single_voxel(torch, sb):
    """One bright voxel lands where the closed form says: axis 0 <- flipped tilt, axis 1 <- flipped x."""
    Z, Y, X = 40, 6, 24
    raw = np.zeros((Z, Y, X), dtype=np.uint16)
    z, y, x = 17, 4, 9
    raw[z, y, x] = 1000
    got = _gpu(torch, sb, raw, 30.0, 0.5, True, 1)
    nz = np.argwhere(got > 0)
    assert set(nz[:, 0]) == {Y - 1 - y} and set(nz[:, 1]) == {X - 1 - x}
    g = sb.deskew_geometry((Z, Y, X), 30.0, 0.5, True, 1)
    o0 = Y - 1 - y
    for o2 in nz[:, 2]:
        z_in = (g.shift + o0 * g.m00) + o2 * g.m02
        assert abs(z_in - z) < 1.0
        assert abs(got[o0, X - 1 - x, o2] - 1000 * (1 - abs(z_in - z))) < 1e-2


def test_average_edge_replication_and_n1(torch, sb):
    raw = synthetic_stack((33, 10, 32), seed=7)
    full = _gpu(torch, sb, raw, 30.0, 0.39, True, 1)
    avg = _gpu(torch, sb, raw, 30.0, 0.39, True, 3)          # 10 % 3 != 0
    padded = np.concatenate([full, full[-1:], full[-1:]], axis=0).reshape(4, 3, *full.shape[1:])
    want = padded.mean(axis=1, dtype=np.float32)
    assert avg.shape[0] == 4
    assert_close_range(avg, want, TIGHT_TOL, "edge replication")


def test_cval_none_uses_min(torch, sb, oracle):
    raw = synthetic_stack((30, 8, 40), seed=8)
    got = _gpu(torch, sb, raw, 30.0, 0.39, True, 2, cval=None)
    want = oracle[0].deskew_data(raw, 30.0, 0.39, True, 2, cval=None)
    assert_close_range(got, want, TIGHT_TOL, "cval=min")
    neg = (raw.astype(np.float32) - 30000.0)
    got = _gpu(torch, sb, neg, 30.0, 0.39, True, 1, cval=None)
    assert got.min() == neg.min()


def test_windows_reassemble_bitwise(torch, sb):
    """Tilt-block / column windows over slabs equal the un-windowed result bit for bit."""
    raw = synthetic_stack((150, 25, 64), seed=9)
    t = torch.from_numpy(raw).cuda()
    g = sb.deskew_geometry(raw.shape, 30.0, 0.39, True, 3)
    whole = sb.deskew_zyx(t, 30.0, 0.39, True, 3)
    Yn, X, Xp = g.out_shape
    out = torch.full_like(whole, float("nan"))
    p_edges = [0, 3, 4, Yn]
    c_edges = [0, 100, 257, Xp]
    for p0, p1 in zip(p_edges[:-1], p_edges[1:]):
        for c0, c1 in zip(c_edges[:-1], c_edges[1:]):
            (y0, y1), (z0, z1) = sb.window_needs(g, p0, p1 - p0, c0, c1 - c0)
            if z1 <= z0:
                z0, z1 = 0, 1
            slab = t[z0:z1, y0:y1, :].contiguous()
            for kernel in ("direct", "auto"):
                piece = sb.deskew_window(slab, g, p_begin=p0, p_count=p1 - p0, c_begin=c0, c_count=c1 - c0,
                                         y_origin=y0, z_origin=z0, kernel=kernel)
                out[p0:p1, :, c0:c1] = piece
                assert torch.equal(out[p0:p1, :, c0:c1], whole[p0:p1, :, c0:c1]), (p0, c0, kernel)
    assert torch.equal(out, whole)


def test_host_pipeline_matches_device_path(torch, sb, oracle):
    """deskew_data (numpy in/out through the streaming pipeline) == device path == oracle."""
    raw = synthetic_stack((160, 41, 128), seed=10)
    dev = _gpu(torch, sb, raw, 30.0, 0.39, False, 3)
    host = sb.deskew_data(raw, 30.0, 0.39, False, 3, device="cuda")
    assert isinstance(host, np.ndarray) and host.dtype == np.float32
    assert np.array_equal(host, dev)
    want = oracle[1].deskew_data(raw, 30.0, 0.39, False, 3)
    assert_close_range(host, want, TIGHT_TOL, "host pipeline")
    # int16 / float64 inputs are cast to float32 first (preprocessing.py:316)
    host64 = sb.deskew_data(raw.astype(np.float64), 30.0, 0.39, False, 3)
    assert np.array_equal(host64, host)


def test_fast_deskew_zyx_contract(torch, sb):
    """Keyword call as shrimpy/preprocessing.py:408-413 makes it; result stays on the input device."""
    raw = torch.from_numpy(synthetic_stack((60, 16, 64), seed=11).astype(np.float32)).cuda()
    out = sb.fast_deskew_zyx(raw_data=raw, ls_angle_deg=30.0, px_to_scan_ratio=0.39, keep_overhang=False,
                             average_n_slices=3)
    assert out.device == raw.device and out.dtype == torch.float32
    assert tuple(out.shape) == sb.get_deskewed_data_shape(tuple(raw.shape), 30.0, 0.39, False, 3)[0]
    cpu = sb.fast_deskew_zyx(raw_data=raw.cpu(), ls_angle_deg=30.0, px_to_scan_ratio=0.39, keep_overhang=False,
                             average_n_slices=3)
    assert cpu.device.type == "cpu" and torch.equal(cpu, out.cpu())


def test_errors_are_python_exceptions(torch, sb):
    from shrimpy_b200._cabi import ShrimpyB200Error

    raw = torch.zeros((8, 4, 10), dtype=torch.uint16, device="cuda")   # X*2 bytes not a multiple of 16
    with pytest.raises(ShrimpyB200Error):
        sb.deskew_zyx(raw, 30.0, 0.39, True, 1, kernel="tma")
    sb.deskew_zyx(raw, 30.0, 0.39, True, 1)                            # auto falls back to the direct kernel
    with pytest.raises(RuntimeError):
        sb.deskew_data(np.zeros((4, 4, 4), np.uint16), 30.0, 0.39, True, device="cpu")
    with pytest.raises(ValueError):
        sb.deskew_zyx(torch.zeros((4, 4), device="cuda"), 30.0, 0.39, True)


def test_launch_counter_moves(torch, sb):
    from shrimpy_b200 import _cabi

    before = _cabi.launch_count()
    sb.deskew_zyx(torch.zeros((8, 4, 16), dtype=torch.uint16, device="cuda"), 30.0, 0.39, True, 1)
    assert _cabi.launch_count() == before + 1


@pytest.mark.parametrize("n,keep,kernel", [(1, False, "tma"), (1, True, "tma"), (3, False, "tma"), (2, True, "direct")])
def test_padded_row_output_is_the_same_volume(torch, sb, n, keep, kernel):
    """``empty_deskewed``: rows padded to whole 32-byte sectors (the write-dominated deskews run 21-26 % faster into such
    a buffer, tools/probe/padded_out_probe.py).  Same voxels, bit for bit; the padding is never written."""
    from helpers import synthetic_stack

    raw = torch.from_numpy(synthetic_stack((90, 13, 128), seed=4)).cuda()
    want = sb.deskew_zyx(raw, 30.0, 0.39, keep, n, kernel=kernel)
    g = sb.deskew_geometry((90, 13, 128), 30.0, 0.39, keep, n)
    out = sb.empty_deskewed(g, raw.device)
    assert tuple(out.shape) == g.out_shape and out.stride(1) % 8 == 0 and out.stride(2) == 1
    assert (out.stride(1) != g.out_shape[2]) == (g.out_shape[2] % 8 != 0)
    storage = out.as_strided((g.out_shape[0], g.out_shape[1], out.stride(1)), out.stride())
    storage.fill_(-123.0)
    got = sb.deskew_zyx(raw, 30.0, 0.39, keep, n, out=out, kernel=kernel)
    assert got.data_ptr() == out.data_ptr() and torch.equal(got, want)
    assert bool((storage[:, :, g.out_shape[2]:] == -123.0).all())
    with pytest.raises(ValueError):
        sb.deskew_zyx(raw, 30.0, 0.39, keep, n, out=out.transpose(1, 2))
    with pytest.raises(ValueError):
        sb.deskew_zyx(raw, 30.0, 0.39, keep, n, out=out, value_range=torch.empty(2, device="cuda"))


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
@pytest.mark.parametrize("shape,r,keep,n", [((90, 13, 128), 0.39, False, 1), ((120, 31, 200), 0.39, True, 3),
                                            ((260, 11, 2048), 0.39, True, 1), ((64, 7, 72), 0.77, False, 2),
                                            ((600, 10, 264), 0.39, False, 1)])
def test_staged_variant_is_the_same_volume(torch, sb, shape, r, keep, n, dtype):
    """``SHRIMPY_KERNEL_TMA_STAGED`` (what AUTO runs for average_n_slices == 1): results staged through shared memory
    and stored as 16-byte vectors on 32-byte boundaries, ragged ends as scalars.  Same voxels as the plain kernel bit for
    bit -- into a contiguous result of odd pitch, into padded rows, and into windows that start inside a sector -- and
    nothing outside the target is touched."""
    raw = torch.from_numpy(synthetic_stack(shape, seed=6, dtype=dtype)).cuda()
    want = sb.deskew_zyx(raw, 30.0, r, keep, n, kernel="tma")
    got = torch.full_like(want, -7.0)
    sb.deskew_zyx(raw, 30.0, r, keep, n, out=got, kernel="tma_staged")
    assert torch.equal(got, want)
    if n == 1:
        assert torch.equal(sb.deskew_zyx(raw, 30.0, r, keep, n), want)          # AUTO
    g = sb.deskew_geometry(shape, 30.0, r, keep, n)
    P, X, Xp = g.out_shape
    padded = sb.empty_deskewed(g, raw.device)
    storage = padded.as_strided((P, X, padded.stride(1)), padded.stride())
    storage.fill_(-7.0)
    sb.deskew_zyx(raw, 30.0, r, keep, n, out=padded, kernel="tma_staged")
    assert torch.equal(padded, want) and bool((storage[:, :, Xp:] == -7.0).all())
    canvas = torch.full((P, X, Xp), -7.0, device="cuda")
    for c0, c1 in ((3, min(Xp, 3 + 61)), (Xp // 3 + 1, Xp - 2), (5, min(Xp, 5 + 256 + 9))):
        if c1 <= c0:
            continue
        _, zr = sb.window_needs(g, 0, P, c0, c1 - c0)
        z0, z1 = (int(zr[0]), int(zr[1])) if zr[1] > zr[0] else (0, 1)
        sb.deskew_window(raw[z0:z1], g, p_begin=0, p_count=P, c_begin=c0, c_count=c1 - c0, y_origin=0, z_origin=z0,
                         out=canvas[:, :, c0:c1], kernel="tma_staged")
        assert torch.equal(canvas[:, :, c0:c1], want[:, :, c0:c1])
        canvas[:, :, c0:c1] = -7.0
        assert bool((canvas == -7.0).all())
    with pytest.raises(Exception):
        sb.deskew_zyx(raw, 30.0, r, keep, n, kernel="tma_staged", value_range=torch.empty(2, device="cuda"))


def test_staged_variant_refuses_what_does_not_fit_and_auto_falls_back(torch, sb):
    from shrimpy_b200._cabi import ShrimpyB200Error

    raw = torch.from_numpy(synthetic_stack((300, 12, 264), seed=8)).cuda()
    with pytest.raises(ShrimpyB200Error):
        sb.deskew_zyx(raw, 30.0, 1.3, True, 1, kernel="tma_staged")      # a 256-column tile spans > 256 scan slices
    assert torch.equal(sb.deskew_zyx(raw, 30.0, 1.3, True, 1), sb.deskew_zyx(raw, 30.0, 1.3, True, 1, kernel="direct"))


def test_broadcast_and_odd_stride_views_are_compacted_not_misread(torch, sb, oracle):
    """A broadcast stack (stride 0 along z) used to reach the C-ABI as "stride 0 = contiguous default" and read
    Z*Y*X elements from a Y*X allocation; every view now deskews to the same volume as its compact copy."""
    plane = torch.from_numpy(synthetic_stack((1, 12, 64), seed=40)[0]).cuda()
    view = plane[None].expand(50, 12, 64)
    assert view.stride(0) == 0
    want = sb.deskew_zyx(view.contiguous(), 30.0, 0.39, True, 2)
    assert torch.equal(sb.deskew_zyx(view, 30.0, 0.39, True, 2), want)
    raw = torch.from_numpy(synthetic_stack((40, 12, 64), seed=41)).cuda()
    for v in (raw.permute(0, 2, 1).contiguous().permute(0, 2, 1),      # x not the fastest axis
              raw.flip(0),                                             # negative-stride-like copy (torch makes it dense)
              raw[::2],                                                # every other slice: stride_z = 2*Y*X, no copy
              raw[:, ::3]):                                            # every third row:   stride_y = 3*X, no copy
        assert torch.equal(sb.deskew_zyx(v, 30.0, 0.39, False, 1), sb.deskew_zyx(v.contiguous(), 30.0, 0.39, False, 1))
    from shrimpy_b200 import flatfield
    assert torch.equal(flatfield.flat_field_pattern(view), plane.to(torch.float32))
    with pytest.raises(ValueError):
        sb.deskew_window(raw, sb.deskew_geometry((40, 12, 64), 30.0, 0.39, True, 1), p_begin=0, p_count=12, c_begin=0,
                         c_count=32, y_origin=0, z_origin=0, out=torch.empty((12, 32, 64), device="cuda").transpose(1, 2))


def test_cval_none_on_an_x_chunk_view_pads_with_the_chunk_minimum(torch, sb, oracle):
    """``cval=None`` pads with min(raw) of the VIEW (what deskew_data's ``raw.min()`` sees), not of the 128
    consecutive-in-memory columns the un-compacted reduction used to walk."""
    raw = synthetic_stack((48, 9, 256), seed=42)
    raw[:, :, :64] = 5                     # a smaller value outside the chunk
    chunk = torch.from_numpy(raw).cuda()[:, :, 128:256]
    assert not chunk.is_contiguous()
    got = sb.deskew_zyx(chunk, 30.0, 0.39, True, 1, cval=None).cpu().numpy()
    want = oracle[0].deskew_data(raw[:, :, 128:256], 30.0, 0.39, True, 1, cval=None)
    assert float(got.min()) == float(raw[:, :, 128:256].min()) != 5.0
    assert np.array_equal(got == got.min(), want == want.min())
    assert_close_range(got, want, TIGHT_TOL, "cval=None on a view")
    assert np.array_equal(got, sb.deskew_data(raw[:, :, 128:256], 30.0, 0.39, True, 1, cval=None))


def test_pageable_and_pinned_callers_get_the_same_bytes(torch, sb):
    """numpy in / numpy out as scripts/measure_psf.py:239-246 calls it: an ordinary (pageable) array goes through the
    pipeline's own page-locked staging rings, a pinned one is copied directly; same result, and the staging really ran."""
    from shrimpy_b200 import _cabi
    from shrimpy_b200.deskew import _pipeline_for
    import ctypes

    raw = synthetic_stack((300, 60, 512), seed=43)
    pinned_in = torch.from_numpy(raw).pin_memory()
    g = sb.deskew_geometry(raw.shape, 30.0, 0.39, False, 3)
    pinned_out = torch.empty(g.out_shape, dtype=torch.float32).pin_memory()
    a = sb.deskew_data(pinned_in.numpy(), 30.0, 0.39, False, 3, out=pinned_out.numpy())
    pipe = _pipeline_for(torch.cuda.current_device())
    si, so = ctypes.c_int64(), ctypes.c_int64()

    def staged():
        _cabi.check(_cabi.lib().shrimpy_pipeline_staged_bytes(pipe._handle, ctypes.byref(si), ctypes.byref(so)))
        return si.value, so.value

    assert staged() == (0, 0)
    b = sb.deskew_data(raw, 30.0, 0.39, False, 3, out=np.empty(g.out_shape, np.float32))
    assert staged() == (raw.nbytes, b.nbytes)
    c = sb.deskew_data(raw, 30.0, 0.39, False, 3)                  # default result: torch's pinned cache
    assert staged() == (raw.nbytes, 0)
    dev = sb.deskew_zyx(torch.from_numpy(raw).cuda(), 30.0, 0.39, False, 3).cpu().numpy()
    assert np.array_equal(a, dev) and np.array_equal(b, dev) and np.array_equal(c, dev)


def test_two_threads_calling_deskew_data_on_one_gpu(torch, sb):
    """An IO pool over positions: each thread gets its own pipeline (slots, streams, events), and one pipeline that IS
    shared is serialised inside the C-ABI -- either way every call returns its own volume."""
    from concurrent.futures import ThreadPoolExecutor
    from shrimpy_b200.deskew import HostPipeline

    stacks = [synthetic_stack((200, 30, 256), seed=50 + i) for i in range(4)]
    want = [sb.deskew_zyx(torch.from_numpy(s).cuda(), 30.0, 0.39, False, 3).cpu().numpy() for s in stacks]
    with ThreadPoolExecutor(4) as pool:
        for _ in range(3):
            got = list(pool.map(lambda s: sb.deskew_data(s, 30.0, 0.39, False, 3), stacks))
            assert all(np.array_equal(g, w) for g, w in zip(got, want))
    g = sb.deskew_geometry(stacks[0].shape, 30.0, 0.39, False, 3)
    with HostPipeline(torch.cuda.current_device()) as shared, ThreadPoolExecutor(4) as pool:
        got = list(pool.map(lambda s: shared.deskew(s, g, 0.0), stacks * 2))
    assert all(np.array_equal(g_, w) for g_, w in zip(got, want * 2))


def test_random_geometries_every_kernel_agrees_bitwise(torch, sb):
    """40 seeded random stacks (ragged X tiles, Y % n != 0, windows, both dtypes, angles 10..45, ratios 0.2..0.9):
    direct == plain TMA == staged TMA == AUTO, bit for bit, and the result equals the C oracle to 2e-6 of range."""
    from oracle import c_oracle
    from shrimpy_b200._cabi import ShrimpyB200Error

    rng = np.random.default_rng(2024)
    refused = 0
    for case in range(40):
        Z, Y = int(rng.integers(40, 400)), int(rng.integers(1, 14))
        X = int(rng.integers(1, 40)) * 8                        # TMA rows want 16-byte multiples
        ang, r = round(float(rng.uniform(10, 45)), 2), round(float(rng.uniform(0.2, 0.9)), 3)
        keep, n = bool(rng.integers(0, 2)), int(rng.integers(1, 5))
        dtype = np.uint16 if case % 3 else np.float32
        raw_np = synthetic_stack((Z, Y, X), seed=100 + case, dtype=dtype)
        g = sb.deskew_geometry((Z, Y, X), ang, r, keep, n)
        if g.out_shape[2] == 0:
            continue
        raw = torch.from_numpy(raw_np).cuda()
        want = sb.deskew_zyx(raw, ang, r, keep, n, kernel="direct")
        for kernel in ("tma", "tma_staged", "auto"):
            try:
                got = sb.deskew_zyx(raw, ang, r, keep, n, kernel=kernel)
            except ShrimpyB200Error as exc:       # forced staged tiles that do not fit shared memory are refused, by name
                assert kernel == "tma_staged" and "too large" in str(exc), (case, kernel, str(exc))
                refused += 1
                continue
            assert torch.equal(got, want), (case, kernel, (Z, Y, X), ang, r, keep, n, dtype)
        if case % 8 == 0:
            ref = c_oracle.deskew_data(raw_np, ang, r, keep, n)
            assert np.array_equal(want.cpu().numpy() == 0.0, ref == 0.0)
            assert_close_range(want.cpu().numpy(), ref, TIGHT_TOL, f"case {case}")
    assert refused < 20


@pytest.mark.parametrize("n,keep", [(1, True), (1, False), (2, False)])
def test_host_pipeline_with_the_staged_kernel(torch, sb, n, keep):
    """``deskew_data`` cuts the stack into tilt slabs and launches window calls: with average_n_slices 1 or 2 those
    run the staged kernel (AUTO) on slabs with a tilt origin -- same volume as the one-launch device path."""
    raw = synthetic_stack((220, 37, 256), seed=80 + n)
    dev = sb.deskew_zyx(torch.from_numpy(raw).cuda(), 30.0, 0.39, keep, n, kernel="tma").cpu().numpy()
    host = sb.deskew_data(raw, 30.0, 0.39, keep, n)
    assert np.array_equal(host, dev)
