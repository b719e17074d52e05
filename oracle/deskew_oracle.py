"""CPU oracle for the light-sheet deskew / affine-resample hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``shrimpy_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs use it, as the checker (never as the product path).

PARITY UNPINNED.  shrimPy does not contain the deskew arithmetic: it calls
``biahub.deskew.fast_deskew_zyx`` / ``get_deskewed_data_shape`` /
``biahub.settings.DeskewSettings`` (reference call sites
``shrimpy/preprocessing.py:138,226-231,408-413`` and
``scripts/measure_psf.py:15-17,230-246``).  ``biahub`` is a git dependency pinned
at ``b011bca57d5bf15c777839e4250e594f7471af3e`` (``pyproject.toml:91``,
``uv.lock:323-325``, version ``0.0.1rc2.post17``), it is not vendored under
``/root/reference`` and not installable offline, and no reference test holds a
golden deskewed voxel or shape (``shrimpy/tests/test_preprocessing.py:12-13``).
This file therefore restates the *published* scipy/numpy form of that
algorithm (the "mantis" deskew: one ``scipy.ndimage.affine_transform`` with
``order=1`` followed by an edge-padded block mean), and anchors it on the
in-repo witnesses:

* ``scripts/measure_psf.py:218-249`` -- X-chunks deskewed independently and
  concatenated in *reverse* along output axis -2 reproduce the whole volume,
  hence output axis 1 is raw axis 2 flipped and the resample never mixes
  raw-x columns;
* ``shrimpy/viewer/_napari_process.py:202-216`` -- one deskewed plane needs
  exactly one tilt row across all scan slices, hence output axis 0 is the raw
  tilt axis and interpolation is 1-D along the scan axis;
* ``scripts/measure_psf.py:225`` -- ``px_to_scan_ratio`` is rounded to three
  decimals.

Two independent statements of the same arithmetic are kept so the oracle
checks itself: the scipy form (normative) and a closed-form numpy 1-D lerp
gather (``*_closed_form``).  scipy's coordinate arithmetic was probed in this
container (scipy 1.18.1): coordinates are accumulated in float64 as
``((shift + o0*M00) + o1*M01) + o2*M02`` with no fused multiply-add, and the
``mode="constant"`` inside test is strict (``0 <= c <= len-1``, no tolerance).

Scope: FINITE raw voxels (camera stacks are uint16).  With a NaN or infinite float32 voxel the statements stop
agreeing, by construction: scipy multiplies all eight trilinear taps, zero weights included, so the voxel poisons
outputs whose zero-weight y/x taps touch it (``0 * nan``), while the closed form here, the C restatement and the CUDA
kernels interpolate along scan only; and the kernels evaluate ``a + w (b - a)`` where scipy evaluates
``(1 - w) a + w b``, so an infinite tap becomes NaN there and stays infinite here
(``tests/test_oracle.py::test_non_finite_voxels_are_outside_the_parity_contract``).

Open ambiguities that cannot be resolved offline (recorded, not hidden):
default ``cval`` (scipy generation used ``min(raw)``, the torch/monai
generation pads with zeros -- the oracle takes ``cval`` explicitly, default
0.0); whether the pinned ``fast_deskew_zyx`` rotates its result (the stale
comment at ``shrimpy/preprocessing.py:224`` mentions ``np.rot90`` but
``_deskew`` at ``:406-417`` performs none -- the oracle does not rotate).
"""

from __future__ import annotations

import numpy as np

__all__ = [
    "get_deskewed_data_shape",
    "deskew_affine_matrix",
    "average_n_slices",
    "deskew_data",
    "deskew_data_closed_form",
    "apply_affine_transform",
    "apply_affine_transform_closed_form",
    "round_settings",
]


def round_settings(ls_angle_deg: float, pixel_size_um: float | None = None,
                   scan_step_um: float | None = None,
                   px_to_scan_ratio: float | None = None) -> tuple[float, float]:
    """Rounding rules of the settings model.

    ``ls_angle_deg`` -> 2 decimals, ``px_to_scan_ratio`` -> 3 decimals (derived
    from ``pixel_size_um / scan_step_um`` when absent;
    ``scripts/measure_psf.py:225`` shows the same 3-decimal rounding).
    """
    if px_to_scan_ratio is None:
        if pixel_size_um is None or scan_step_um is None:
            raise ValueError("need px_to_scan_ratio or pixel_size_um and scan_step_um")
        px_to_scan_ratio = pixel_size_um / scan_step_um
    return round(float(ls_angle_deg), 2), round(float(px_to_scan_ratio), 3)


def get_deskewed_data_shape(raw_data_shape, ls_angle_deg, px_to_scan_ratio, keep_overhang,
                            average_n_slices=1, pixel_size_um=1):
    """Deskewed (post-average) shape and voxel size.

    Call sites: ``shrimpy/preprocessing.py:228-231``, ``scripts/measure_psf.py:230-234``.
    float64 arithmetic then ``ceil``; integers must be bit-exact.
    """
    theta = ls_angle_deg * np.pi / 180
    st = np.sin(theta)
    ct = np.cos(theta)
    Z, Y, X = raw_data_shape
    if keep_overhang:
        Xp = int(np.ceil((Z / px_to_scan_ratio) + (Y * ct)))
    else:
        Xp = int(np.ceil((Z / px_to_scan_ratio) - (Y * ct)))
    shape = (int(np.ceil(Y / average_n_slices)), X, Xp)
    voxel_size = (average_n_slices * st * pixel_size_um, pixel_size_um, pixel_size_um)
    return shape, voxel_size


def deskew_affine_matrix(raw_data_shape, ls_angle_deg, px_to_scan_ratio, keep_overhang):
    """4x4 homogeneous matrix, output index -> input index (scipy convention)."""
    Z, Y, X = raw_data_shape
    ct = np.cos(ls_angle_deg * np.pi / 180)
    Z_shift = 0
    if not keep_overhang:
        Z_shift = int(np.floor(Y * ct * px_to_scan_ratio))
    return np.array(
        [
            [-px_to_scan_ratio * ct, 0, px_to_scan_ratio, Z_shift],
            [-1, 0, 0, Y - 1],
            [0, -1, 0, X - 1],
            [0, 0, 0, 1],
        ],
        dtype=np.float64,
    )


def average_n_slices(data: np.ndarray, average_window_width: int = 1) -> np.ndarray:
    """Edge-padded block mean over axis 0, kept in the input dtype."""
    n = int(average_window_width)
    rem = data.shape[0] % n
    if rem:
        data = np.pad(data, [(0, n - rem)] + [(0, 0)] * (data.ndim - 1), mode="edge")
    blocks = data.reshape((data.shape[0] // n, n) + data.shape[1:])
    return np.mean(blocks, axis=1)


def deskew_data(raw_data: np.ndarray, ls_angle_deg: float, px_to_scan_ratio: float,
                keep_overhang: bool, average_n_slices_: int = 1, cval: float | None = 0.0,
                **kw) -> np.ndarray:
    """Normative scipy form.

    The shrimPy path converts to float32 before the call
    (``shrimpy/preprocessing.py:316``), so does the oracle.  ``cval=None``
    reproduces the scipy-generation default ``min(raw)``.
    """
    import scipy.ndimage

    if "average_n_slices" in kw:
        average_n_slices_ = kw.pop("average_n_slices")
    if kw:
        raise TypeError(f"unexpected arguments {sorted(kw)}")
    raw = np.asarray(raw_data).astype(np.float32)
    if cval is None:
        cval = float(raw.min())
    matrix = deskew_affine_matrix(raw.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang)
    (_, X, Xp), _ = get_deskewed_data_shape(raw.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang)
    Y = raw.shape[1]
    deskewed = scipy.ndimage.affine_transform(
        raw, matrix, output_shape=(Y, X, Xp), order=1, mode="constant", cval=cval
    )
    return average_n_slices(deskewed, average_n_slices_)


def deskew_data_closed_form(raw_data: np.ndarray, ls_angle_deg: float, px_to_scan_ratio: float,
                            keep_overhang: bool, average_n_slices_: int = 1,
                            cval: float | None = 0.0) -> np.ndarray:
    """Independent restatement without scipy: per-(tilt row) 1-D lerp gather along scan.

    ``out[p,o1,o2] = mean_k lerp(raw[:, Y-1-o0_k, X-1-o1], z_in(o0_k,o2))`` with
    ``z_in = (Z_shift + o0*M00) + o2*M02`` in float64.
    The lerp itself is evaluated in float64 and rounded once to float32, as
    scipy does; the block mean is float32.
    """
    raw = np.asarray(raw_data).astype(np.float32)
    if cval is None:
        cval = float(raw.min())
    Z, Y, X = raw.shape
    M = deskew_affine_matrix(raw.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang)
    (_, _, Xp), _ = get_deskewed_data_shape(raw.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang)
    n = int(average_n_slices_)
    Yn = -(-Y // n)
    o2 = np.arange(Xp, dtype=np.float64)
    out = np.empty((Yn, X, Xp), dtype=np.float32)
    plane = np.empty((n, X, Xp), dtype=np.float32)
    for p in range(Yn):
        for k in range(n):
            o0 = min(n * p + k, Y - 1)
            z_in = (M[0, 3] + np.float64(o0) * M[0, 0]) + o2 * M[0, 2]
            inside = (z_in >= 0) & (z_in <= Z - 1)
            zc = np.where(inside, z_in, 0.0)
            z0 = np.floor(zc).astype(np.int64)
            w = zc - z0
            z1 = np.minimum(z0 + 1, Z - 1)
            col = raw[:, Y - 1 - o0, ::-1].astype(np.float64)  # (Z, X) with x flipped
            v = (1.0 - w)[None, :] * col[z0, :].T + w[None, :] * col[z1, :].T  # (X, Xp)
            v = np.where(inside[None, :], v, np.float64(cval))
            plane[k] = v.astype(np.float32)
        out[p] = np.mean(plane, axis=0)
    return out


def apply_affine_transform(zyx_data: np.ndarray, matrix: np.ndarray, output_shape_zyx,
                           cval: float = 0.0) -> np.ndarray:
    """Registration resample: trilinear, constant outside.

    ``matrix`` is 4x4 (or 3x4), output index -> input index in ZYX voxel units.
    NaNs are zeroed first, as the upstream scipy method does.
    """
    import scipy.ndimage

    vol = np.nan_to_num(np.asarray(zyx_data).astype(np.float32))
    return scipy.ndimage.affine_transform(
        vol, np.asarray(matrix, dtype=np.float64), output_shape=tuple(output_shape_zyx),
        order=1, mode="constant", cval=cval,
    )


def apply_affine_transform_closed_form(zyx_data: np.ndarray, matrix: np.ndarray, output_shape_zyx,
                                       cval: float = 0.0) -> np.ndarray:
    """Independent numpy trilinear gather with scipy's coordinate order and inside rule."""
    vol = np.nan_to_num(np.asarray(zyx_data).astype(np.float32)).astype(np.float64)
    M = np.asarray(matrix, dtype=np.float64)
    dims = vol.shape
    oz, oy, ox = (int(s) for s in output_shape_zyx)
    o = np.meshgrid(np.arange(oz, dtype=np.float64), np.arange(oy, dtype=np.float64),
                    np.arange(ox, dtype=np.float64), indexing="ij")
    inside = np.ones((oz, oy, ox), dtype=bool)
    lo, hi, wt = [], [], []
    for a in range(3):
        c = M[a, 3] + o[0] * M[a, 0]
        c = c + o[1] * M[a, 1]
        c = c + o[2] * M[a, 2]
        inside &= (c >= 0) & (c <= dims[a] - 1)
        c = np.where((c >= 0) & (c <= dims[a] - 1), c, 0.0)
        i0 = np.floor(c).astype(np.int64)
        lo.append(i0)
        hi.append(np.minimum(i0 + 1, dims[a] - 1))
        wt.append(c - i0)
    acc = np.zeros((oz, oy, ox), dtype=np.float64)
    for dz in (0, 1):
        iz = hi[0] if dz else lo[0]
        wz = wt[0] if dz else 1.0 - wt[0]
        for dy in (0, 1):
            iy = hi[1] if dy else lo[1]
            wy = wt[1] if dy else 1.0 - wt[1]
            for dx in (0, 1):
                ix = hi[2] if dx else lo[2]
                wx = wt[2] if dx else 1.0 - wt[2]
                acc += wz * wy * wx * vol[iz, iy, ix]
    return np.where(inside, acc, np.float64(cval)).astype(np.float32)
