"""The deskew as the torch / MONAI generation of biahub evaluates it -- TEST INFRASTRUCTURE ONLY, parity unpinned.

SURVEY.md section 8 (a3/a4, [RECALL]): the ``deskew_data`` generation that shrimPy's pinned biahub revision ships hands
the output-index -> input-index matrix of the deskew to ``monai.transforms.Affine(..., padding_mode="zeros")`` with
``mode="bilinear"``, which ends in ``torch.nn.functional.grid_sample`` on a float32 sampling grid.  Neither biahub nor
MONAI is available offline, so this module restates only what is certain about that generation -- the resampling
machinery -- for the SAME voxel map as the normative scipy form (``deskew_oracle.deskew_affine_matrix``):

* ``affine_grid`` + ``grid_sample(mode="bilinear", padding_mode="zeros", align_corners=True)``, everything float32;
* outside rule: ``padding_mode="zeros"`` blends with zero over one voxel beyond either end of the scan axis
  (``-1 < z_in < 0`` and ``Z-1 < z_in < Z``) where scipy's ``mode="constant"`` returns exactly ``cval`` -- the rim;
* coordinates: float32 instead of scipy's float64, so interior voxels differ by the slope times ~1e-7 * Z voxels, and
  the integer row/column maps (axes 1 and 2) are only integer to float32 rounding.

``tests/test_grid_sample_form.py`` measures both differences on BASELINE config 1 and pins them; the numbers are quoted
in INTEGRATION.md next to the statement of which form this package implements (the scipy form, as the north star says).
Nothing under ``shrimpy_b200/`` imports this module.
"""

from __future__ import annotations

import numpy as np

from . import deskew_oracle as o


def theta_from_index_matrix(M: np.ndarray, in_shape, out_shape) -> np.ndarray:
    """3x4 ``theta`` for ``affine_grid(align_corners=True)`` that realises the index-space map ``i_in = M @ [i_out, 1]``.

    ``affine_grid`` works in normalised coordinates ``u = 2 i / (S - 1) - 1`` with the axes in x, y, z order (last
    array axis first), so ``theta = N_in . M . N_out^-1`` with the axis order reversed on both sides (float64 here,
    cast to float32 by the caller like torch would)."""
    M = np.asarray(M, dtype=np.float64)
    if M.shape == (3, 4):
        M = np.vstack([M, [0, 0, 0, 1]])

    def norm(shape):        # index (z, y, x, 1) -> normalised (z, y, x, 1)
        N = np.eye(4)
        for a, s in enumerate(shape):
            d = max(int(s) - 1, 1)
            N[a, a], N[a, 3] = 2.0 / d, -1.0
        return N

    T = norm(in_shape) @ M @ np.linalg.inv(norm(out_shape))
    flip = np.eye(4)[[2, 1, 0, 3]]           # (z, y, x) <-> (x, y, z)
    return (flip @ T @ flip)[:3]


def deskew_data_grid_sample(raw: np.ndarray, ls_angle_deg: float, px_to_scan_ratio: float, keep_overhang: bool,
                            average_n_slices: int = 1) -> np.ndarray:
    """Deskew through ``affine_grid`` / ``grid_sample`` (bilinear, zero padding, float32), then the edge-padded block
    mean -- the same voxel map, shapes and averaging as ``deskew_oracle.deskew_data``."""
    import torch
    import torch.nn.functional as F

    raw = np.asarray(raw)
    Z, Y, X = raw.shape
    shape, _ = o.get_deskewed_data_shape(raw.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang, 1)
    M = o.deskew_affine_matrix(raw.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang)
    theta = torch.from_numpy(theta_from_index_matrix(M, (Z, Y, X), shape).astype(np.float32))[None]
    vol = torch.from_numpy(raw.astype(np.float32))[None, None]
    grid = F.affine_grid(theta, (1, 1) + tuple(int(s) for s in shape), align_corners=True)
    out = F.grid_sample(vol, grid, mode="bilinear", padding_mode="zeros", align_corners=True)[0, 0].numpy()
    return o.average_n_slices(out, average_n_slices).astype(np.float32)


def compare_forms(raw: np.ndarray, ls_angle_deg: float, px_to_scan_ratio: float, keep_overhang: bool,
                  average_n_slices: int = 1) -> dict:
    """Where and by how much the two forms differ on one stack (cval = 0 on the scipy side)."""
    a = o.deskew_data(raw, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices, cval=0.0)
    b = deskew_data_grid_sample(raw, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices)
    rng = float(a.max() - a.min()) or 1.0
    delta = np.abs(a.astype(np.float64) - b.astype(np.float64)) / rng
    # the rim: output voxels with at least one averaged row whose scan coordinate lies within one voxel OUTSIDE the stack
    g = o.deskew_affine_matrix(raw.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang)
    Z, Y, _ = raw.shape
    P, _, Xp = a.shape
    n = int(average_n_slices)
    o0 = np.minimum(np.arange(P * n), Y - 1).reshape(P, n)
    z = (g[0, 3] + o0[:, :, None] * g[0, 0]) + np.arange(Xp)[None, None, :] * g[0, 2]
    rim = (((z > -1) & (z < 0)) | ((z > Z - 1) & (z < Z))).any(axis=1)              # (P, Xp)
    rim3 = np.broadcast_to(rim[:, None, :], a.shape)
    interior = ~rim3
    return {
        "shape": list(a.shape), "rim_voxel_fraction": float(rim3.mean()),
        "rim_max_delta_of_range": float(delta[rim3].max()) if rim3.any() else 0.0,
        "rim_mean_delta_of_range": float(delta[rim3].mean()) if rim3.any() else 0.0,
        "interior_max_delta_of_range": float(delta[interior].max()),
        "interior_mean_delta_of_range": float(delta[interior].mean()),
        "voxels_beyond_contract_tolerance_1e-3": float((delta > 1e-3).mean()),
        "voxels_beyond_1e-3_outside_the_rim": float((delta[interior] > 1e-3).mean()),
    }
