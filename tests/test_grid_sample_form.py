"""How far the torch / MONAI generation of the reference deskew (``affine_grid`` + ``grid_sample``, bilinear, zero
padding, float32 -- ``oracle/grid_sample_form.py``) lies from the scipy form this package implements, on BASELINE
config 1.  The two differ only (i) in a rim one voxel thick beyond either end of the scan axis, where zero padding
blends with 0 and scipy's ``mode="constant"`` returns ``cval``, and (ii) by float32 coordinate rounding elsewhere.
The measured numbers are pinned here and quoted in INTEGRATION.md."""

import numpy as np
import pytest

from oracle import deskew_oracle as o
from oracle import grid_sample_form as gs


def test_theta_reproduces_the_index_map():
    """The normalised ``theta`` handed to ``affine_grid`` is the scipy-convention matrix in other units."""
    import torch
    import torch.nn.functional as F

    shape_in, keep = (23, 10, 7), True
    (shape_out, _), M = o.get_deskewed_data_shape(shape_in, 30.0, 0.39, keep, 1), o.deskew_affine_matrix(shape_in, 30.0, 0.39, keep)
    theta = torch.from_numpy(gs.theta_from_index_matrix(M, shape_in, shape_out))[None]
    grid = F.affine_grid(theta, (1, 1) + tuple(shape_out), align_corners=True)[0].numpy()      # (..., xyz) in [-1, 1]
    idx = np.stack(np.meshgrid(*[np.arange(s) for s in shape_out], indexing="ij"), axis=-1)
    want = idx @ M[:3, :3].T + M[:3, 3]                                                        # (..., zyx) input indices
    got = (grid[..., ::-1] + 1) / 2 * (np.array(shape_in) - 1)
    assert np.max(np.abs(got - want)) < 1e-9


@pytest.mark.parametrize("keep,rim_fraction", [(False, 4.2e-4), (True, 1.07e-2)])
def test_rim_and_interior_delta_on_config_1(keep, rim_fraction):
    raw = np.random.default_rng(0).integers(100, 60000, size=(101, 256, 256), dtype=np.uint16)
    d = gs.compare_forms(raw, 30.0, 0.39, keep, 1)
    assert d["shape"] == [256, 256, 38 if not keep else 481]
    # (i) the rim: a few voxels per output row, but there the two forms differ by up to the whole dynamic range
    assert d["rim_voxel_fraction"] == pytest.approx(rim_fraction, rel=0.05)
    assert 0.5 < d["rim_max_delta_of_range"] <= 1.0
    # (ii) everywhere else: float32 coordinates against float64, 1e-5 of range on white noise, never near the contract
    assert d["interior_max_delta_of_range"] < 1e-4 and d["interior_mean_delta_of_range"] < 1e-5
    assert d["voxels_beyond_1e-3_outside_the_rim"] == 0.0
    assert d["voxels_beyond_contract_tolerance_1e-3"] <= d["rim_voxel_fraction"]
