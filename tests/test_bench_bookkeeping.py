"""The byte counts bench.py divides by are checked against brute force on small cases (CPU) and, for the affine
block's touched-input count, against a numpy scatter on the GPU."""

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
import shrimpy_b200 as sb  # noqa: E402


@pytest.mark.parametrize("shape,keep,n", [((40, 11, 5), False, 3), ((40, 11, 5), True, 1), ((101, 32, 4), False, 1),
                                          ((23, 10, 7), False, 2)])
def test_needed_input_voxels_is_the_span_of_taps_per_tilt_row(shape, keep, n):
    """``roofline.frac_needed_bytes``: per tilt row the kernel addresses the scan slices between the taps of its first
    and last inside column (a contiguous range: the scan coordinate is monotone in o2)."""
    g = sb.deskew_geometry(shape, 30.0, 0.39, keep, n)
    Z, Y, X = shape
    Xp = g.out_shape[2]
    touched = np.zeros((Z, Y), dtype=bool)
    for o0 in range(Y):
        z = (g.shift + o0 * g.m00) + np.arange(Xp) * g.m02
        z = z[(z >= 0) & (z <= Z - 1)]
        if z.size:
            lo, hi = int(np.floor(z.min())), min(int(np.floor(z.max())) + 1, Z - 1)
            touched[lo:hi + 1, Y - 1 - o0] = True
    assert bench.needed_input_voxels(g) == int(touched.sum()) * X
    assert bench.needed_input_voxels(g) <= Z * Y * X


def test_config_2_reads_five_sixths_of_the_stack():
    g = sb.deskew_geometry((600, 300, 2048), 30.0, 0.39, False, 3)
    assert bench.needed_input_voxels(g) * 2 == 614_907_904          # ncu: 620 MB of DRAM reads per launch
    assert g.algorithmic_bytes == (368_640_000, 261_939_200)


@pytest.mark.gpu
def test_touched_input_voxels_against_a_numpy_scatter():
    from tools import bench_blocks

    in_shape, out_shape = (9, 20, 24), (7, 18, 30)
    th = np.deg2rad(4.0)
    M = np.array([[1.02, 0.01, 0.02, -0.4], [0.0, np.cos(th) * 0.9, -np.sin(th), 1.5], [0.03, np.sin(th), np.cos(th) * 0.8, 0.7],
                  [0, 0, 0, 1.0]])
    idx = np.stack(np.meshgrid(*[np.arange(s) for s in out_shape], indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float64)
    c = np.stack([((M[a, 3] + idx[:, 0] * M[a, 0]) + idx[:, 1] * M[a, 1]) + idx[:, 2] * M[a, 2] for a in range(3)], axis=1)
    inside = np.all((c >= 0) & (c <= np.array(in_shape) - 1), axis=1)
    f = np.floor(c[inside]).astype(int)
    hit = np.zeros(in_shape, dtype=bool)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                z = np.minimum(f[:, 0] + dz, in_shape[0] - 1)
                y = np.minimum(f[:, 1] + dy, in_shape[1] - 1)
                x = np.minimum(f[:, 2] + dx, in_shape[2] - 1)
                hit[z, y, x] = True
    assert bench_blocks.touched_input_voxels(in_shape, M, out_shape) == int(hit.sum())
    assert bench_blocks.touched_input_voxels(in_shape, np.eye(4), in_shape) == int(np.prod(in_shape))
