"""Store rule of the whole-sector deskew variant (``SHRIMPY_KERNEL_TMA_ALIGNED``): CPU proof on the host mirror
(``tools/sector_spans.py``) that every voxel has exactly one owner tile and that only whole sectors leave a tile."""

import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))

import sector_spans as ss  # noqa: E402


@pytest.mark.parametrize("T2", [32, 64, 128, 256])
def test_every_voxel_has_exactly_one_owner(T2):
    rng = np.random.default_rng(T2)
    for _ in range(40):
        width = int(rng.integers(1, 4 * T2))
        row_stride = width + int(rng.integers(0, 9))            # contiguous or padded rows
        base = int(rng.integers(0, 8))                           # any 4-byte alignment of the window's first voxel
        counts = ss.store_counts(base, row_stride, 9, width, T2)
        assert counts.min() == 1 and counts.max() == 1, (width, row_stride, base)


def test_mantis_row_length_leaves_no_partial_sector_inside_a_row():
    # config 2: Xp = 1279 floats per row, contiguous; config 5: 10517
    for width in (1279, 1799, 10517):
        assert ss.partial_sector_ends(0, width, 16, width, 256, aligned=True) == 0
        # the plain tiling leaves two partial ends per tile boundary in almost every row
        plain = ss.partial_sector_ends(0, width, 16, width, 256, aligned=False)
        assert plain >= 2 * (ss.tile_layout(width, 256, False)[1] - 1) * 13
        assert (ss.store_counts(0, width, 16, width, 256, aligned=False) == 1).all()


def test_window_edges_and_padded_rows():
    # a window that starts inside a sector: tile 0 also owns the leading partial sector
    counts = ss.store_counts(base_elem=5, row_stride=1279, rows=8, width=700, T2=256)
    assert (counts == 1).all()
    assert ss.partial_sector_ends(5, 1279, 8, 700, 256) == 0
    # rows padded to whole sectors: the aligned rule degenerates to 248-column tiles, still one owner each
    assert (ss.store_counts(0, 1280, 4, 1279, 256) == 1).all()
    assert ss.tile_layout(1279, 256, True) == (248, 6) and ss.tile_layout(1279, 256, False) == (256, 5)
