"""Host mirror of the store rule of ``deskew_tma_kernel<..., ALIGNED=true>`` (``csrc/deskew.cu``,
``SHRIMPY_KERNEL_TMA_ALIGNED``): which tile stores which voxel of an output row.

A deskewed row is ``Xp`` floats long and ``Xp`` is usually odd, so in a contiguous result a tile's 1 KB row segment
starts and ends inside 32-byte DRAM sectors, and the other half of each of those sectors is written by another CTA at
another time (measured cost on B200: 21-26 % of the write-dominated deskews).  The aligned
variant lets neighbouring o2 tiles overlap by 8 columns -- tiles advance by ``T2 - 8`` columns, the CTA's threads still
cover ``T2`` -- and gives every sector of a row to exactly one tile: the tile holding the column of the sector's first
float (clamped to the window's first column).  A thread decides from its own store address:

    a = (address / 4) & 7          place of the voxel inside its sector
    t = column - tile's first column
    store  iff  (t - a >= 0 or tile == 0)  and  t - a < T2 - 8

``store_counts`` replays that rule for every (tile, thread, row) exactly as the kernel evaluates it;
``tests/test_sector_spans.py`` checks that every voxel is stored exactly once and that every span a tile stores
begins and ends on a sector boundary unless it touches the window's edge.
"""

from __future__ import annotations

import numpy as np

SECTOR_FLOATS = 8


def tile_layout(width: int, T2: int, aligned: bool):
    """(step between the first columns of neighbouring tiles, number of tiles) for a window ``width`` columns wide."""
    step = T2 - SECTOR_FLOATS if aligned else T2
    return step, -(-width // step)


def stores(base_elem: int, row_stride: int, row: int, width: int, T2: int, tile: int, aligned: bool = True) -> np.ndarray:
    """Window-local columns that ``tile`` stores in ``row``.  ``base_elem`` is the float index of the window's voxel
    (row 0, column 0) counted from a 32-byte boundary (``address / 4``); ``row_stride`` the row pitch in floats."""
    step, _ = tile_layout(width, T2, aligned)
    t = np.arange(T2)
    col = tile * step + t
    ok = col < width
    if aligned:
        a = (base_elem + row * row_stride + col) & (SECTOR_FLOATS - 1)
        own_hi = np.full(T2, 7) if tile == 0 else np.minimum(t, 7)
        own_lo = t - step
        ok &= (a > own_lo) & (a <= own_hi)
    return col[ok]


def store_counts(base_elem: int, row_stride: int, rows: int, width: int, T2: int, aligned: bool = True) -> np.ndarray:
    """How many tiles store each voxel of a ``(rows, width)`` window (must be 1 everywhere)."""
    _, tiles = tile_layout(width, T2, aligned)
    counts = np.zeros((rows, width), dtype=np.int32)
    for row in range(rows):
        for tile in range(tiles):
            counts[row, stores(base_elem, row_stride, row, width, T2, tile, aligned)] += 1
    return counts


def partial_sector_ends(base_elem: int, row_stride: int, rows: int, width: int, T2: int, aligned: bool = True) -> int:
    """Number of (tile, row) span ends that fall inside a sector without being an end of the window's row."""
    _, tiles = tile_layout(width, T2, aligned)
    partial = 0
    for row in range(rows):
        for tile in range(tiles):
            cols = stores(base_elem, row_stride, row, width, T2, tile, aligned)
            if cols.size == 0:
                continue
            assert np.array_equal(cols, np.arange(cols[0], cols[-1] + 1)), "a tile's span of a row must be contiguous"
            first = base_elem + row * row_stride + int(cols[0])
            last = base_elem + row * row_stride + int(cols[-1]) + 1
            partial += int(first % SECTOR_FLOATS != 0 and cols[0] != 0)
            partial += int(last % SECTOR_FLOATS != 0 and cols[-1] != width - 1)
    return partial
