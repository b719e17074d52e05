"""Byte-paged scan split on ONE GPU: every rank's pages in one process, windows stitched with cuMemMap, the window
kernel reading through a stitched range (the driver maps whole handles only, hence the segments of the plan)."""

import pytest

from helpers import synthetic_stack

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("world,keep,n", [(1, True, 1), (3, True, 1), (4, False, 3)])
def test_stitched_windows_on_one_device_match_the_whole_stack(world, keep, n):
    import torch

    import shrimpy_b200 as sb
    from shrimpy_b200 import paged_stack as ps

    shape = (700, 24, 2048)                                 # 96 KiB slices: a 2 MiB page holds 21.3 of them
    raw = torch.from_numpy(synthetic_stack(shape, seed=31)).cuda()
    g = sb.deskew_geometry(shape, 30.0, 0.39, keep, n)
    page = ps.PagedStack.granularity(0)
    shards = ps.plan_paged_split(g, world, shape[1] * shape[2] * 2, page)
    stacks = ps.PagedStack.on_one_device(shards, shape[1:], torch.uint16, 0, page)
    try:
        for st in stacks:
            st.fill_own(lambda z0, z1: raw[z0:z1])
        for st in stacks:
            st.barrier()
        want = sb.deskew_zyx(raw, 30.0, 0.39, keep, n)
        pieces = [ps.deskew_paged_split(st, g, st.shard) for st in stacks]
        torch.cuda.synchronize()
        assert torch.equal(torch.cat(pieces, dim=2), want)
        if world > 1:
            assert any(len(st.shard.maps) > 1 for st in stacks)      # some window really is stitched from two handles
    finally:
        for st in stacks:
            st.close()
