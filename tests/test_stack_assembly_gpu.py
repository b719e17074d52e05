"""FrameStack: frames land once in pinned memory, ship to the GPU while the stack is being collected."""

import numpy as np
import pytest

from helpers import TIGHT_TOL, assert_close_range, synthetic_stack

pytestmark = pytest.mark.gpu


def test_frames_in_any_order_reach_the_device_and_deskew():
    import torch

    import shrimpy_b200 as sb
    from oracle import c_oracle
    from shrimpy_b200.stack_assembly import FrameStack

    raw = synthetic_stack((70, 12, 64), seed=3)
    fs = FrameStack(raw.shape, h2d_batch=16, slots=2)
    order = np.random.default_rng(0).permutation(70)
    for i, z in enumerate(order):
        done = fs.put(int(z), raw[z])
        assert done == (i == 69)
    dev = fs.finish()
    assert dev.dtype == torch.uint16 and dev.is_cuda and tuple(dev.shape) == raw.shape
    out = sb.fast_deskew_zyx(raw_data=dev, ls_angle_deg=30.0, px_to_scan_ratio=0.39, keep_overhang=False,
                             average_n_slices=3)
    fs.release()
    assert np.array_equal(dev.cpu().numpy(), raw)
    assert fs.h2d_bytes == raw.nbytes
    want = c_oracle.deskew_data(raw, 30.0, 0.39, False, 3)
    assert_close_range(out.cpu().numpy(), want, TIGHT_TOL, "frame stack")


def test_slots_rotate_and_errors():
    from shrimpy_b200.stack_assembly import FrameStack

    fs = FrameStack((5, 4, 8), slots=2, h2d_batch=2)
    stacks = [synthetic_stack((5, 4, 8), seed=s) for s in range(3)]
    devs = []
    for raw in stacks:
        for z in range(5):
            fs.put(z, raw[z])
        devs.append(fs.finish())
        fs.release()
    assert np.array_equal(devs[1].cpu().numpy(), stacks[1])
    assert np.array_equal(devs[2].cpu().numpy(), stacks[2])          # slot 0 reused by the third stack
    assert devs[0].data_ptr() == devs[2].data_ptr()
    fs.put(0, stacks[0][0])
    with pytest.raises(RuntimeError, match="incomplete"):
        fs.finish()
    with pytest.raises(IndexError):
        fs.put(9, stacks[0][0])
    with pytest.raises(ValueError):
        fs.put(1, np.zeros((4, 9), np.uint16))
