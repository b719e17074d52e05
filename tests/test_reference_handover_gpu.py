"""The reference's own hand-over to the deskew, replayed on a GPU.

``_LabelfreePreprocessor`` does, per acquired stack (``/root/reference`` is not on the GPU box, so the steps are
restated here with the lines they come from):

    preprocessing.py:137-140   DeskewSettings(**deskew)                       from biahub.settings
    preprocessing.py:225-231   get_deskewed_data_shape(raw_data_shape=zyx, **_settings_kwargs(...))   (warm_up)
    preprocessing.py:316       volume = torch.as_tensor(volume_bf, device=self._device, dtype=torch.float32)
    preprocessing.py:44-56     kwargs = fields of settings.model_dump() that fast_deskew_zyx's signature accepts
    preprocessing.py:408-413   result = fast_deskew_zyx(raw_data=volume, **kwargs)         from biahub.deskew
    preprocessing.py:357-363   require_gpu: no output channel may sit on the CPU
    preprocessing.py:244       the warm-up shape is what the next step (phase transfer function) was built for
"""

import inspect

import numpy as np
import pytest

from helpers import TIGHT_TOL, assert_close_range, synthetic_stack

pytestmark = pytest.mark.gpu


def _settings_kwargs(func, settings):
    accepted = set(inspect.signature(func).parameters)
    return {k: v for k, v in settings.model_dump().items() if k in accepted}


@pytest.mark.parametrize("deskew", [
    {"ls_angle_deg": 30.0, "keep_overhang": False, "average_n_slices": 3, "pixel_size_um": 0.1133, "scan_step_um": 0.174},
    {"ls_angle_deg": 30, "px_to_scan_ratio": 0.374, "pixel_size_um": 0.116, "keep_overhang": True, "average_n_slices": 1},
], ids=["dynatrack_demo.yaml", "test_dynatrack.py:1073"])
def test_the_reference_hand_over_runs_on_the_gpu_through_the_shim(deskew):
    import torch

    import shrimpy_b200 as sb
    from oracle import deskew_oracle

    sb.install_biahub_shim()
    from biahub.deskew import fast_deskew_zyx, get_deskewed_data_shape      # the literal imports of the reference
    from biahub.settings import DeskewSettings

    device = torch.device("cuda", torch.cuda.current_device())
    zyx = (120, 40, 192)
    settings = DeskewSettings(**deskew)
    warm_shape, voxel = get_deskewed_data_shape(raw_data_shape=zyx, **_settings_kwargs(get_deskewed_data_shape, settings))

    volume_bf = synthetic_stack(zyx, seed=60)                                  # np.stack(frames): uint16 (Z, Y, X)
    volume = torch.as_tensor(volume_bf, device=device, dtype=torch.float32)    # :316
    kwargs = _settings_kwargs(fast_deskew_zyx, settings)
    assert set(kwargs) == {"ls_angle_deg", "px_to_scan_ratio", "keep_overhang", "average_n_slices"}
    result = fast_deskew_zyx(raw_data=volume, **kwargs)                        # :408-413
    torch.cuda.synchronize()

    channels = {"GFP": result}
    assert not [n for n, t in channels.items() if t.device.type == "cpu"]     # :357-363 require_gpu
    assert result.device == volume.device and result.dtype == torch.float32
    assert tuple(result.shape) == tuple(warm_shape)                            # :244
    assert volume.dtype == torch.float32 and torch.equal(volume.cpu(), torch.from_numpy(volume_bf.astype(np.float32)))

    want = deskew_oracle.deskew_data(volume_bf, settings.ls_angle_deg, settings.px_to_scan_ratio,
                                     settings.keep_overhang, settings.average_n_slices)
    got = result.cpu().numpy()
    assert np.array_equal(got == 0.0, want == 0.0)                             # the padded set, bit for bit
    assert_close_range(got, want, TIGHT_TOL, "reference hand-over")
    # the uint16 stack handed over directly (the fused convert) is the same volume, and a later empty_cache()
    # (worker.py:273-281) leaves the result intact
    direct = sb.deskew_zyx(torch.from_numpy(volume_bf).to(device), **kwargs)
    torch.cuda.empty_cache()
    assert torch.equal(direct, result)
