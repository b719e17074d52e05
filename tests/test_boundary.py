"""Host-side boundary: settings model, geometry, shim imports, and the reference's own call pattern
(``inspect.signature`` filtering, ``shrimpy/preprocessing.py:44-56``).  CPU only."""

import inspect
import json
import sys
from pathlib import Path

import numpy as np
import pydantic
import pytest

import shrimpy_b200 as sb
from oracle import deskew_oracle as o

GOLDEN = Path(__file__).parent / "golden"


# ---- a1: DeskewSettings ---------------------------------------------------------------------------
def test_settings_from_the_reference_config_block():
    # config/mda/mantis/dynatrack_demo.yaml:161-164 + the keys manager.py:297-299 injects
    s = sb.DeskewSettings(ls_angle_deg=30.0, keep_overhang=False, average_n_slices=3,
                          pixel_size_um=0.1133, scan_step_um=0.174)
    assert s.px_to_scan_ratio == 0.651
    assert s.model_dump() == {"pixel_size_um": 0.1133, "ls_angle_deg": 30.0, "px_to_scan_ratio": 0.651,
                              "scan_step_um": 0.174, "keep_overhang": False, "average_n_slices": 3}
    # shrimpy/tests/test_dynatrack.py:1073 uses these numbers
    assert sb.DeskewSettings(pixel_size_um=0.116, scan_step_um=0.31, ls_angle_deg=30).px_to_scan_ratio == 0.374


def test_settings_defaults_and_rounding():
    s = sb.DeskewSettings(pixel_size_um=0.116, ls_angle_deg=30.004, px_to_scan_ratio=0.38961)
    assert (s.ls_angle_deg, s.px_to_scan_ratio, s.keep_overhang, s.average_n_slices) == (30.0, 0.39, False, 3)
    assert getattr(s, "scan_step_um", None) is None


@pytest.mark.parametrize("bad", [
    {"pixel_size_um": 0.1, "ls_angle_deg": 30},                                # neither ratio nor step
    {"pixel_size_um": 0.1, "ls_angle_deg": 50, "px_to_scan_ratio": 0.3},       # angle > 45
    {"pixel_size_um": 0.1, "ls_angle_deg": -5, "px_to_scan_ratio": 0.3},
    {"pixel_size_um": -1, "ls_angle_deg": 30, "px_to_scan_ratio": 0.3},
    {"pixel_size_um": 0.1, "ls_angle_deg": 30, "px_to_scan_ratio": 0.3, "average_n_slices": 0},
    {"pixel_size_um": 0.1, "ls_angle_deg": 30, "px_to_scan_ratio": 0.3, "typo_key": 1},
    {"ls_angle_deg": 30, "px_to_scan_ratio": 0.3},                             # pixel size is required
])
def test_settings_reject_bad_input(bad):
    with pytest.raises((pydantic.ValidationError, ValueError)):
        sb.DeskewSettings(**bad)


# ---- a2: geometry ------------------------------------------------------------------------------------
def test_shape_known_answers():
    assert sb.get_deskewed_data_shape((101, 256, 256), 30, 0.39, False)[0] == (256, 256, 38)
    assert sb.get_deskewed_data_shape((101, 256, 256), 30, 0.39, True)[0] == (256, 256, 481)
    assert sb.get_deskewed_data_shape((600, 300, 2048), 30, 0.39, False, 3)[0] == (100, 2048, 1279)
    assert sb.get_deskewed_data_shape((4000, 300, 2048), 30, 0.39, True)[0] == (300, 2048, 10517)
    assert sb.get_deskewed_data_shape((2, 3, 4), 36, 0.386, True)[0] == (3, 4, 8)
    shape, vox = sb.get_deskewed_data_shape(raw_data_shape=(600, 300, 2048), ls_angle_deg=30, px_to_scan_ratio=0.39,
                                            keep_overhang=False, average_n_slices=3, pixel_size_um=0.116)
    assert isinstance(shape, tuple) and all(isinstance(v, int) for v in shape)
    assert vox == pytest.approx((3 * 0.5 * 0.116, 0.116, 0.116))


def test_geometry_equals_oracle_bitwise():
    rng = np.random.default_rng(0)
    for _ in range(500):
        shape = tuple(int(v) for v in rng.integers(1, 900, 3))
        th, r = round(float(rng.uniform(1, 45)), 2), round(float(rng.uniform(0.2, 1.2)), 3)
        keep, n = bool(rng.integers(0, 2)), int(rng.integers(1, 5))
        (a, b, c), vox = o.get_deskewed_data_shape(shape, th, r, keep, n, 0.116)
        got, gvox = sb.get_deskewed_data_shape(shape, th, r, keep, n, 0.116)
        assert got == (a, b, max(c, 0))
        assert tuple(float(v) for v in gvox) == tuple(float(v) for v in vox)
        assert np.array_equal(sb.deskew_geometry(shape, th, r, keep, n).matrix(),
                              o.deskew_affine_matrix(shape, th, r, keep))


def test_geometry_rejects_nonsense():
    with pytest.raises(ValueError):
        sb.get_deskewed_data_shape((10, 10), 30, 0.39, True)
    with pytest.raises(ValueError):
        sb.get_deskewed_data_shape((10, 0, 10), 30, 0.39, True)
    with pytest.raises(ValueError):
        sb.get_deskewed_data_shape((10, 10, 10), 30, 0.0, True)


# ---- b: the names and parameter names shrimPy relies on ---------------------------------------------
def test_signatures_carry_the_settings_field_names():
    fields = set(sb.DeskewSettings.model_fields)
    shape_params = set(inspect.signature(sb.get_deskewed_data_shape).parameters)
    fast_params = set(inspect.signature(sb.fast_deskew_zyx).parameters)
    data_params = set(inspect.signature(sb.deskew_data).parameters)
    assert {"raw_data_shape", "ls_angle_deg", "px_to_scan_ratio", "keep_overhang", "average_n_slices",
            "pixel_size_um"} <= shape_params
    assert {"raw_data", "ls_angle_deg", "px_to_scan_ratio", "keep_overhang", "average_n_slices"} <= fast_params
    assert {"raw_data", "ls_angle_deg", "px_to_scan_ratio", "keep_overhang", "average_n_slices", "device"} <= data_params
    # the reference filters model_dump() by signature: nothing the kernels need may be dropped,
    # and nothing the settings carry beyond that may leak into fast_deskew_zyx
    assert fields & fast_params == {"ls_angle_deg", "px_to_scan_ratio", "keep_overhang", "average_n_slices"}
    assert list(inspect.signature(sb.fast_deskew_zyx).parameters)[0] == "raw_data"


def test_shim_imports_resolve():
    sb.install_biahub_shim()
    from biahub.analysis.deskew import deskew_data, get_deskewed_data_shape as g2
    from biahub.deskew import fast_deskew_zyx, get_deskewed_data_shape
    from biahub.settings import DeskewSettings

    assert fast_deskew_zyx is sb.fast_deskew_zyx and get_deskewed_data_shape is sb.get_deskewed_data_shape
    assert deskew_data is sb.deskew_data and g2 is sb.get_deskewed_data_shape
    assert DeskewSettings is sb.DeskewSettings


def test_committed_reference_boundary_fixture():
    """What the UNMODIFIED reference passed across the boundary (tests/golden/make_golden.py)."""
    cases = json.loads((GOLDEN / "reference_boundary.json").read_text())
    assert len(cases) == 4
    for case in cases:
        s = sb.DeskewSettings(**case["deskew"])
        assert s.model_dump() == case["settings_dump"]
        accepted = set(inspect.signature(sb.get_deskewed_data_shape).parameters)
        kwargs = {k: v for k, v in s.model_dump().items() if k in accepted}
        assert kwargs == {k: v for k, v in case["shape_call_kwargs"].items() if k != "raw_data_shape"}
        shape, _ = sb.get_deskewed_data_shape(raw_data_shape=tuple(case["zyx"]), **kwargs)
        assert list(shape) == case["stored_zyx_shape"]
        accepted = set(inspect.signature(sb.fast_deskew_zyx).parameters)
        assert {k: v for k, v in s.model_dump().items() if k in accepted} == case["fast_deskew_kwargs"]
        assert case["log"] and str(tuple(case["stored_zyx_shape"])) in case["log"][0]


@pytest.mark.skipif(not Path("/root/reference/shrimpy/preprocessing.py").exists(),
                    reason="reference tree only exists in the authoring container")
def test_unmodified_reference_caller_against_the_shim(monkeypatch):
    """Reference caller, new callee: build_preprocessor -> DeskewSettings -> warm_up -> shape (CPU part)."""
    sb.install_biahub_shim()
    monkeypatch.syspath_prepend("/root/reference")
    for name in [m for m in sys.modules if m == "shrimpy" or m.startswith("shrimpy.")]:
        monkeypatch.delitem(sys.modules, name)
    from shrimpy import preprocessing as ref_pre

    import torch

    monkeypatch.setattr(ref_pre, "_resolve_device", lambda use_waveorder=True: torch.device("cpu"))
    pre = ref_pre.build_preprocessor((40, 20, 24), ["deskew"],
                                     deskew={"ls_angle_deg": 30.0, "keep_overhang": False, "average_n_slices": 3,
                                             "pixel_size_um": 0.1133, "scan_step_um": 0.174},
                                     output_channel="GFP")
    assert pre._zyx_shape == (7, 24, 45)
    assert type(pre._deskew_settings) is sb.DeskewSettings
