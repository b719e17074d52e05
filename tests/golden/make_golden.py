"""Regenerates the committed fixtures under ``tests/golden/``.  Run in the AUTHORING container:

    python tests/golden/make_golden.py

Four fixtures (``flatfield.npz`` and ``reductions.json`` hold outputs of unmodified reference functions --
``_LabelfreePreprocessor._flat_field_BF``, ``tracking._percentile``, ``tracking._intensity_center_of_mass`` --
i.e. results pinned by the reference itself):

``reference_boundary.json``
    Produced by importing the UNMODIFIED reference module ``/root/reference/shrimpy/preprocessing.py``
    and driving ``build_preprocessor`` / ``warm_up`` (``preprocessing.py:85-158, 209-252``) against the
    ``biahub``-named shim of this repo.  It records, for several ``deskew:`` dicts, the keyword
    arguments the reference passes across the boundary (after its ``inspect.signature`` filtering,
    ``preprocessing.py:44-56``), the deskewed shape it stores (``:244``) and the log line it emits
    (``:235-243``).  The arithmetic behind the shape is this repo's host code, so this fixture pins
    the *call contract*, not the voxel values.

``deskew_small.npz``
    Outputs of ``scipy.ndimage.affine_transform(order=1, mode="constant")`` + edge-padded mean (the
    normative scipy form, ``oracle/deskew_oracle.py``) on small seeded stacks, and of the affine
    resample.  No reference test holds a deskewed voxel (``shrimpy/tests/test_preprocessing.py:12-13``),
    so these vectors come from scipy itself (scipy 1.18.1, numpy 2.3.5): parity stays "unpinned"
    with respect to biahub, and pinned with respect to the scipy form the north star names.
"""

import json
import logging
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from helpers import synthetic_stack  # noqa: E402
from oracle import deskew_oracle as o  # noqa: E402

BOUNDARY_CASES = [
    {"zyx": [40, 20, 24], "deskew": {"ls_angle_deg": 30.0, "keep_overhang": False, "average_n_slices": 3,
                                     "pixel_size_um": 0.1133, "scan_step_um": 0.174}},
    {"zyx": [101, 256, 256], "deskew": {"ls_angle_deg": 30.0, "keep_overhang": True, "average_n_slices": 1,
                                        "pixel_size_um": 0.39, "px_to_scan_ratio": 0.39}},
    {"zyx": [600, 300, 2048], "deskew": {"ls_angle_deg": 30.0, "keep_overhang": False, "average_n_slices": 3,
                                         "pixel_size_um": 0.116, "scan_step_um": 0.31}},
    {"zyx": [592, 256, 1600], "deskew": {"ls_angle_deg": 30, "keep_overhang": False, "average_n_slices": 3,
                                         "pixel_size_um": 0.116, "scan_step_um": 0.313}},
]

VOXEL_CASES = [
    # name, shape, dtype, angle, ratio, keep, n, cval
    ("tiny_keep_n1", (23, 10, 7), "uint16", 30.0, 0.39, True, 1, 0.0),
    ("tiny_crop_n3", (23, 10, 7), "uint16", 30.0, 0.39, False, 3, 0.0),
    ("odd_keep_n2", (40, 11, 5), "uint16", 36.0, 0.651, True, 2, 5.0),
    ("aligned_crop_n3", (64, 9, 64), "uint16", 30.0, 0.39, False, 3, 0.0),
    ("float_keep_n4", (50, 12, 72), "float32", 30.0, 0.374, True, 4, -1.0),
    ("steep_crop_n2", (31, 7, 130), "uint16", 45.0, 0.77, False, 2, 0.0),
]


def boundary_fixture():
    ref = Path("/root/reference")
    if not ref.exists():
        raise SystemExit("/root/reference is required to regenerate reference_boundary.json")
    import shrimpy_b200

    shrimpy_b200.install_biahub_shim()
    sys.path.insert(0, str(ref))
    import biahub.deskew as bd
    from shrimpy import preprocessing as ref_pre  # the unmodified reference module

    seen = {}
    real_shape = bd.get_deskewed_data_shape

    def spy(*args, **kwargs):
        seen["args"] = list(args)
        seen["kwargs"] = {k: (list(v) if isinstance(v, tuple) else v) for k, v in kwargs.items()}
        return real_shape(*args, **kwargs)

    spy.__signature__ = __import__("inspect").signature(real_shape)
    bd.get_deskewed_data_shape = spy

    class Capture(logging.Handler):
        def __init__(self):
            super().__init__()
            self.lines = []

        def emit(self, record):
            self.lines.append(record.getMessage())

    cap = Capture()
    ref_pre.logger.addHandler(cap)
    ref_pre.logger.setLevel(logging.INFO)
    ref_pre._resolve_device = lambda use_waveorder=True: __import__("torch").device("cpu")  # no waveorder here
    out = []
    for case in BOUNDARY_CASES:
        cap.lines.clear()
        pre = ref_pre.build_preprocessor(tuple(case["zyx"]), ["deskew"], deskew=case["deskew"], output_channel="GFP")
        fast_kwargs = ref_pre._settings_kwargs(bd.fast_deskew_zyx, pre._deskew_settings)
        out.append({
            "zyx": case["zyx"], "deskew": case["deskew"],
            "shape_call_kwargs": seen["kwargs"], "shape_call_args": seen["args"],
            "stored_zyx_shape": list(pre._zyx_shape),
            "fast_deskew_kwargs": fast_kwargs,
            "settings_dump": pre._deskew_settings.model_dump(),
            "log": [ln for ln in cap.lines if "deskew will reshape" in ln],
        })
    bd.get_deskewed_data_shape = real_shape
    (HERE / "reference_boundary.json").write_text(json.dumps(out, indent=1) + "\n")
    print("wrote reference_boundary.json", len(out), "cases")


def voxel_fixture():
    arrays = {}
    for name, shape, dtype, ang, r, keep, n, cval in VOXEL_CASES:
        raw = synthetic_stack(shape, seed=len(name) + sum(shape), dtype=np.dtype(dtype).type)
        arrays[f"{name}__raw"] = raw
        arrays[f"{name}__params"] = np.array([ang, r, float(keep), n, cval], dtype=np.float64)
        arrays[f"{name}__out"] = o.deskew_data(raw, ang, r, keep, n, cval=cval)
    rng = np.random.default_rng(3)
    vol = rng.standard_normal((9, 20, 24)).astype(np.float32)
    a, b, c = np.deg2rad([2.0, 1.0, 3.0])
    Rz = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rx = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    M = np.eye(4)
    M[:3, :3] = Rz @ Ry @ Rx @ np.diag([1.03, 0.97, 1.1])
    M[:3, 3] = [0.4, -1.2, 2.3]
    arrays["affine_general__vol"] = vol
    arrays["affine_general__matrix"] = M
    arrays["affine_general__out"] = o.apply_affine_transform(vol, M, (10, 22, 26))
    M90 = np.array([[1.0, 0, 0, 0.25], [0, 0, -1.288, 23.0], [0, 1.288, 0, -2.0], [0, 0, 0, 1]])
    arrays["affine_rot90__matrix"] = M90
    arrays["affine_rot90__out"] = o.apply_affine_transform(vol, M90, (9, 16, 18))
    np.savez_compressed(HERE / "deskew_small.npz", **arrays)
    print("wrote deskew_small.npz", sum(v.nbytes for v in arrays.values()), "bytes raw")


def flatfield_fixture():
    """Golden output of the UNMODIFIED reference method _LabelfreePreprocessor._flat_field_BF (pure torch, CPU)."""
    ref = Path("/root/reference")
    if not ref.exists():
        raise SystemExit("/root/reference is required to regenerate flatfield.npz")
    sys.path.insert(0, str(ref))
    import torch
    from shrimpy import preprocessing as ref_pre

    pre = ref_pre._LabelfreePreprocessor.__new__(ref_pre._LabelfreePreprocessor)   # the method uses no instance state
    rng = np.random.default_rng(3)
    arrays = {}
    for name, shape in (("even_z", (8, 6, 10)), ("odd_z", (9, 5, 12)), ("tall", (70, 4, 66))):
        vol = rng.integers(80, 600, shape).astype(np.float32)
        # a smooth illumination pattern on top, so that the correction is not the identity
        yy, xx = np.meshgrid(np.linspace(0.7, 1.3, shape[1]), np.linspace(0.8, 1.2, shape[2]), indexing="ij")
        vol = np.round(vol * (yy * xx)[None]).astype(np.float32)
        arrays[f"{name}__vol"] = vol.astype(np.uint16)
        arrays[f"{name}__out"] = pre._flat_field_BF(torch.as_tensor(vol, dtype=torch.float32)).numpy()
    np.savez_compressed(HERE / "flatfield.npz", **arrays)
    print("wrote flatfield.npz", sum(v.nbytes for v in arrays.values()), "bytes raw")


def reductions_fixture():
    """Golden outputs of the UNMODIFIED reference helpers _percentile and _intensity_center_of_mass
    (shrimpy/dynatrack/tracking.py:572-649; pure torch, run on CPU)."""
    ref = Path("/root/reference")
    if not ref.exists():
        raise SystemExit("/root/reference is required to regenerate reductions.json")
    sys.path.insert(0, str(ref))
    import torch
    from shrimpy.dynatrack import tracking as ref_tr

    cases = []
    for seed, shape in ((0, (12, 20, 33)), (1, (7, 64, 50)), (2, (30, 9, 130))):
        rng = np.random.default_rng(seed)
        vol = rng.gamma(2.0, 300.0, size=shape).astype(np.float32)
        vol[shape[0] // 2, shape[1] // 3, shape[2] // 4] += 50000.0      # a bright blob pulls the centroid
        t = torch.as_tensor(vol)
        case = {"seed": seed, "shape": list(shape), "percentiles": {}, "com": {}}
        for p in (1.0, 25.0, 50.0, 90.0, 99.5):
            case["percentiles"][str(p)] = ref_tr._percentile(t, p)
        for bg in (0.0, 400.0, 1e9):
            case["com"][str(bg)] = [float(v) for v in ref_tr._intensity_center_of_mass(t, background=bg)]
        cases.append(case)
    (HERE / "reductions.json").write_text(json.dumps(cases, indent=1) + "\n")
    print("wrote reductions.json", len(cases), "cases")


if __name__ == "__main__":
    voxel_fixture()
    boundary_fixture()
    flatfield_fixture()
    reductions_fixture()
