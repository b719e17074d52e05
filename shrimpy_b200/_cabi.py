"""ctypes binding of ``libshrimpy_b200.so`` (the C-ABI declared in ``include/shrimpy_b200.h``).

There is deliberately no fallback: if the library is missing or cannot be
loaded, every entry point raises.  Build it with ``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C shrimpy_b200/csrc``.
"""

from __future__ import annotations

import ctypes
import os
from pathlib import Path

LIB_DIR = Path(__file__).resolve().parent / "_lib"
LIB_PATH = LIB_DIR / "libshrimpy_b200.so"

OK, EINVAL, ECUDA, ENOGPU, ENOMEM = range(5)
U16, F32 = 0, 1
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_TMA, KERNEL_TMA_STAGED = 0, 1, 2, 3
KERNELS = {"auto": KERNEL_AUTO, "direct": KERNEL_DIRECT, "tma": KERNEL_TMA, "tma_staged": KERNEL_TMA_STAGED}

# every symbol include/shrimpy_b200.h declares; tests check the library exports all of them
EXPORTS = (
    "shrimpy_abi_version",
    "shrimpy_last_error",
    "shrimpy_deskew_geometry",
    "shrimpy_deskew_geometry_trig",
    "shrimpy_deskew_device",
    "shrimpy_deskew_window_needs",
    "shrimpy_deskew_window_device",
    "shrimpy_flatfield_pattern_device",
    "shrimpy_flatfield_scale_device",
    "shrimpy_flatfield_apply_device",
    "shrimpy_deskew_flatfield_device",
    "shrimpy_deskew_range_device",
    "shrimpy_affine_device",
    "shrimpy_affine_strided_device",
    "shrimpy_minmax_device",
    "shrimpy_hist256_device",
    "shrimpy_center_of_mass_device",
    "shrimpy_zmax_projection_device",
    "shrimpy_min_device",
    "shrimpy_pipeline_create",
    "shrimpy_pipeline_destroy",
    "shrimpy_deskew_host",
    "shrimpy_pipeline_stats",
    "shrimpy_pipeline_staged_bytes",
    "shrimpy_host_alloc",
    "shrimpy_host_free",
    "shrimpy_blosc_info",
    "shrimpy_blosc_decode",
    "shrimpy_blosc_encode_bound",
    "shrimpy_blosc_encode",
    "shrimpy_launch_count",
)


class Window(ctypes.Structure):
    """``shrimpy_window`` (include/shrimpy_b200.h)."""

    _fields_ = [
        ("p_begin", ctypes.c_int32), ("p_count", ctypes.c_int32),
        ("c_begin", ctypes.c_int32), ("c_count", ctypes.c_int32),
        ("y_origin", ctypes.c_int32), ("y_count", ctypes.c_int32),
        ("z_origin", ctypes.c_int32), ("z_count", ctypes.c_int32),
    ]


class ShrimpyB200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"shrimpy_b200 error {code}: {message}")
        self.code = code


_lib = None


def _declare(lib) -> None:
    c_int, c_i64, c_dbl, c_flt, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_float, ctypes.c_void_p
    lib.shrimpy_abi_version.restype = c_int
    lib.shrimpy_abi_version.argtypes = []
    lib.shrimpy_last_error.restype = ctypes.c_char_p
    lib.shrimpy_last_error.argtypes = []
    lib.shrimpy_launch_count.restype = c_i64
    lib.shrimpy_launch_count.argtypes = []
    lib.shrimpy_deskew_geometry.restype = c_int
    lib.shrimpy_deskew_geometry.argtypes = [c_int, c_int, c_int, c_dbl, c_dbl, c_int, c_int, c_dbl,
                                            ctypes.POINTER(c_i64), ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl)]
    lib.shrimpy_deskew_geometry_trig.restype = c_int
    lib.shrimpy_deskew_geometry_trig.argtypes = [c_int, c_int, c_int, c_dbl, c_dbl, c_dbl, c_int, c_int, c_dbl,
                                                 ctypes.POINTER(c_i64), ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl)]
    deskew_common = [c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl, c_flt,
                     c_i64, c_i64, c_i64, c_i64]
    lib.shrimpy_deskew_device.restype = c_int
    lib.shrimpy_deskew_device.argtypes = deskew_common + [c_int, c_vp]
    lib.shrimpy_deskew_window_device.restype = c_int
    lib.shrimpy_deskew_window_device.argtypes = deskew_common + [ctypes.POINTER(Window), c_int, c_vp]
    lib.shrimpy_deskew_window_needs.restype = c_int
    lib.shrimpy_deskew_window_needs.argtypes = [c_int, c_int, c_int, c_dbl, c_dbl, c_dbl, c_int, c_int, c_int, c_int,
                                                ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]
    lib.shrimpy_flatfield_pattern_device.restype = c_int
    lib.shrimpy_flatfield_pattern_device.argtypes = [c_vp, c_int, c_vp, c_int, c_int, c_int, c_i64, c_i64, c_vp]
    lib.shrimpy_flatfield_scale_device.restype = c_int
    lib.shrimpy_flatfield_scale_device.argtypes = [c_vp, c_i64, c_vp, c_vp, c_vp]
    lib.shrimpy_flatfield_apply_device.restype = c_int
    lib.shrimpy_flatfield_apply_device.argtypes = [c_vp, c_int, c_vp, c_vp, c_int, c_int, c_int, c_vp]
    lib.shrimpy_deskew_flatfield_device.restype = c_int
    lib.shrimpy_deskew_flatfield_device.argtypes = [c_vp, c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int,
                                                    c_dbl, c_dbl, c_dbl, c_flt, c_i64, c_i64,
                                                    ctypes.POINTER(Window), c_int, c_vp]
    lib.shrimpy_deskew_range_device.restype = c_int
    lib.shrimpy_deskew_range_device.argtypes = [c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int,
                                                c_dbl, c_dbl, c_dbl, c_flt, c_i64, c_i64, c_int, c_vp]
    lib.shrimpy_affine_device.restype = c_int
    lib.shrimpy_affine_device.argtypes = [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int,
                                          ctypes.POINTER(c_dbl), c_flt, c_int, c_vp]
    lib.shrimpy_affine_strided_device.restype = c_int
    lib.shrimpy_affine_strided_device.argtypes = [c_vp, c_vp, c_int, c_int, c_int, c_i64, c_i64, c_int, c_int, c_int,
                                                  ctypes.POINTER(c_dbl), c_flt, c_int, c_vp]
    lib.shrimpy_minmax_device.restype = c_int
    lib.shrimpy_minmax_device.argtypes = [c_vp, c_i64, c_vp, c_vp]
    lib.shrimpy_hist256_device.restype = c_int
    lib.shrimpy_hist256_device.argtypes = [c_vp, c_i64, c_flt, c_flt, c_vp, c_vp]
    lib.shrimpy_center_of_mass_device.restype = c_int
    lib.shrimpy_center_of_mass_device.argtypes = [c_vp, c_int, c_int, c_int, c_flt, c_vp, c_vp]
    lib.shrimpy_zmax_projection_device.restype = c_int
    lib.shrimpy_zmax_projection_device.argtypes = [c_vp, c_int, c_int, c_int, c_flt, c_vp, c_vp]
    lib.shrimpy_min_device.restype = c_int
    lib.shrimpy_min_device.argtypes = [c_vp, c_int, c_i64, c_vp, c_vp]
    lib.shrimpy_pipeline_create.restype = c_int
    lib.shrimpy_pipeline_create.argtypes = [c_int, ctypes.c_size_t, ctypes.POINTER(c_vp)]
    lib.shrimpy_pipeline_destroy.restype = None
    lib.shrimpy_pipeline_destroy.argtypes = [c_vp]
    lib.shrimpy_deskew_host.restype = c_int
    lib.shrimpy_deskew_host.argtypes = [c_vp, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int,
                                        c_dbl, c_dbl, c_dbl, c_flt]
    lib.shrimpy_pipeline_stats.restype = c_int
    lib.shrimpy_pipeline_stats.argtypes = [c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]
    lib.shrimpy_pipeline_staged_bytes.restype = c_int
    lib.shrimpy_pipeline_staged_bytes.argtypes = [c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]
    lib.shrimpy_host_alloc.restype = c_int
    lib.shrimpy_host_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(c_vp)]
    lib.shrimpy_host_free.restype = c_int
    lib.shrimpy_host_free.argtypes = [c_vp]
    c_sz, c_i32 = ctypes.c_size_t, ctypes.c_int32
    lib.shrimpy_blosc_info.restype = c_int
    lib.shrimpy_blosc_info.argtypes = [c_vp, c_sz, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(c_i32),
                                       ctypes.POINTER(c_i32), ctypes.POINTER(c_i32)]
    lib.shrimpy_blosc_decode.restype = c_int
    lib.shrimpy_blosc_decode.argtypes = [c_vp, c_sz, c_vp, c_sz, c_int]
    lib.shrimpy_blosc_encode_bound.restype = c_sz
    lib.shrimpy_blosc_encode_bound.argtypes = [c_sz, c_i32, c_int]
    lib.shrimpy_blosc_encode.restype = c_int
    lib.shrimpy_blosc_encode.argtypes = [c_vp, c_sz, c_int, c_int, c_int, c_int, c_i32, c_int, c_vp, c_sz,
                                         ctypes.POINTER(c_sz)]


def lib():
    """The loaded library; raises ``ImportError`` with build instructions when it is absent."""
    global _lib
    if _lib is None:
        path = Path(os.environ.get("SHRIMPY_B200_LIB", LIB_PATH))
        if not path.exists():
            raise ImportError(
                f"{path} not found: the CUDA library is not built. Run `make -C shrimpy_b200/csrc` "
                "(needs nvcc with sm_100a support). shrimpy_b200 has no CPU or PyTorch fallback."
            )
        loaded = ctypes.CDLL(str(path))
        _declare(loaded)
        if loaded.shrimpy_abi_version() != 1:
            raise ImportError(f"{path}: ABI version {loaded.shrimpy_abi_version()} != 1; rebuild the library")
        _lib = loaded
    return _lib


def check(code: int) -> None:
    if code != OK:
        raise ShrimpyB200Error(code, lib().shrimpy_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(lib().shrimpy_launch_count())
