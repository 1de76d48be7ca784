"""ctypes binding of include/img2sgf_b200.h.  No fallback: if the CUDA library is missing or
cannot be loaded this module raises, it never routes work to a CPU path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "_lib", "libimg2sgf_b200.so")

BOARD_SIZE = 19
MAX_GRID = 32
N_CALLS = 10
N_UNIQUE = 8

ST_CAND_OVERFLOW = 1
ST_CIRCLE_OVERFLOW = 2
ST_LINE_OVERFLOW = 4
ST_HYST_NOT_CONVERGED = 8
ST_GRID_OVERFLOW = 16


class Limits(C.Structure):
    _fields_ = [("cand_cap", C.c_int32), ("circle_cap", C.c_int32), ("line_cap", C.c_int32),
                ("hyst_passes", C.c_int32)]


class Batch(C.Structure):           # i2s_batch_t
    _fields_ = [("n", C.c_int32), ("channels", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("pitch", C.c_int32), ("pad_", C.c_int32), ("images", C.c_void_p)]


class Params(C.Structure):          # i2s_params_t
    _fields_ = [("line_threshold", C.c_int32), ("black_threshold", C.c_int32), ("canny_low", C.c_int32),
                ("canny_high", C.c_int32), ("contrast_factor", C.c_float), ("brightness_factor", C.c_float)]


class Taps(C.Structure):            # i2s_taps_t
    _fields_ = [("plane_pitch", C.c_int32), ("pad_", C.c_int32), ("grey", C.c_void_p), ("edges", C.c_void_p),
                ("masked", C.c_void_p), ("circles", C.c_void_p), ("counts", C.c_void_p), ("rho", C.c_void_p),
                ("line_counts", C.c_void_p), ("grids", C.c_void_p), ("brightness", C.c_void_p)]


IMAGE_DTYPE = np.dtype([("offset", "<i8"), ("h", "<i4"), ("w", "<i4"), ("pitch", "<i4"), ("line_threshold", "<i4")])
GRID_DTYPE = np.dtype([("valid", "<i4"), ("hsize", "<i4"), ("vsize", "<i4"), ("pad_", "<i4"),
                       ("hspace", "<f8"), ("vspace", "<f8"),
                       ("hcentres", "<f8", (MAX_GRID,)), ("vcentres", "<f8", (MAX_GRID,))])
RECORD_DTYPE = np.dtype([("board", "u1", (BOARD_SIZE * BOARD_SIZE,)), ("valid", "u1"), ("board_ready", "u1"),
                         ("hsize", "u1"), ("vsize", "u1"), ("pad_", "u1", (3,)),
                         ("n_black", "<i4"), ("n_white", "<i4"), ("n_circles", "<i4"), ("status", "<i4")])
assert RECORD_DTYPE.itemsize == 384
assert IMAGE_DTYPE.itemsize == 24
assert GRID_DTYPE.itemsize == 32 + 2 * 8 * MAX_GRID

_P = C.c_void_p
_I = C.c_int
_SIGNATURES = {
    # name: (restype, argtypes)
    "i2s_last_error": (C.c_char_p, []),
    "i2s_version": (_I, []),
    "i2s_default_limits": (None, [C.POINTER(Limits)]),
    "i2s_default_params": (None, [C.POINTER(Params)]),
    "i2s_canvas_pitch": (_I, [_I]),
    "i2s_grey": (_I, [_P, _I, _P, _I, _I, _I, _I, _P]),
    "i2s_enhance": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, C.c_double, C.c_double, _P]),
    "i2s_gauss357": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "i2s_median": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "i2s_canny_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "i2s_canny": (_I, [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, C.c_size_t, _P]),
    "i2s_hough_circles_workspace_bytes": (C.c_size_t, [_I, _I, _I, C.POINTER(Limits)]),
    "i2s_hough_circles": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, C.POINTER(Limits), _P, C.c_size_t, _P]),
    "i2s_mask_circles": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _I, _P]),
    "i2s_find_circles_workspace_bytes": (C.c_size_t, [_I, _I, _I, C.POINTER(Limits)]),
    "i2s_find_circles": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, C.POINTER(Limits), _P, C.c_size_t, _P]),
    "i2s_find_lines_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "i2s_find_lines": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, C.c_size_t, _P]),
    "i2s_cluster": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "i2s_validate_grid": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "i2s_classify_stones": (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _P, _I, _P, _P, _P]),
    "i2s_profile_enable": (_I, [_I]),
    "i2s_profile_section_name": (C.c_char_p, [_I]),
    "i2s_profile_read": (_I, [_P, _P, _I]),
    "i2s_launch_count": (C.c_longlong, [_I]),
    "i2s_pipeline_workspace_bytes": (C.c_size_t, [_I, _I, _I, C.POINTER(Limits)]),
    "i2s_pipeline": (_I, [_P, C.POINTER(Batch), C.POINTER(Params), _P, C.POINTER(Taps), C.POINTER(Limits), _P,
                          C.c_size_t, _P]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    """Load the CUDA library (built by `python -m img2sgf_b200.build`).  Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise NativeError(f"{SO_PATH} is missing: run `python -m img2sgf_b200.build` "
                              "(there is no CPU fallback)")
        l = C.CDLL(SO_PATH)
        for name, (res, args) in _SIGNATURES.items():
            f = getattr(l, name)        # AttributeError if the .so does not export a declared symbol
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise NativeError(f"{what} failed ({rc}): {lib().i2s_last_error().decode()}")


def default_limits() -> Limits:
    lim = Limits()
    lib().i2s_default_limits(C.byref(lim))
    return lim


def default_params() -> Params:
    p = Params()
    lib().i2s_default_params(C.byref(p))
    return p


def canvas_pitch(w: int) -> int:
    """Row pitch (bytes) of the library's own planes for images w pixels wide (multiple of 128)."""
    return (int(w) + 127) // 128 * 128


def describe_status(st: int) -> str:
    names = [(ST_CAND_OVERFLOW, "candidate capacity exceeded"), (ST_CIRCLE_OVERFLOW, "circle capacity exceeded"),
             (ST_LINE_OVERFLOW, "line capacity exceeded"), (ST_HYST_NOT_CONVERGED, "hysteresis pass budget exceeded"),
             (ST_GRID_OVERFLOW, "more than 32 grid lines on an axis")]
    return ", ".join(n for b, n in names if st & b) or "ok"
