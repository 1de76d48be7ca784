"""ctypes binding of the C restatement (`oracle/img2sgf_oracle.c`).  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this.
The product package never does.  Function names mirror the reference call sites they restate
(see the C file header for file:line citations).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "img2sgf_oracle.c")
_SO = os.path.join(_HERE, "_build", "libimg2sgf_oracle.so")

HORIZONTAL, VERTICAL = 1, 2
BOARD_SIZE = 19


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", _SO, _SRC, "-lm"])
    return _SO


class Grid(C.Structure):
    _fields_ = [("valid", C.c_int), ("hsize", C.c_int), ("vsize", C.c_int),
                ("hspace", C.c_double), ("vspace", C.c_double),
                ("hc", C.c_double * 64), ("vc", C.c_double * 64)]


class Result(C.Structure):
    _fields_ = [("n_circles", C.c_int), ("n_hlines", C.c_int), ("n_vlines", C.c_int),
                ("n_hcentres", C.c_int), ("n_vcentres", C.c_int), ("board_ready", C.c_int),
                ("n_black", C.c_int), ("n_white", C.c_int), ("grid", Grid),
                ("board", C.c_uint8 * (BOARD_SIZE * BOARD_SIZE))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.o_average_intensity.restype = C.c_double
    return _lib


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.c_void_p)


def grey(rgb):
    rgb, p = _u8(rgb)
    h, w = rgb.shape[:2]
    out = np.empty((h, w), np.uint8)
    lib().o_grey(p, h, w, out.ctypes.data_as(C.c_void_p))
    return out


def contrast(rgb, factor):
    rgb, p = _u8(rgb)
    h, w = rgb.shape[:2]
    out = np.empty_like(rgb)
    lib().o_contrast(p, h, w, C.c_double(factor), out.ctypes.data_as(C.c_void_p))
    return out


def enhance(rgb, contrast_factor=1.0, brightness_factor=1.0):
    rgb, p = _u8(rgb)
    h, w = rgb.shape[:2]
    out = np.empty_like(rgb)
    lib().o_enhance(p, h, w, C.c_double(contrast_factor), C.c_double(brightness_factor), out.ctypes.data_as(C.c_void_p))
    return out


def _unary(fn, img, *args):
    img, p = _u8(img)
    h, w = img.shape[:2]
    out = np.empty((h, w), np.uint8)
    getattr(lib(), fn)(p, h, w, *args, out.ctypes.data_as(C.c_void_p))
    return out


def gauss(img, b):
    return _unary("o_gauss", img, int(b))


def median(img, b):
    return _unary("o_median", img, int(b))


def canny_rgb(rgb, low=50, high=200):
    return _unary("o_canny_rgb", rgb, int(low), int(high))


def canny_grey(img, low=50, high=100):
    return _unary("o_canny_grey", img, int(low), int(high))


def hough_circles(img, cap=65536, taps=False):
    img, p = _u8(img)
    h, w = img.shape
    out = np.empty((cap, 3), np.float32)
    if taps:
        edges = np.empty((h, w), np.uint8)
        acc = np.empty((h + 2, w + 2), np.int32)
        n = lib().o_hough_circles_ex(p, h, w, out.ctypes.data_as(C.c_void_p), cap,
                                     edges.ctypes.data_as(C.c_void_p), acc.ctypes.data_as(C.c_void_p))
        assert n <= cap
        return out[:n].copy(), edges, acc
    n = lib().o_hough_circles(p, h, w, out.ctypes.data_as(C.c_void_p), cap)
    assert n <= cap
    return out[:n].copy()


def mask_circles(edges, circles):
    out = np.ascontiguousarray(edges, dtype=np.uint8).copy()
    c = np.ascontiguousarray(circles, dtype=np.float32).reshape(-1, 3)
    h, w = out.shape
    lib().o_mask(out.ctypes.data_as(C.c_void_p), h, w, c.ctypes.data_as(C.c_void_p), len(c))
    return out


def hough_lines(img, thr, min_theta, max_theta, cap=1 << 16):
    img, p = _u8(img)
    h, w = img.shape
    rho = np.empty(cap, np.float32)
    theta = np.empty(cap, np.float32)
    n = lib().o_hough_lines(p, h, w, int(thr), C.c_double(min_theta), C.c_double(max_theta),
                            rho.ctypes.data_as(C.c_void_p), theta.ctypes.data_as(C.c_void_p), cap)
    assert n <= cap
    return rho[:n].copy(), theta[:n].copy()


def find_lines(img, thr, direction, cap=1 << 16):
    img, p = _u8(img)
    h, w = img.shape
    rho = np.empty(cap, np.float32)
    n = lib().o_find_lines(p, h, w, int(thr), int(direction), rho.ctypes.data_as(C.c_void_p), cap)
    assert n <= cap
    return rho[:n].reshape(-1, 1).copy()


def cluster(lines):
    r = np.ascontiguousarray(np.asarray(lines, dtype=np.float32).reshape(-1))
    out = np.empty(max(len(r), 1), np.float64)
    k = lib().o_cluster(r.ctypes.data_as(C.c_void_p), len(r), out.ctypes.data_as(C.c_void_p))
    return out[:k].copy()


def complete_grid(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(1100, np.float64)
    n = lib().o_complete_grid(x.ctypes.data_as(C.c_void_p), len(x), out.ctypes.data_as(C.c_void_p))
    return None if n < 0 else out[:n].copy()


def validate_grid(hcentres, vcentres) -> Grid:
    hc = np.ascontiguousarray(hcentres, dtype=np.float64)
    vc = np.ascontiguousarray(vcentres, dtype=np.float64)
    g = Grid()
    lib().o_validate_grid(hc.ctypes.data_as(C.c_void_p), len(hc), vc.ctypes.data_as(C.c_void_p), len(vc),
                          C.byref(g))
    return g


def classify(grey_u8, circles, g: Grid, black_thr=128):
    grey_u8, p = _u8(grey_u8)
    h, w = grey_u8.shape
    c = np.ascontiguousarray(circles, dtype=np.float32).reshape(-1, 3)
    board = np.zeros(BOARD_SIZE * BOARD_SIZE, np.uint8)
    br = np.zeros(BOARD_SIZE * BOARD_SIZE, np.float64)
    nb, nw = C.c_int(), C.c_int()
    k = lib().o_classify(p, h, w, c.ctypes.data_as(C.c_void_p), len(c), C.byref(g), int(black_thr),
                         board.ctypes.data_as(C.c_void_p), br.ctypes.data_as(C.c_void_p),
                         C.byref(nb), C.byref(nw))
    return board[:g.hsize * g.vsize].reshape(g.hsize, g.vsize).copy(), br[:k].copy()


def choose_threshold(width, height):
    """img2sgf.py:606-613"""
    t = int(min(width, height) / 12.8 + 16)
    return int(min(max(t, 20), 200))


def pipeline(rgb, line_thr=None, black_thr=128, circ_cap=1 << 16):
    """Whole path for one contrast-enhanced RGB image.  Returns (Result, circles, masked)."""
    rgb, p = _u8(rgb)
    h, w = rgb.shape[:2]
    if line_thr is None:
        line_thr = choose_threshold(w, h)
    circles = np.empty((circ_cap, 3), np.float32)
    masked = np.empty((h, w), np.uint8)
    res = Result()
    rc = lib().o_pipeline(p, h, w, int(line_thr), int(black_thr), circles.ctypes.data_as(C.c_void_p),
                          circ_cap, masked.ctypes.data_as(C.c_void_p), C.byref(res))
    if rc != 0:
        raise RuntimeError("oracle circle capacity exceeded")
    return res, circles[:res.n_circles].copy(), masked


def board_of(res: Result):
    g = res.grid
    if not res.board_ready:
        return None
    return np.frombuffer(bytes(res.board), np.uint8)[:g.hsize * g.vsize].reshape(g.hsize, g.vsize).copy()
