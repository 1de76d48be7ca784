"""CPU emulation of the register-rolling kernels (Gaussian 3/5/7, Sobel + NMS) against the oracle.

tests/host/roll_host.cpp runs img2sgf_b200/csrc/roll_cores.cuh -- the arithmetic the CUDA kernels
k_gauss357_roll / k_canny_roll execute -- lane by lane on the CPU.  Bit-exact comparison
(tolerance 0) with the C oracle (oracle/img2sgf_oracle.c, pinned against cv2)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "roll_host.cpp")
CORES = os.path.join(HERE, "..", "img2sgf_b200", "csrc", "roll_cores.cuh")
SO = os.path.join(HERE, "host", "_build", "libroll_host.so")


@pytest.fixture(scope="module")
def host():
    newest = max(os.path.getmtime(SRC), os.path.getmtime(CORES))
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO, SRC])
    return C.CDLL(SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _images(rng):
    out = []
    for (h, w) in [(1, 1), (2, 3), (5, 4), (7, 129), (64, 120), (65, 121), (70, 250), (131, 37), (33, 480), (300, 130), (257, 250)]:
        out.append(rng.integers(0, 256, (h, w), dtype=np.uint8))
    # smooth / tie-heavy content: quantised blobs and flat areas
    h, w = 150, 260
    yy, xx = np.mgrid[0:h, 0:w]
    out.append(((np.sin(xx / 9.0) + np.cos(yy / 7.0)) * 60 + 128).astype(np.uint8))
    out.append((((xx // 8 + yy // 8) % 2) * 255).astype(np.uint8))
    out.append(np.full((40, 50), 255, np.uint8))
    q = rng.integers(0, 4, (90, 140), dtype=np.uint8) * 64
    out.append(q)
    return out


def test_gauss357_cores_match_oracle(host):
    rng = np.random.default_rng(11)
    for img in _images(rng):
        img = np.ascontiguousarray(img)
        h, w = img.shape
        d = [np.zeros((h, w), np.uint8) for _ in range(3)]
        host.rh_gauss357(_p(img), h, w, _p(d[0]), _p(d[1]), _p(d[2]))
        for b, got in zip((3, 5, 7), d):
            assert np.array_equal(got, O.gauss(img, b)), (h, w, b)


def _canny_host(host, img, ch, low, high, always_diag=0):
    h, w = img.shape[:2]
    st = np.zeros((h, w), np.uint8)
    host.rh_sobel_nms(_p(img), ch, h, w, low, high, always_diag, _p(st))
    ed = np.zeros((h, w), np.uint8)
    host.rh_hysteresis(_p(st), h, w, _p(ed))
    return st, ed


@pytest.mark.parametrize("low,high", [(50, 100), (50, 200), (0, 10), (300, 900)])
def test_sobel_nms_cores_match_oracle_grey(host, low, high):
    rng = np.random.default_rng(5)
    for img in _images(rng):
        img = np.ascontiguousarray(img)
        st, ed = _canny_host(host, img, 1, low, high)
        st2, _ = _canny_host(host, img, 1, low, high, always_diag=1)
        assert np.array_equal(st, st2)
        assert np.array_equal(ed, O.canny_grey(img, low, high)), (img.shape, low, high)


def test_sobel_nms_cores_match_oracle_rgb(host):
    rng = np.random.default_rng(6)
    for g in _images(rng):
        h, w = g.shape
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        rgb[..., 1] = g                                  # one structured channel, two noisy ones
        if h * w > 2000:
            rgb[..., 0] = g                              # and a tie between channels 0 and 1
        rgb = np.ascontiguousarray(rgb)
        _, ed = _canny_host(host, rgb, 3, 50, 200)
        assert np.array_equal(ed, O.canny_rgb(rgb, 50, 200)), (h, w)


def test_cores_on_synthetic_diagram(host):
    from img2sgf_b200 import synth
    grey, _ = synth.diagram(400, 20, 9, seed=3)
    for b in (3, 5, 7):
        blurred = O.gauss(grey, b)
        _, ed = _canny_host(host, blurred, 1, 50, 100)
        assert np.array_equal(ed, O.canny_grey(blurred, 50, 100))
    rgb = np.ascontiguousarray(synth.to_rgb(grey))
    _, ed = _canny_host(host, rgb, 3, 50, 200)
    assert np.array_equal(ed, O.canny_rgb(rgb, 50, 200))
