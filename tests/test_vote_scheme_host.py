"""The tile / clipped-ray / guard-band scheme of the Hough-circle vote kernel, restated on the CPU
(tests/host/vote_host.cpp mirrors k_vote_peaks2), against the reference accumulator of the oracle:
the set of accumulator peaks must be identical, no vote may ever leave the shared tile, and the
result must not depend on how conservative the clipping interval is."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "vote_host.cpp")
SO = os.path.join(HERE, "host", "_build", "libvote_host.so")


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", SO, SRC])
    return C.CDLL(SO)


def _oracle_peaks(img):
    _, edges, acc = O.hough_circles(img, taps=True)
    v = acc
    c = v[1:-1, 1:-1]
    pk = (c > 30) & (c > v[1:-1, :-2]) & (c >= v[1:-1, 2:]) & (c > v[:-2, 1:-1]) & (c >= v[2:, 1:-1])
    ys, xs = np.nonzero(pk)
    return edges, np.sort((ys + 1) * v.shape[1] + (xs + 1)), int(acc.sum())


def _host_peaks(host, img, edges, slack):
    h, w = img.shape
    out = np.zeros(1 << 16, np.int32)
    cast = C.c_longlong(0)
    n = host.vh_vote_peaks(img.ctypes.data_as(C.c_void_p), edges.ctypes.data_as(C.c_void_p), h, w, C.c_float(slack),
                           out.ctypes.data_as(C.c_void_p), len(out), C.byref(cast))
    assert n >= 0, "a vote left the shared tile: guard band too small"
    assert n <= len(out)
    return np.sort(out[:n].astype(np.int64)), cast.value


@pytest.mark.parametrize("size,spacing,radius,seed", [(300, 16, 7, 1), (257, 12, 5, 2), (450, 24, 11, 3), (230, 12, 5, 4)])
def test_tile_scheme_matches_reference_accumulator(host, size, spacing, radius, seed):
    from img2sgf_b200 import synth
    grey, _ = synth.diagram(size, spacing, radius, seed=seed, noise=1.5 if seed % 2 else 0.0)
    for img in (np.ascontiguousarray(grey), O.gauss(grey, 5), np.ascontiguousarray(grey[:, : size - 37])):
        edges, want, total = _oracle_peaks(img)
        got, cast = _host_peaks(host, img, edges, 0.25)
        assert np.array_equal(got, want), (img.shape, len(got), len(want))
        wide, cast_wide = _host_peaks(host, img, edges, 1.0)          # a more conservative interval changes nothing
        assert np.array_equal(wide, want)
        assert cast >= total and cast_wide >= cast                     # extra votes only ever land in guard cells


def test_tile_scheme_random_noise(host):
    rng = np.random.default_rng(9)
    img = (rng.integers(0, 2, (140, 200)) * 255).astype(np.uint8)
    img = O.gauss(O.gauss(img, 7), 7)
    edges, want, _ = _oracle_peaks(img)
    got, _ = _host_peaks(host, img, edges, 0.25)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("b", [3, 5, 7])
def test_saturated_window_shortcut_is_exact(b):
    """The median kernels' shortcut: a value held by more than half of the b x b window (REPLICATE
    border) is the median.  Checked against the oracle's medianBlur on diagram-like content."""
    from img2sgf_b200 import synth
    grey, _ = synth.diagram(300, 16, 7, seed=11, noise=0.0)
    noisy, _ = synth.diagram(300, 16, 7, seed=12, noise=3.0)
    for img in (grey, 255 - grey, noisy):
        img = np.ascontiguousarray(img)
        med = O.median(img, b)
        r = b // 2
        pad = np.pad(img, r, mode="edge")
        win = np.lib.stride_tricks.sliding_window_view(pad, (b, b))
        km = (b * b) // 2 + 1
        for value in (255, 0):
            dominated = (win == value).sum(axis=(2, 3)) >= km
            assert dominated.any()
            assert (med[dominated] == value).all()
