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


# ---------------------------------------------------------------------------------------------------
# Arithmetic identities two other kernels of the circle detector rely on, restated with numpy.

def test_edge_gradient_from_aligned_words():
    """k_edge_list (circles.cu): for an interior pixel the three columns of a row come out of the one or two
    aligned 32-bit words that hold them by a funnel shift, and the Sobel sums are byte dot products with the
    signed weights (-1,0,1), (-2,0,2), (1,2,1), -(1,2,1): equal to the plain 3x3 Sobel for every alignment."""
    rng = np.random.default_rng(5)
    w = 64
    rows = rng.integers(0, 256, (3, w), dtype=np.uint8)
    words = rows.view("<u4")                                            # aligned words of each row

    def s8(b):
        return b - 256 if b >= 128 else b

    def dp4a_us(a, wts, c):
        return c + sum(((a >> (8 * i)) & 0xff) * s8((wts >> (8 * i)) & 0xff) for i in range(4))

    for px in range(1, w - 1):
        a, b, sh = ((px - 1) & ~3) // 4, ((px + 1) & ~3) // 4, 8 * ((px - 1) & 3)
        r = []
        for k in range(3):
            lo = int(words[k, a]); hi = int(words[k, b]) if sh >= 16 else 0   # second word only when straddling
            r.append(((hi << 32 | lo) >> sh) & 0xffffffff)
        dx = dp4a_us(r[0], 0x000100FF, dp4a_us(r[1], 0x000200FE, dp4a_us(r[2], 0x000100FF, 0)))
        dy = dp4a_us(r[2], 0x00010201, dp4a_us(r[0], 0x00FFFEFF, 0))
        p = rows.astype(np.int64)
        want_dx = (p[0, px + 1] + 2 * p[1, px + 1] + p[2, px + 1]) - (p[0, px - 1] + 2 * p[1, px - 1] + p[2, px - 1])
        want_dy = (p[2, px - 1] + 2 * p[2, px] + p[2, px + 1]) - (p[0, px - 1] + 2 * p[0, px] + p[0, px + 1])
        assert (dx, dy) == (want_dx, want_dy), px


def test_radius_scan_on_packed_prefix():
    """k_radius: the histogram is overwritten in place by (inclusive prefix sum << 16) | (1 + highest non-empty
    bin at or below); OpenCV's scan from the top bin (every non-zero bin opens a 10-bin window, the bin just
    below the window is skipped -- oracle/img2sgf_oracle.c, radius estimation) then needs three words per
    window.  Same windows and counts as the plain scan, on random histograms."""
    rng = np.random.default_rng(9)
    NB = 290
    for trial in range(200):
        bins = np.zeros(NB, np.int64)
        k = rng.integers(0, 60)
        bins[rng.integers(0, NB, k)] += rng.integers(1, 40, k)
        # plain scan (the oracle's loop)
        want, j = [], NB - 1
        while j > 0:
            if bins[j]:
                up, cur = j, 0
                while j > up - 10 and j >= 0:
                    cur += bins[j]; j -= 1
                want.append((up, j, int(cur)))
            j -= 1
        # packed form
        pref = np.cumsum(bins)
        below, P = 0, np.zeros(NB, np.int64)
        for b in range(NB):
            if bins[b]:
                below = b + 1
            P[b] = (pref[b] << 16) | below
        got, j = [], NB - 1
        while j > 0:
            up = int(P[j] & 0xffff) - 1
            if up <= 0:
                break
            jn, cur = up - 10, int(P[up] >> 16)
            if jn >= 0:
                cur -= int(P[jn] >> 16)
            else:
                jn = -1
            got.append((up, jn, cur))
            j = jn - 1
        assert got == want, trial
