"""Synthetic go-diagram generator for the BASELINE.json configs 3/4/5 (numpy only).

Recipe follows SURVEY.md section 8(d): white background, 19x19 grid of 1-px black lines with
spacing `s`, centred; each intersection occupied with probability 0.45, colour uniform
{black, white}; black = filled anti-aliased disc, white = white disc with a 2-px black
anti-aliased outline; optional Gaussian pixel noise; `numpy.random.default_rng(seed)`.
The drawing is done with a distance-field coverage ramp instead of cv2 drawing calls so the
bench has no OpenCV dependency; inputs are only ever compared implementation-vs-checker on the
SAME array, so the exact rasteriser is irrelevant to parity.
"""
from __future__ import annotations

import numpy as np

BOARD = 19

CONFIGS = {
    # name: (size, spacing, radius, pinned line threshold)
    "synth2048": (2048, 60, 28, 176),   # BASELINE.json configs[2]
    "synth1024": (1024, 50, 24, 150),   # BASELINE.json configs[3]
}


def _disc(img, cx, cy, r, colour, outline=0):
    """Anti-aliased filled disc (and optional dark outline of `outline` px) blended into img."""
    h, w = img.shape
    R = int(r + outline + 2)
    x0, x1 = max(0, cx - R), min(w, cx + R + 1)
    y0, y1 = max(0, cy - R), min(h, cy + R + 1)
    yy, xx = np.mgrid[y0:y1, x0:x1]
    d = np.sqrt((xx - cx) ** 2 + (yy - cy) ** 2)
    win = img[y0:y1, x0:x1].astype(np.float32)
    if outline:
        cov = np.clip(r + 0.5 - d, 0, 1)                 # outer edge of the ring
        win = win * (1 - cov) + 0.0 * cov
        cov_in = np.clip(r - outline + 0.5 - d, 0, 1)    # inner white fill
        win = win * (1 - cov_in) + float(colour) * cov_in
    else:
        cov = np.clip(r + 0.5 - d, 0, 1)
        win = win * (1 - cov) + float(colour) * cov
    img[y0:y1, x0:x1] = np.clip(np.rint(win), 0, 255).astype(np.uint8)


# a tiny 3x5 digit font for the "numbered stones" set (config 5)
_FONT = {
    "0": "111101101101111", "1": "010110010010111", "2": "111001111100111", "3": "111001111001111",
    "4": "101101111001001", "5": "111100111001111", "6": "111100111101111", "7": "111001001001001",
    "8": "111101111101111", "9": "111101111001111",
}


def _number(img, cx, cy, text, scale, colour):
    cw, ch = 3 * scale, 5 * scale
    total = len(text) * (cw + scale) - scale
    x = cx - total // 2
    y = cy - ch // 2
    for chr_ in text:
        bits = _FONT[chr_]
        for r in range(5):
            for c in range(3):
                if bits[r * 3 + c] == "1":
                    ya, xa = y + r * scale, x + c * scale
                    img[max(0, ya):max(0, ya + scale), max(0, xa):max(0, xa + scale)] = colour
        x += cw + scale


def diagram(size: int, spacing: int, radius: int, seed: int, noise: float = 0.0,
            numbered: bool = False, fill: float = 0.45):
    """Return (grey u8 [size,size], truth int8 [19,19] with 0 empty / 1 black / 2 white).

    truth[i, j]: i = column (x index), j = row (y index) -- same convention as the
    reference's detected_board (img2sgf.py:502-505).
    """
    rng = np.random.default_rng(seed)
    img = np.full((size, size), 255, np.uint8)
    span = spacing * (BOARD - 1)
    o = (size - span) // 2
    for k in range(BOARD):
        p = o + k * spacing
        img[p, o:o + span + 1] = 0
        img[o:o + span + 1, p] = 0
    occ = rng.random((BOARD, BOARD)) < fill
    col = rng.integers(1, 3, (BOARD, BOARD))
    truth = np.where(occ, col, 0).astype(np.int8)
    num = 1
    for i in range(BOARD):
        for j in range(BOARD):
            if not truth[i, j]:
                continue
            cx, cy = o + i * spacing, o + j * spacing
            if truth[i, j] == 1:
                _disc(img, cx, cy, radius, 0)
            else:
                _disc(img, cx, cy, radius, 255, outline=2)
            if numbered:
                _number(img, cx, cy, str(num), max(1, spacing // 24), 255 if truth[i, j] == 1 else 0)
                num += 1
    if noise > 0:
        img = np.clip(np.rint(img.astype(np.float32) + rng.normal(0, noise, img.shape)), 0, 255).astype(np.uint8)
    return img, truth


def batch(config: str, start: int, count: int, noise: float = 0.0, numbered: bool = False):
    """[count, size, size] u8 greyscale diagrams with seeds start..start+count-1, plus truths."""
    size, s, r, _ = CONFIGS[config]
    imgs = np.empty((count, size, size), np.uint8)
    truths = np.empty((count, BOARD, BOARD), np.int8)
    for k in range(count):
        imgs[k], truths[k] = diagram(size, s, r, start + k, noise, numbered)
    return imgs, truths


def to_rgb(grey: np.ndarray) -> np.ndarray:
    """Replicate to 3 channels (the reference's Canny runs on the colour array, img2sgf.py:162)."""
    return np.ascontiguousarray(np.repeat(grey[..., None], 3, axis=-1))
