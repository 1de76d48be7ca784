"""Headless replay of the reference's diagram-recognition path (TEST INFRASTRUCTURE ONLY).

This module is the *executable reference*: it issues the same third-party library calls
(cv2 / scikit-learn / numpy / Pillow) with the same argument values as
`/root/reference/img2sgf.py` Part 2, with Tk getters and module globals turned into
explicit arguments.  It exists to (1) generate the golden vectors under `tests/golden/`,
(2) pin the C restatement in `oracle/img2sgf_oracle.c`, and (3) serve as the CPU baseline
arm of `bench.py` (`--impl reference`, `cpu_baseline.kind == "reference"`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import it.
The product package `img2sgf_b200` never does.

Pinned library versions (the reference pins none, README.md:29-41): opencv-python-headless
4.13.0, scikit-learn 1.9.0, numpy 2.3, Pillow 12.2.

Reference sites followed (file:line into /root/reference/img2sgf.py):
  enhance()            :142-150   contrast/brightness prologue (PIL ImageEnhance)
  grey()               :153       cv.cvtColor(BGR2GRAY) on an RGB-ordered array
  edge_map()           :162-165   cv.Canny on the 3-channel array, 50/200, aperture 3, L1
  blur_pyramid()       :171-175   [grey, edges] + (median b, gaussian b) for b = 1,3,5,7
  hough_circles()      :180       cv.HoughCircles(.., HOUGH_GRADIENT, 1, 10, [], 100, 30, 1, 30)
  find_circles()       :169-198   ten HoughCircles calls stacked, then the masking loop
  find_lines()         :230-255   near-axis cv.HoughLines, V2 rho negated, V1 before V2
  cluster()            :268-292   single-linkage agglomerative clustering at distance 10
  complete_grid()      :335-397
  truncate_grid()      :400-417
  validate_grid()      :420-445
  closest_index()      :448-459
  average_intensity()  :468-481
  classify_stones()    :497-515, :537-542
  choose_threshold()   :606-613
"""
from __future__ import annotations

import math
from bisect import bisect_left
from dataclasses import dataclass, field

import numpy as np

BOARD_SIZE = 19                      # img2sgf.py:43
BLACK_STONE_THRESHOLD = 128          # :45
EDGE_MIN, EDGE_MAX = 50, 200         # :47-48
MAXBLUR = 3                          # :51
ANGLE_DELTA = math.pi / 180 * 1.0    # :52-53
MIN_GRID_SPACING = 10                # :54
BIG_SPACE_RATIO = 1.6                # :55
CONTRAST_DEFAULT, BRIGHTNESS_DEFAULT = 70, 50   # :56-57
EMPTY, BLACK, WHITE, STONE = range(4)           # :82-83
HORIZONTAL, VERTICAL = 1, 2                     # :74-80


def _cv():
    import cv2
    return cv2


def choose_threshold(width: int, height: int) -> int:
    t = int(min(width, height) / 12.8 + 16)
    return int(min(max(t, 20), 200))


def enhance(pil_rgb, contrast: int = CONTRAST_DEFAULT, brightness: int = BRIGHTNESS_DEFAULT):
    from PIL import ImageEnhance
    out = ImageEnhance.Contrast(pil_rgb).enhance(102 / (101 - contrast) - 1)
    out = ImageEnhance.Brightness(out).enhance(450 / (200 - brightness) - 2)
    return np.array(out)


def grey(rgb: np.ndarray) -> np.ndarray:
    cv = _cv()
    return cv.cvtColor(rgb, cv.COLOR_BGR2GRAY)


def edge_map(rgb: np.ndarray, low: int = EDGE_MIN, high: int = EDGE_MAX) -> np.ndarray:
    cv = _cv()
    return cv.Canny(rgb, low, high, apertureSize=3, L2gradient=False)


def blur_pyramid(grey_u8: np.ndarray, edges_u8: np.ndarray) -> list:
    cv = _cv()
    out = [grey_u8, edges_u8]
    for i in range(MAXBLUR + 1):
        b = 2 * i + 1
        out.append(cv.medianBlur(grey_u8, b))
        out.append(cv.GaussianBlur(grey_u8, (b, b), b))
    return out


def hough_circles(img_u8: np.ndarray) -> np.ndarray:
    """One reference HoughCircles call -> (n,3) float32, n may be 0."""
    cv = _cv()
    c = cv.HoughCircles(img_u8, cv.HOUGH_GRADIENT, 1, 10, np.array([]), 100, 30, 1, 30)
    if c is None or len(c) == 0:
        return np.zeros((0, 3), np.float32)
    return np.ascontiguousarray(c[0], dtype=np.float32)


def mask_circles(edges_u8: np.ndarray, circles: np.ndarray) -> np.ndarray:
    cv = _cv()
    out = edges_u8.copy()
    for i in range(len(circles)):
        xc, yc, r = circles[i, :]
        r = r + 2
        ul = (int(round(xc - r)), int(round(yc - r)))
        lr = (int(round(xc + r)), int(round(yc + r)))
        mid = (int(round(xc)), int(round(yc)))
        cv.rectangle(out, ul, lr, (0, 0, 0), -1)
        cv.circle(out, mid, 1, (255, 255, 255), -1)
    return out


def find_circles(grey_u8: np.ndarray, edges_u8: np.ndarray, return_per_call: bool = False):
    per_call = [hough_circles(b) for b in blur_pyramid(grey_u8, edges_u8)]
    nonempty = [c for c in per_call if len(c)]
    circles = np.vstack(nonempty) if nonempty else np.zeros((0, 3), np.float32)
    masked = mask_circles(edges_u8, circles)
    if return_per_call:
        return circles, masked, per_call
    return circles, masked


def hough_lines_raw(masked_u8: np.ndarray, threshold: int, min_theta: float, max_theta: float):
    cv = _cv()
    return cv.HoughLines(masked_u8, rho=1, theta=math.pi / 180.0, threshold=threshold,
                         min_theta=min_theta, max_theta=max_theta)


def find_lines(masked_u8: np.ndarray, threshold: int, direction: int):
    if direction == HORIZONTAL:
        lines = hough_lines_raw(masked_u8, threshold, math.pi / 2 - ANGLE_DELTA,
                                math.pi / 2 + ANGLE_DELTA)
    else:
        v1 = hough_lines_raw(masked_u8, threshold, 0, ANGLE_DELTA)
        v2 = hough_lines_raw(masked_u8, threshold, math.pi - ANGLE_DELTA, math.pi)
        if v2 is not None:
            v2[:, 0, 0] = -v2[:, 0, 0]
            v2[:, 0, 1] = v2[:, 0, 1] - math.pi
            lines = np.vstack((v1, v2)) if v1 is not None else v2
        else:
            lines = v1
    return [] if lines is None else lines[:, 0, 0].reshape(-1, 1)


def cluster(lines):
    """find_clusters_fixed_threshold + get_cluster_centres on one direction's rho column."""
    from sklearn.cluster import AgglomerativeClustering
    model = AgglomerativeClustering(n_clusters=None, linkage='single',
                                    distance_threshold=MIN_GRID_SPACING)
    try:
        model = model.fit(lines)
    except Exception:
        return []
    n = model.n_clusters_
    centres = np.zeros(n)
    for i in range(n):
        centres[i] = lines[model.labels_ == i].mean()
    centres.sort()
    return centres


def complete_grid(x):
    if x is None or len(x) == 0 or len(x) == 1:
        return None
    spaces = x[1:] - x[:-1]
    min_space = min(spaces)
    if min_space < MIN_GRID_SPACING:
        return None
    bound = min_space * BIG_SPACE_RATIO
    big_spaces = spaces[spaces > bound]
    if len(big_spaces) == 0:
        return x
    small_spaces = spaces[spaces <= bound]
    max_space = max(small_spaces)
    average_space = (min_space + max_space) / 2
    n = len(small_spaces)
    for s in big_spaces:
        n += int(round(s / average_space))
    if n > BOARD_SIZE + 2:
        return None
    n += 1
    if len(x) < n:
        answer = np.zeros(n)
        answer[0] = x[0]
        i, j = 1, 1
        for s in spaces:
            if s <= max_space:
                answer[i] = x[j]
                i += 1
                j += 1
            else:
                m = int(round(s / average_space))
                for k in range(m):
                    answer[i] = x[j - 1] + (k + 1) * s / m
                    i += 1
                j += 1
        return answer
    return x


def truncate_grid(x):
    if x is None:
        return None
    if len(x) == BOARD_SIZE + 2:
        return x[1:-1]
    if len(x) == BOARD_SIZE + 1:
        return x[:-1]
    return x


@dataclass
class Grid:
    valid: bool
    circles: object = None
    vsize: int = 0
    hsize: int = 0
    hcentres_complete: object = None
    vcentres_complete: object = None
    hspace: object = None
    vspace: object = None


def validate_grid(hcentres, vcentres, circles) -> Grid:
    hc = truncate_grid(complete_grid(truncate_grid(_as_grid(hcentres))))
    if hc is None:
        return Grid(False, circles)
    vc = truncate_grid(complete_grid(truncate_grid(_as_grid(vcentres))))
    if vc is None:
        return Grid(False, circles)
    vsize, hsize = len(hc), len(vc)
    hspace = (hc[-1] - hc[0]) / vsize
    vspace = (vc[-1] - vc[0]) / hsize
    lo = min(hspace, vspace) * 0.3
    hi = max(hspace, vspace) * 0.65
    kept = [c for c in circles if lo < c[2] < hi]
    return Grid(True, kept, vsize, hsize, hc, vc, hspace, vspace)


def _as_grid(x):
    # the reference passes either a sorted float64 array or [] (img2sgf.py:283-292)
    if x is None:
        return None
    return np.asarray(x, dtype=np.float64) if len(x) else x


def closest_index(a, x) -> int:
    i = bisect_left(x, a)
    if i == 0:
        return 0
    if i == len(x):
        return i - 1
    return i - 1 if a - x[i - 1] <= x[i] - a else i


def average_intensity(grey_u8, i, j, g: Grid):
    x = g.vcentres_complete[i]
    xmin, xmax = int(round(x - g.hspace / 2)), int(round(x + g.hspace / 2))
    y = g.hcentres_complete[j]
    ymin, ymax = int(round(y - g.vspace / 2)), int(round(y + g.vspace / 2))
    xmin = max(0, xmin)
    ymin = max(0, ymin)
    xmax = min(grey_u8.shape[1], xmax)
    ymax = min(grey_u8.shape[0], ymax)
    with np.errstate(all='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            return np.mean(grey_u8[ymin:ymax, xmin:xmax])


def classify_stones(grey_u8, g: Grid, black_stone_threshold=BLACK_STONE_THRESHOLD):
    board = np.zeros((g.hsize, g.vsize))
    for c in g.circles:
        board[closest_index(c[0], g.vcentres_complete),
              closest_index(c[1], g.hcentres_complete)] = STONE
    n = np.count_nonzero(board)
    brightness = np.zeros(n)
    k = 0
    for i in range(g.hsize):
        for j in range(g.vsize):
            if board[i, j] == STONE:
                v = average_intensity(grey_u8, i, j, g)
                brightness[k] = v
                k += 1
                board[i, j] = BLACK if v <= black_stone_threshold else WHITE
    return board, brightness


@dataclass
class Replay:
    """Every intermediate of one image through the path."""
    rgb: np.ndarray
    threshold: int
    grey: np.ndarray = None
    edges: np.ndarray = None
    blurs: list = field(default_factory=list)
    per_call_circles: list = field(default_factory=list)
    circles: np.ndarray = None
    masked: np.ndarray = None
    hlines: object = None
    vlines: object = None
    hcentres: object = None
    vcentres: object = None
    grid: Grid = None
    board: np.ndarray = None
    brightness: np.ndarray = None


def run(rgb: np.ndarray, threshold: int | None = None,
        black_stone_threshold: int = BLACK_STONE_THRESHOLD, classify: bool = True) -> Replay:
    """process_image() from 'contrast-enhanced RGB array exists' (:150) to detected_board (:542)."""
    h, w = rgb.shape[:2]
    if threshold is None:
        threshold = choose_threshold(w, h)
    r = Replay(rgb=rgb, threshold=threshold)
    r.grey = grey(rgb)
    r.edges = edge_map(rgb)
    r.blurs = blur_pyramid(r.grey, r.edges)
    r.circles, r.masked, r.per_call_circles = find_circles(r.grey, r.edges, True)
    r.hlines = find_lines(r.masked, threshold, HORIZONTAL)
    r.vlines = find_lines(r.masked, threshold, VERTICAL)
    # the reference re-runs find_lines inside the clustering (:269); deterministic, same result
    find_lines(r.masked, threshold, HORIZONTAL)
    find_lines(r.masked, threshold, VERTICAL)
    r.hcentres = cluster(r.hlines) if len(r.hlines) else []
    r.vcentres = cluster(r.vlines) if len(r.vlines) else []
    if not classify:
        return r
    r.grid = validate_grid(r.hcentres, r.vcentres, r.circles)
    if r.grid.valid and r.grid.hsize <= BOARD_SIZE and r.grid.vsize <= BOARD_SIZE:
        r.board, r.brightness = classify_stones(r.grey, r.grid, black_stone_threshold)
    return r


def load_enhanced(path: str) -> np.ndarray:
    """open_file (:651) + prologue (:142-150) at GUI defaults -> contrast-enhanced RGB array."""
    from PIL import Image
    return enhance(Image.open(path).convert('RGB'))


# --------------------------------------------------------------------------- output stage
# Replay of align_board (img2sgf.py:484-494) and to_SGF (:781-810) with the globals / Tk variable
# turned into arguments; loops kept element by element like the reference (checker for
# img2sgf_b200/sgf.py).
ALIGN_TOP, ALIGN_BOTTOM, ALIGN_LEFT, ALIGN_RIGHT = range(4)


def align_board(b, hsize, vsize, a=(ALIGN_LEFT, ALIGN_TOP)):
    board = np.zeros((BOARD_SIZE, BOARD_SIZE))
    xoffset = BOARD_SIZE - hsize if a[0] == ALIGN_RIGHT else 0
    yoffset = BOARD_SIZE - vsize if a[1] == ALIGN_BOTTOM else 0
    for i in range(hsize):
        for j in range(vsize):
            board[i + xoffset, j + yoffset] = b[i, j]
    return board


def to_SGF(board, side_to_move):
    import string
    letters = string.ascii_lowercase
    lines = ["(;GM[1]FF[4]SZ[" + str(BOARD_SIZE) + "]", "PL[B]" if side_to_move == 1 else "PL[W]"]
    moves = {}
    for colour, tag in ((1, "AB"), (2, "AW")):
        text = ""
        if colour in board:
            text = tag
            for i in range(BOARD_SIZE):
                for j in range(BOARD_SIZE):
                    if board[i, j] == colour:
                        text += "[" + letters[i] + letters[j] + "]"
        moves[colour] = text
    order = (1, 2) if side_to_move == 1 else (2, 1)
    lines += [moves[order[0]], moves[order[1]], ")"]
    return "\n".join(lines) + "\n"
