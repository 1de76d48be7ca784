"""Host-side mirror of the reference's image-processing functions (img2sgf.py Part 2), backed by
the sm_100a kernels through the C ABI in include/img2sgf_b200.h.

The reference passes everything through module globals and Tk getters; here
every input is an explicit argument, names and return conventions are the reference's:

    enhance(rgb, fc, fb)                            img2sgf.py:142-149
    edge_map(rgb)                                   img2sgf.py:162-165
    find_circles(grey, edges) -> circles, masked    img2sgf.py:169-198
    find_lines(masked, threshold, direction)        img2sgf.py:230-255   ([] when nothing found)
    cluster(lines)                                  img2sgf.py:268-292   ([] when < 2 lines)
    validate_grid(hcentres, vcentres, circles)      img2sgf.py:420-445
    classify_stones(grey, circles, ...)             img2sgf.py:497-515,537-542
    process_image(rgb, ...)                         img2sgf.py:142-204 + find_grid :546-576
    process_images([rgb, ...])                      the same for a batch of images of any sizes

PyTorch is used only as the device-buffer carrier (allocation, H2D/D2H copies, stream handle).
There is no CPU fallback: without a CUDA device or the built library these functions raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from enum import Enum

import zlib

import numpy as np
import torch

from . import _native as N

BOARD_SIZE = 19
threshold_default = 80
black_stone_threshold_default = 128
edge_min_default, edge_max_default = 50, 200


class Direction(Enum):          # img2sgf.py:74-80
    HORIZONTAL = 1
    HORIZ = 1
    H = 1
    VERTICAL = 2
    VERT = 2
    V = 2


def choose_threshold(width: int, height: int) -> int:
    """img2sgf.py:606-613 on the image size."""
    t = int(min(width, height) / 12.8 + 16)
    return int(min(max(t, 20), 200))


# ------------------------------------------------------------------ device plumbing
def _require_cuda():
    if not torch.cuda.is_available():
        raise N.NativeError("img2sgf_b200 needs a CUDA device (there is no CPU fallback)")


def _dev(a: np.ndarray, dtype=None) -> torch.Tensor:
    a = np.ascontiguousarray(a, dtype=dtype)
    return torch.from_numpy(a).cuda(non_blocking=False)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _empty(shape, dtype):
    return torch.empty(shape, dtype=dtype, device="cuda")


def _grow(lim: N.Limits, status: int) -> N.Limits:
    new = N.Limits(lim.cand_cap, lim.circle_cap, lim.line_cap, lim.hyst_passes)
    if status & N.ST_CAND_OVERFLOW:
        new.cand_cap = min(lim.cand_cap * 2, 16384)
    if status & N.ST_CIRCLE_OVERFLOW:
        new.circle_cap = lim.circle_cap * 4
    if status & N.ST_LINE_OVERFLOW:
        new.line_cap = min(lim.line_cap * 4, 4096)
    if status & N.ST_HYST_NOT_CONVERGED:
        new.hyst_passes = lim.hyst_passes * 4
    if (new.cand_cap, new.circle_cap, new.line_cap, new.hyst_passes) == \
            (lim.cand_cap, lim.circle_cap, lim.line_cap, lim.hyst_passes):
        raise N.NativeError("limits exhausted: " + N.describe_status(status))
    return new


_RETRY_BITS = N.ST_CAND_OVERFLOW | N.ST_CIRCLE_OVERFLOW | N.ST_LINE_OVERFLOW | N.ST_HYST_NOT_CONVERGED


def _retrying(fn, lim=None):
    """Run fn(limits) -> (result, status int); enlarge the failing limit and retry."""
    lim = lim or N.default_limits()
    for _ in range(8):
        out, st = fn(lim)
        if not (st & _RETRY_BITS):
            return out
        lim = _grow(lim, st)
    raise N.NativeError("retry budget exhausted: " + N.describe_status(st))


# ------------------------------------------------------------------ stage-level entry points
def grey_image(rgb: np.ndarray) -> np.ndarray:
    """cv.cvtColor(rgb, COLOR_BGR2GRAY) on the RGB-ordered array -- img2sgf.py:153."""
    _require_cuda()
    h, w = rgb.shape[:2]
    d = _dev(rgb, np.uint8)
    out = _empty((h, w), torch.uint8)
    N.check(N.lib().i2s_grey(_ptr(d), 0, _ptr(out), 0, 1, h, w, _stream()), "i2s_grey")
    return out.cpu().numpy()


def enhance(rgb: np.ndarray, contrast_factor: float = 1.0, brightness_factor: float = 1.0) -> np.ndarray:
    """ImageEnhance.Contrast(img).enhance(fc) then ImageEnhance.Brightness(img).enhance(fb) --
    img2sgf.py:142-149 (fc = 102/(101-contrast)-1, fb = 450/(200-brightness)-2)."""
    _require_cuda()
    h, w = rgb.shape[:2]
    d = _dev(rgb, np.uint8)
    out = torch.empty_like(d)
    scratch = _empty((8,), torch.uint8)
    N.check(N.lib().i2s_enhance(_ptr(d), 0, _ptr(out), 0, _ptr(scratch), 1, h, w, float(contrast_factor),
                                float(brightness_factor), _stream()), "i2s_enhance")
    return out.cpu().numpy()


def contrast(rgb: np.ndarray, factor: float) -> np.ndarray:
    """ImageEnhance.Contrast(img).enhance(factor) -- img2sgf.py:142-144."""
    return enhance(rgb, factor, 1.0)


def scaled_contrast(slider: float) -> float:
    """img2sgf.py:142: slider range 0-100 -> factor 0.01-101, 50 -> 1.0."""
    return 102 / (101 - slider) - 1


def scaled_brightness(slider: float) -> float:
    """img2sgf.py:147: slider range 0-100 -> factor 0.25-2.5, 50 -> 1.0."""
    return 450 / (200 - slider) - 2


def gaussian_blurs(grey: np.ndarray):
    """(GaussianBlur(grey,(b,b),b) for b in 3,5,7) -- img2sgf.py:175."""
    _require_cuda()
    h, w = grey.shape
    d = _dev(grey, np.uint8)
    outs = [_empty((h, w), torch.uint8) for _ in range(3)]
    N.check(N.lib().i2s_gauss357(_ptr(d), _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), 1, h, w, 0, _stream()),
            "i2s_gauss357")
    return [o.cpu().numpy() for o in outs]


def median_blur(grey: np.ndarray, b: int) -> np.ndarray:
    """cv.medianBlur(grey, b) -- img2sgf.py:174."""
    _require_cuda()
    h, w = grey.shape
    d = _dev(grey, np.uint8)
    out = _empty((h, w), torch.uint8)
    N.check(N.lib().i2s_median(_ptr(d), _ptr(out), 1, h, w, 0, int(b), _stream()), "i2s_median")
    return out.cpu().numpy()


def _canny(img: np.ndarray, channels: int, low: int, high: int, hyst_passes=None) -> np.ndarray:
    _require_cuda()
    h, w = img.shape[:2]
    d = _dev(img, np.uint8)

    def run(lim):
        if hyst_passes is not None:
            lim = N.Limits(lim.cand_cap, lim.circle_cap, lim.line_cap, max(int(hyst_passes), lim.hyst_passes))
        out = _empty((h, w), torch.uint8)
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        nb = N.lib().i2s_canny_workspace_bytes(1, h, w)
        ws = _empty((nb,), torch.uint8)
        N.check(N.lib().i2s_canny(_ptr(d), channels, 0, _ptr(out), 0, 1, h, w, int(low), int(high), lim.hyst_passes,
                                  _ptr(status), _ptr(ws), nb, _stream()), "i2s_canny")
        return out.cpu().numpy(), int(status.item())

    return _retrying(run)


def edge_map(rgb: np.ndarray, low: int = edge_min_default, high: int = edge_max_default) -> np.ndarray:
    """cv.Canny(rgb, low, high, apertureSize=3, L2gradient=False) -- img2sgf.py:162-165."""
    return _canny(rgb, 3, low, high)


def canny_grey(img: np.ndarray, low: int = 50, high: int = 100, hyst_passes=None) -> np.ndarray:
    """The single-channel Canny cv.HoughCircles runs on its input (param1=100) -- img2sgf.py:180.
    `hyst_passes` starts the cross-tile pass budget higher than the default (it is enlarged on demand anyway)."""
    return _canny(img, 1, low, high, hyst_passes)


def hough_circles(img: np.ndarray, limits: N.Limits | None = None) -> np.ndarray:
    """cv.HoughCircles(img, HOUGH_GRADIENT, 1, 10, [], 100, 30, 1, 30)[0] -- img2sgf.py:180; (n,3) float32."""
    _require_cuda()
    h, w = img.shape
    d = _dev(img, np.uint8)

    def run(lim):
        circ = _empty((lim.circle_cap, 3), torch.float32)
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        nb = N.lib().i2s_hough_circles_workspace_bytes(1, h, w, C.byref(lim))
        ws = _empty((nb,), torch.uint8)
        N.check(N.lib().i2s_hough_circles(_ptr(d), 0, 1, h, w, _ptr(circ), _ptr(cnt), _ptr(status), C.byref(lim),
                                          _ptr(ws), nb, _stream()), "i2s_hough_circles")
        n = int(cnt.item())
        return circ[:min(n, lim.circle_cap)].cpu().numpy(), int(status.item())

    return _retrying(run, limits)


def mask_circles(edges: np.ndarray, circles: np.ndarray) -> np.ndarray:
    """The masking loop -- img2sgf.py:169,191-198."""
    _require_cuda()
    h, w = edges.shape
    c = np.ascontiguousarray(circles, np.float32).reshape(-1, 3)
    d = _dev(edges, np.uint8)
    out = torch.empty_like(d)
    cap = max(len(c), 1)
    dc = _dev(c if len(c) else np.zeros((1, 3), np.float32))
    cnt = torch.tensor([len(c)], dtype=torch.int32, device="cuda")
    N.check(N.lib().i2s_mask_circles(_ptr(d), _ptr(out), 0, 1, h, w, _ptr(dc), _ptr(cnt), cap, _stream()),
            "i2s_mask_circles")
    return out.cpu().numpy()


def find_circles(grey: np.ndarray, edges: np.ndarray):
    """Blur pyramid + ten HoughCircles calls stacked + masking -- img2sgf.py:169-198.

    Returns (circles (N,3) float32 in the reference's stacking order, duplicates kept;
    circles_removed_image u8).  N may be 0.
    """
    _require_cuda()
    h, w = grey.shape
    dg, de = _dev(grey, np.uint8), _dev(edges, np.uint8)

    def run(lim):
        circ = _empty((lim.circle_cap, 3), torch.float32)
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        masked = _empty((h, w), torch.uint8)
        nb = N.lib().i2s_find_circles_workspace_bytes(1, h, w, C.byref(lim))
        ws = _empty((nb,), torch.uint8)
        N.check(N.lib().i2s_find_circles(_ptr(dg), _ptr(de), 0, 1, h, w, _ptr(circ), _ptr(cnt), _ptr(masked),
                                         _ptr(status), C.byref(lim), _ptr(ws), nb, _stream()), "i2s_find_circles")
        n = int(cnt.item())
        return (circ[:min(n, lim.circle_cap)].cpu().numpy(), masked.cpu().numpy()), int(status.item())

    return _retrying(run)


_lines_memo = None      # (key, result) of the last _find_lines_both call


def _find_lines_both(masked: np.ndarray, threshold: int):
    """Both directions come out of one pass over the image.  The reference asks for them one at a time
    (find_lines(t, H) then find_lines(t, V), img2sgf.py:259-261): the last result is kept, keyed on the
    image's bytes, so the second call of such a pair costs a checksum instead of a second pass."""
    global _lines_memo
    _require_cuda()
    masked = np.ascontiguousarray(masked, np.uint8)
    h, w = masked.shape
    key = (h, w, int(threshold), zlib.crc32(masked), zlib.adler32(masked))
    memo = _lines_memo
    if memo is not None and memo[0] == key:
        return memo[1][0].copy(), memo[1][1].copy()
    d = _dev(masked, np.uint8)

    def run(lim):
        rho = _empty((2, lim.line_cap), torch.float32)
        cnt = torch.zeros(2, dtype=torch.int32, device="cuda")
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        nb = N.lib().i2s_find_lines_workspace_bytes(1, h, w)
        ws = _empty((nb,), torch.uint8)
        N.check(N.lib().i2s_find_lines(_ptr(d), 0, 1, h, w, int(threshold), _ptr(rho), _ptr(cnt), lim.line_cap,
                                       _ptr(status), _ptr(ws), nb, _stream()), "i2s_find_lines")
        c = cnt.cpu().numpy()
        r = rho.cpu().numpy()
        return (r[0, :c[0]].copy(), r[1, :c[1]].copy()), int(status.item())

    res = _retrying(run)
    _lines_memo = (key, res)
    return res[0].copy(), res[1].copy()


def find_lines(masked: np.ndarray, threshold: int, direction):
    """find_lines(threshold, direction) -- img2sgf.py:230-255.  (n,1) float32 column of rho, or []."""
    hl, vl = _find_lines_both(masked, threshold)
    is_h = direction in (Direction.H, 1, "H", "h")
    col = hl if is_h else vl
    return [] if len(col) == 0 else col.reshape(-1, 1)


def find_all_lines(masked: np.ndarray, threshold: int):
    """find_all_lines() -- img2sgf.py:258-265 (both directions from one pass over the image)."""
    hl, vl = _find_lines_both(masked, threshold)
    f = lambda c: [] if len(c) == 0 else c.reshape(-1, 1)
    return f(hl), f(vl)


def cluster(lines):
    """find_clusters_fixed_threshold + get_cluster_centres -- img2sgf.py:268-292.
    Sorted float64 centres, or [] when fewer than two lines."""
    _require_cuda()
    if lines is None or len(lines) < 2:
        return []
    col = np.ascontiguousarray(np.asarray(lines, np.float32).reshape(-1))
    cap = 2
    while cap < len(col):
        cap *= 2
    if cap > 4096:
        raise N.NativeError("cluster: more than 4096 lines")
    rho = np.zeros((2, cap), np.float32)
    rho[0, :len(col)] = col
    d = _dev(rho)
    cnt = torch.tensor([len(col), 0], dtype=torch.int32, device="cuda")
    cen = _empty((2, cap), torch.float64)
    k = torch.zeros(2, dtype=torch.int32, device="cuda")
    N.check(N.lib().i2s_cluster(_ptr(d), _ptr(cnt), 1, cap, _ptr(cen), _ptr(k), _stream()), "i2s_cluster")
    n = int(k[0].item())
    return cen[0, :n].cpu().numpy()


def _validate_raw(hcentres, vcentres):
    nh = 0 if hcentres is None else len(hcentres)
    nv = 0 if vcentres is None else len(vcentres)
    cap = 2
    while cap < max(nh, nv):
        cap *= 2
    cen = np.zeros((2, cap), np.float64)
    if nh:
        cen[0, :nh] = np.asarray(hcentres, np.float64)
    if nv:
        cen[1, :nv] = np.asarray(vcentres, np.float64)
    d = _dev(cen)
    k = torch.tensor([nh, nv], dtype=torch.int32, device="cuda")
    grid = torch.zeros(N.GRID_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    N.check(N.lib().i2s_validate_grid(_ptr(d), _ptr(k), 1, cap, _ptr(grid), _ptr(status), _stream()),
            "i2s_validate_grid")
    return grid.cpu().numpy().view(N.GRID_DTYPE)[0], int(status.item())


def validate_grid(hcentres, vcentres, circles):
    """validate_grid -- img2sgf.py:420-445.  Same 8-element result as the reference."""
    _require_cuda()
    g, st = _validate_raw(hcentres, vcentres)
    if not g["valid"]:
        return [False, circles, 0, 0] + 4 * [None]
    if st & N.ST_GRID_OVERFLOW:
        raise N.NativeError("validate_grid: more than 32 grid lines on an axis")
    vsize, hsize = int(g["vsize"]), int(g["hsize"])
    hspace, vspace = np.float64(g["hspace"]), np.float64(g["vspace"])
    lo, hi = min(hspace, vspace) * 0.3, max(hspace, vspace) * 0.65
    newcircles = [c for c in circles if lo < c[2] < hi]
    return (True, newcircles, vsize, hsize, g["hcentres"][:vsize].copy(), g["vcentres"][:hsize].copy(),
            hspace, vspace)


def classify_stones(grey: np.ndarray, circles, hcentres_complete, vcentres_complete, hspace, vspace,
                    black_stone_threshold: int = black_stone_threshold_default):
    """identify_board -- img2sgf.py:497-515,537-542, callable on its own (black-threshold drag,
    :762-765).  Returns (detected_board float64 (hsize, vsize) with 0/1/2, stone_brightnesses)."""
    _require_cuda()
    h, w = grey.shape
    vsize, hsize = len(hcentres_complete), len(vcentres_complete)
    if hsize > BOARD_SIZE or vsize > BOARD_SIZE:
        raise ValueError("grid larger than 19x19 (the reference does not classify it, img2sgf.py:568-571)")
    g = np.zeros(1, N.GRID_DTYPE)
    g["valid"], g["hsize"], g["vsize"] = 1, hsize, vsize
    g["hspace"], g["vspace"] = float(hspace), float(vspace)
    g["hcentres"][0, :vsize] = np.asarray(hcentres_complete, np.float64)
    g["vcentres"][0, :hsize] = np.asarray(vcentres_complete, np.float64)
    c = np.ascontiguousarray(np.asarray(circles, np.float32).reshape(-1, 3))
    cap = max(len(c), 1)
    dc = _dev(c if len(c) else np.zeros((1, 3), np.float32))
    cnt = torch.tensor([len(c)], dtype=torch.int32, device="cuda")
    dgrid = _dev(g.view(np.uint8))
    dgrey = _dev(grey, np.uint8)
    rec = torch.zeros(N.RECORD_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    br = torch.zeros(BOARD_SIZE * BOARD_SIZE, dtype=torch.float64, device="cuda")
    N.check(N.lib().i2s_classify_stones(_ptr(dgrey), 0, 1, h, w, _ptr(dc), _ptr(cnt), cap, _ptr(dgrid),
                                        int(black_stone_threshold), _ptr(rec), _ptr(br), _stream()),
            "i2s_classify_stones")
    r = rec.cpu().numpy().view(N.RECORD_DTYPE)[0]
    board = r["board"].reshape(BOARD_SIZE, BOARD_SIZE)[:hsize, :vsize].astype(np.float64)
    k = int(r["n_black"]) + int(r["n_white"])
    return board, br.cpu().numpy()[:k]


# ------------------------------------------------------------------ whole path, one image
@dataclass
class Processed:
    """Everything process_image()/find_grid() leave in the reference's globals (img2sgf.py:118-120,
    498-499, 547-548), for one image.  `circles` is the stacked list the ten HoughCircles calls give
    (:179-186, what the masking loop uses); `circles_in_grid` is what the global `circles` holds after
    find_grid(): the radius-filtered list when the grid is valid (:441-443, :555), else the same list."""
    grey_image_np: np.ndarray
    edge_detected_image_np: np.ndarray
    circles: np.ndarray
    circles_in_grid: np.ndarray
    circles_removed_image_np: np.ndarray
    hlines: object
    vlines: object
    valid_grid: bool
    hsize: int
    vsize: int
    hspace: float
    vspace: float
    hcentres_complete: object
    vcentres_complete: object
    board_ready: bool
    detected_board: object
    full_board: object
    stone_brightnesses: object
    num_black_stones: int
    num_white_stones: int
    record: object


def process_image(rgb: np.ndarray, threshold: int | None = None,
                  black_stone_threshold: int = black_stone_threshold_default, contrast_slider: float | None = None,
                  brightness_slider: float | None = None, selection=None) -> Processed:
    """process_image() (img2sgf.py:117-204) through find_grid() and identify_board() -- one i2s_pipeline
    call.  `rgb`: [h,w,3] u8 in PIL's RGB order, or [h,w] for a greyscale source.  By default the array
    is taken as already contrast-enhanced (:150); with `contrast_slider` / `brightness_slider` (the GUI's
    0..100 values, defaults 70 / 50) the prologue :142-149 runs on the device first.

    `selection` = (x0, y0, x1, y1): the crop of crop_and_rotate_image() (img2sgf.py:110-114, PIL box
    convention, rotation 0) without copying pixels: the whole image is uploaded once and the pipeline is
    handed a view of it -- an i2s_image_t with the offset of the box's first pixel and the full image's row
    pitch.  (The GUI re-runs process_image on every rubber-band zoom, :677-723.)"""
    from .batch import Engine, make_params
    _require_cuda()
    H, W = rgb.shape[:2]
    ch = 1 if rgb.ndim == 2 else 3
    x0, y0, x1, y1 = (0, 0, W, H) if selection is None else [int(v) for v in selection]
    if not (0 <= x0 < x1 <= W and 0 <= y0 < y1 <= H):
        raise ValueError("selection outside the image")
    h, w = y1 - y0, x1 - x0
    if threshold is None:
        threshold = choose_threshold(w, h)
    fc = scaled_contrast(contrast_slider) if contrast_slider is not None else 1.0
    fb = scaled_brightness(brightness_slider) if brightness_slider is not None else 1.0
    params = make_params(threshold, black_stone_threshold, contrast_factor=fc, brightness_factor=fb)
    full = np.ascontiguousarray(rgb, np.uint8)

    def run(lim):
        eng = Engine(1, h, w, limits=lim, taps=True)
        if selection is None:
            recs = eng.run_host(full[None], params, channels=ch)
        else:
            desc = np.zeros(1, N.IMAGE_DTYPE)
            desc[0] = ((y0 * W + x0) * ch, h, w, W * ch, 0)
            d_img, d_desc = _dev(full), _dev(desc.view(np.uint8))
            recs = eng.run(d_img.reshape(-1), params, n=1, channels=ch, images=d_desc).cpu().numpy() \
                      .view(N.RECORD_DTYPE).reshape(1)
        return (eng, recs), int(recs[0]["status"])

    eng, recs = _retrying(run)
    r = recs[0]
    if int(r["status"]) & N.ST_GRID_OVERFLOW:
        raise N.NativeError(f"more than {N.MAX_GRID} grid lines on an axis: the fixed-size grid record cannot hold "
                            "them (the reference reports 'too many lines' above 19, img2sgf.py:568-571)")
    t = eng.taps_host()
    nc = int(t["counts"][0])
    g = t["grids"][0]
    hs, vs = int(g["hsize"]), int(g["vsize"])
    lc = t["line_counts"][0]
    col = lambda a: [] if len(a) == 0 else a.reshape(-1, 1)
    ready = bool(r["board_ready"])
    full = r["board"].reshape(BOARD_SIZE, BOARD_SIZE).astype(np.float64)
    circles = t["circles"][0, :nc].copy()
    kept = circles
    if g["valid"]:
        lo, hi = min(g["hspace"], g["vspace"]) * 0.3, max(g["hspace"], g["vspace"]) * 0.65
        kept = circles[(circles[:, 2] > lo) & (circles[:, 2] < hi)] if nc else circles
    k = int(r["n_black"]) + int(r["n_white"])
    return Processed(
        grey_image_np=t["grey"][0], edge_detected_image_np=t["edges"][0], circles=circles, circles_in_grid=kept,
        circles_removed_image_np=t["masked"][0], hlines=col(t["rho"][0, 0, :lc[0]].copy()),
        vlines=col(t["rho"][0, 1, :lc[1]].copy()), valid_grid=bool(g["valid"]), hsize=hs, vsize=vs,
        hspace=float(g["hspace"]), vspace=float(g["vspace"]),
        hcentres_complete=g["hcentres"][:vs].copy() if g["valid"] else None,
        vcentres_complete=g["vcentres"][:hs].copy() if g["valid"] else None,
        board_ready=ready, detected_board=full[:hs, :vs].copy() if ready else None,
        full_board=full if ready else None, stone_brightnesses=t["brightness"][0, :k].copy() if ready else None,
        num_black_stones=int(r["n_black"]), num_white_stones=int(r["n_white"]), record=r)


def process_images(images, line_threshold=None, black_stone_threshold: int = black_stone_threshold_default,
                   runner=None, **kw) -> np.ndarray:
    """The batched form of process_image() for images of any sizes: one record per image (board, grid
    verdict, counts), no image-sized outputs.  See batch.RaggedRunner."""
    from .batch import RaggedRunner
    runner = runner or RaggedRunner()
    return runner.process_images(images, line_threshold, black_stone_threshold, **kw)
