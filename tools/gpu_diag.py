"""Stage-by-stage GPU-vs-oracle diagnosis for one image (debug aid, run on the GPU box).
Reads the HoughCircles workspace layout directly (state map at offset 0, accumulator after it)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from img2sgf_b200 import _native as N, api, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def diff(name, got, want):
    got, want = np.asarray(got), np.asarray(want)
    if got.shape != want.shape:
        print(f"  {name}: SHAPE {got.shape} vs {want.shape}")
        return False
    n = int((got != want).sum())
    print(f"  {name}: {'ok' if n == 0 else f'{n}/{got.size} differ, first {np.argwhere(got != want)[:4].tolist()}'}")
    return n == 0


def hough_diag(img, tag):
    h, w = img.shape
    lim = N.default_limits()
    d = torch.from_numpy(np.ascontiguousarray(img)).cuda()
    circ = torch.zeros((lim.circle_cap, 3), dtype=torch.float32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    nb = N.lib().i2s_hough_circles_workspace_bytes(1, h, w, C.byref(lim))
    ws = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    N.check(N.lib().i2s_hough_circles(C.c_void_p(d.data_ptr()), 1, h, w, C.c_void_p(circ.data_ptr()),
                                      C.c_void_p(cnt.data_ptr()), C.c_void_p(status.data_ptr()), C.byref(lim),
                                      C.c_void_p(ws.data_ptr()), nb, None), "hough")
    torch.cuda.synchronize()
    wsn = ws.cpu().numpy()
    plane = h * w
    state = wsn[:plane].reshape(h, w)
    oc, oedges, oacc = O.hough_circles(img, taps=True)
    print(f"[{tag}] {w}x{h} status={int(status.item())} count={int(cnt.item())} oracle={len(oc)}")
    diff("canny(50,100) edges", np.where(state & 2, 255, 0).astype(np.uint8), oedges)
    diff("nms candidates superset", (state & 1) >= (oedges > 0), np.ones_like(oedges, bool))
    n = int(cnt.item())
    diff("circles", circ[:n].cpu().numpy(), oc)


def main():
    from conftest import load_input
    names = sys.argv[1:] or ["ex9", "ex7"]
    for name in names:
        if name.startswith("synth"):
            g, _ = synth.diagram(640, 30, 14, seed=int(name[5:] or 0))
            rgb = synth.to_rgb(g)
        else:
            rgb = load_input(name)
        print("=== ", name, rgb.shape)
        grey = O.grey(rgb)
        diff("grey", api.grey_image(rgb), grey)
        edges = O.canny_rgb(rgb)
        diff("canny_rgb", api.edge_map(rgb), edges)
        for b, got in zip((3, 5, 7), api.gaussian_blurs(grey)):
            diff(f"gauss{b}", got, O.gauss(grey, b))
            diff(f"median{b}", api.median_blur(grey, b), O.median(grey, b))
        hough_diag(grey, "grey")
        hough_diag(edges, "edges")
        hough_diag(O.median(grey, 5), "median5")
        circles, masked = api.find_circles(grey, edges)
        res, ocirc, omasked = O.pipeline(rgb)
        diff("find_circles", circles, ocirc)
        diff("masked", masked, omasked)
        diff("mask only", api.mask_circles(edges, ocirc), omasked)
        thr = O.choose_threshold(rgb.shape[1], rgb.shape[0])
        for d, od in ((api.Direction.H, 1), (api.Direction.V, 2)):
            ol = O.find_lines(omasked, thr, od)
            gl = api.find_lines(omasked, thr, d)
            diff(f"lines dir{od}", np.asarray(gl, np.float32).reshape(-1), ol.reshape(-1))
            diff(f"cluster dir{od}", np.asarray(api.cluster(ol), np.float64), O.cluster(ol))
        r = api.process_image(rgb)
        print("  pipeline: valid", r.valid_grid, bool(res.grid.valid), "ready", r.board_ready, bool(res.board_ready),
              "status", int(r.record["status"]))
        if r.board_ready and res.board_ready:
            diff("board", r.detected_board.astype(np.uint8), O.board_of(res))


if __name__ == "__main__":
    main()
