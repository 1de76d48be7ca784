"""Generate the golden vectors under tests/golden/ from the REAL reference library calls.

Run in the build container (needs /root/reference, cv2, sklearn, PIL):
    python tests/golden/make_golden.py

For each of the 18 reference fixtures (/root/reference/test_images/*.jpg) it
  1. runs the reference's prologue (open_file :651 + contrast/brightness :142-150 at GUI defaults)
     and stores the contrast-enhanced RGB array losslessly as inputs/<name>.png -- the input of
     the hot path (greyscale sources are stored single-channel);
  2. replays the hot path with oracle/ref_replay.py (identical cv2/sklearn/numpy calls) and
     stores every small output (circles per call, line columns, cluster centres, grid, board,
     brightnesses) plus SHA-1 hashes of the image-sized ones in golden.npz.
It also stores primitive-level vectors on seeded random images (random_vectors.npz).
"""
import hashlib
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_replay as R  # noqa: E402

REF_IMAGES = "/root/reference/test_images"
NAMES = [f"ex{i}" for i in range(1, 18)] + ["no_circles"]


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import cv2 as cv
    out = {}
    for name in NAMES:
        rgb = R.load_enhanced(os.path.join(REF_IMAGES, name + ".jpg"))
        grey_src = (rgb[..., 0] == rgb[..., 1]).all() and (rgb[..., 1] == rgb[..., 2]).all()
        Image.fromarray(rgb[..., 0] if grey_src else rgb).save(
            os.path.join(HERE, "inputs", name + ".png"), optimize=True)
        r = R.run(rgb)
        p = name + "/"
        out[p + "shape"] = np.array(rgb.shape[:2])
        out[p + "threshold"] = np.array(r.threshold)
        out[p + "grey_sha"] = np.array(sha(r.grey))
        out[p + "edges_sha"] = np.array(sha(r.edges))
        out[p + "masked_sha"] = np.array(sha(r.masked))
        for k, b in enumerate(r.blurs):
            out[p + f"blur{k}_sha"] = np.array(sha(b))
            out[p + f"circles{k}"] = r.per_call_circles[k]
        out[p + "circles"] = r.circles
        out[p + "hlines"] = np.asarray(r.hlines, np.float32).reshape(-1)
        out[p + "vlines"] = np.asarray(r.vlines, np.float32).reshape(-1)
        out[p + "hcentres"] = np.asarray(r.hcentres, np.float64).reshape(-1)
        out[p + "vcentres"] = np.asarray(r.vcentres, np.float64).reshape(-1)
        g = r.grid
        out[p + "valid"] = np.array(bool(g.valid))
        if g.valid:
            out[p + "sizes"] = np.array([g.hsize, g.vsize])
            out[p + "spaces"] = np.array([g.hspace, g.vspace], np.float64)
            out[p + "hcentres_complete"] = np.asarray(g.hcentres_complete, np.float64)
            out[p + "vcentres_complete"] = np.asarray(g.vcentres_complete, np.float64)
        out[p + "board_ready"] = np.array(r.board is not None)
        if r.board is not None:
            out[p + "board"] = r.board.astype(np.uint8)
            out[p + "brightness"] = r.brightness
        print(name, rgb.shape, len(r.circles), "board" if r.board is not None else "no board")
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)

    # primitive-level vectors on seeded random inputs (small, committed)
    rv = {}
    rng = np.random.default_rng(1234)
    for k, (h, w) in enumerate([(37, 53), (64, 64), (50, 131), (97, 40)]):
        noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        smooth = cv.GaussianBlur(noise, (0, 0), 2.0)
        quant = (smooth // 32 * 32).astype(np.uint8)
        for tag, img in (("noise", noise), ("smooth", smooth), ("quant", quant)):
            p = f"{tag}{k}/"
            g = cv.cvtColor(img, cv.COLOR_BGR2GRAY)
            rv[p + "rgb"] = img
            rv[p + "grey"] = g
            rv[p + "canny_rgb"] = cv.Canny(img, 50, 200, apertureSize=3, L2gradient=False)
            rv[p + "canny_grey"] = cv.Canny(g, 50, 100, apertureSize=3, L2gradient=False)
            for b in (3, 5, 7):
                rv[p + f"median{b}"] = cv.medianBlur(g, b)
                rv[p + f"gauss{b}"] = cv.GaussianBlur(g, (b, b), b)
            rv[p + "contrast"] = np.array(R.enhance(Image.fromarray(img)))
    np.savez_compressed(os.path.join(HERE, "random_vectors.npz"), **rv)
    print("wrote golden.npz, random_vectors.npz")
    # the goldens came from the replay (oracle/ref_replay.py): prove they equal what the reference's
    # own source produces (img2sgf.py Parts 1-3 executed unmodified)
    import check_against_reference
    if check_against_reference.main(["--log", os.path.join(HERE, "..", "..", "profiles", "r2_reference_pin.txt")]) != 0:
        raise SystemExit("goldens differ from the reference's own code")


if __name__ == "__main__":
    main()
