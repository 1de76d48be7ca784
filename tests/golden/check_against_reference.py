"""Pin the committed goldens to the reference's OWN code (not to the transcription in oracle/).

    python tests/golden/check_against_reference.py [--log profiles/r2_reference_pin.txt]

Reads /root/reference/img2sgf.py (build container only; the GPU box never runs this), executes
its Parts 1-3 source text unmodified -- everything above "# Part 4", i.e. the imports/constants,
the image-processing functions and the GUI callbacks, but not the widget construction and main
loop -- inside a namespace where the modules that are absent here (tkinter, matplotlib,
pyscreenshot) are replaced by inert stand-ins, and the Tk variables/widgets Part 4 would have
created are stand-ins holding the GUI defaults (img2sgf.py:616-639).  Then, for each of the 18
test images, it calls the reference's own open_file() (:643), which runs
initialise_parameters() -> process_image() -> find_grid() -> identify_board(), and compares the
globals the reference leaves behind with tests/golden/golden.npz:

    grey / edge / masked image hashes, stacked circles, hlines / vlines, cluster centres,
    valid_grid, hsize / vsize / hspace / vspace, completed grids, detected board, brightnesses.

No reference code is copied into the repo: the source is read from /root/reference at run time.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/img2sgf.py"
REF_IMAGES = "/root/reference/test_images"
NAMES = [f"ex{i}" for i in range(1, 18)] + ["no_circles"]


class Inert:
    """Accepts any attribute access, call, item access or iteration and does nothing."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return Inert()

    def __call__(self, *a, **k):
        return Inert()

    def __getitem__(self, k):
        return Inert()

    def __iter__(self):
        return iter(())

    def __len__(self):
        return 0


class Var:
    """Stand-in for tk.Scale / tk.IntVar: holds a value behind get()/set()."""

    def __init__(self, value=0):
        self.value = value

    def get(self):
        return self.value

    def set(self, v):
        self.value = v

    def __getattr__(self, name):       # .configure(), .bind(), ...
        return Inert()


def _stub_module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__getattr__ = lambda attr: Inert()      # PEP 562: any other name resolves to an inert object
    sys.modules[name] = m
    return m


def load_reference_namespace():
    """Exec the reference's Parts 1-3 with stand-ins for the GUI toolkits.  Returns the namespace."""
    tk = _stub_module("tkinter", END="end", ACTIVE="active", DISABLED="disabled", NORMAL="normal",
                      HORIZONTAL="horizontal")
    for sub in ("messagebox", "filedialog", "scrolledtext"):
        setattr(tk, sub, _stub_module("tkinter." + sub))
    mpl = _stub_module("matplotlib", __version__="stub")
    _stub_module("matplotlib.backends")
    _stub_module("matplotlib.backends.backend_tkagg")
    _stub_module("matplotlib.figure")
    _stub_module("pyscreenshot")
    del mpl
    src = open(REF).read()
    cut = src.index("# Part 4")
    ns = {"__name__": "img2sgf_reference", "__file__": REF}
    exec(compile(src[:cut], REF, "exec"), ns)
    # what Part 4 would have created (img2sgf.py:1016-1238), at the values initialise_parameters()
    # and the widget definitions give them: Canny 50/200, Sobel 3, L1 (:47-50, :1147-1181)
    ns.update(
        log_text=Inert(), save_button=Inert(), reset_button=Inert(), threshold_subfigure=Inert(),
        threshold_plot=Inert(), black_thresh_subfigure=Inert(), black_thresh_hist=Inert(),
        input_canvas=Inert(), processed_canvas=Inert(), output_canvas=Inert(), main_window=Inert(),
        show_circles=Var(0), rotate_angle=Var(0), contrast=Var(ns["contrast_default"]),
        brightness=Var(ns["brightness_default"]), edge_min=Var(ns["edge_min_default"]),
        edge_max=Var(ns["edge_max_default"]), sobel=Var(ns["sobel_default"]),
        gradient=Var(ns["gradient_default"]), threshold=Var(ns["threshold_default"]), side_to_move=Var(1),
        threshold_line=None, threshold_hist=None,
    )
    # drawing callbacks of Part 3 touch canvases only: make them no-ops (they compute nothing)
    for fn in ("draw_images", "draw_board", "draw_histogram"):
        ns[fn] = lambda *a, **k: None
    return ns


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_reference(ns, path):
    """open_file() of the reference on one image; returns the globals it leaves behind."""
    ns["open_file"](path)
    g = lambda k: ns.get(k)
    return {k: g(k) for k in (
        "input_image_np", "grey_image_np", "edge_detected_image_np", "circles", "circles_removed_image_np",
        "hcentres", "vcentres", "valid_grid", "board_ready", "hsize", "vsize", "hspace", "vspace",
        "hcentres_complete", "vcentres_complete", "detected_board", "stone_brightnesses", "threshold")}


def compare(name, r, gold, ns):
    p = name + "/"
    errs = []

    def eq(what, got, want):
        got, want = np.asarray(got), np.asarray(want)
        if got.shape != want.shape or not np.array_equal(got, want):
            errs.append(f"{what}: got shape {got.shape}, want {want.shape}")

    if sha(r["grey_image_np"]) != str(gold[p + "grey_sha"]):
        errs.append("grey hash")
    if sha(r["edge_detected_image_np"]) != str(gold[p + "edges_sha"]):
        errs.append("edges hash")
    if sha(r["circles_removed_image_np"]) != str(gold[p + "masked_sha"]):
        errs.append("masked hash")
    if int(r["threshold"].get()) != int(gold[p + "threshold"]):
        errs.append("auto line threshold")
    # `circles` after find_grid() is the radius-filtered list when the grid is valid (:441-443,:555);
    # the stacked list (before the filter) is what golden "circles" holds -> compare through the filter
    stacked = gold[p + "circles"].reshape(-1, 3)
    valid = bool(r["valid_grid"])
    if valid != bool(gold[p + "valid"]):
        errs.append("valid_grid")
    if valid:
        hs, vs = float(gold[p + "spaces"][0]), float(gold[p + "spaces"][1])
        lo, hi = min(hs, vs) * 0.3, max(hs, vs) * 0.65
        want = np.array([c for c in stacked if lo < c[2] < hi], np.float32).reshape(-1, 3)
        eq("filtered circles", np.asarray(r["circles"], np.float32).reshape(-1, 3), want)
        eq("sizes", [r["hsize"], r["vsize"]], gold[p + "sizes"])
        eq("spaces", np.array([r["hspace"], r["vspace"]], np.float64), gold[p + "spaces"])
        eq("hcentres_complete", r["hcentres_complete"], gold[p + "hcentres_complete"])
        eq("vcentres_complete", r["vcentres_complete"], gold[p + "vcentres_complete"])
    else:
        eq("stacked circles", np.asarray(r["circles"], np.float32).reshape(-1, 3), stacked)
    # the line columns are not kept in a global: re-ask the reference's own find_lines on its masked image
    hl = ns["find_lines"](r["threshold"].get(), ns["Direction"].HORIZONTAL)
    vl = ns["find_lines"](r["threshold"].get(), ns["Direction"].VERTICAL)
    eq("hlines", np.asarray(hl, np.float32).reshape(-1), gold[p + "hlines"])
    eq("vlines", np.asarray(vl, np.float32).reshape(-1), gold[p + "vlines"])
    eq("hcentres", np.asarray(r["hcentres"] if r["hcentres"] is not None else [], np.float64).reshape(-1),
       gold[p + "hcentres"])
    eq("vcentres", np.asarray(r["vcentres"] if r["vcentres"] is not None else [], np.float64).reshape(-1),
       gold[p + "vcentres"])
    ready = bool(r["board_ready"])
    if ready != bool(gold[p + "board_ready"]):
        errs.append("board_ready")
    if ready:
        eq("board", np.asarray(r["detected_board"]).astype(np.uint8), gold[p + "board"])
        eq("brightness", np.asarray(r["stone_brightnesses"], np.float64), gold[p + "brightness"])
    return errs


def check_inputs(name, r):
    """tests/golden/inputs/<name>.png must be the array the reference feeds to cvtColor/Canny (:150)."""
    from PIL import Image
    a = np.array(Image.open(os.path.join(HERE, "inputs", name + ".png")))
    if a.ndim == 2:
        a = np.repeat(a[..., None], 3, axis=-1)
    return [] if np.array_equal(a[..., :3], r["input_image_np"]) else ["committed input PNG differs from the reference's array"]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--log", default=None)
    args = ap.parse_args(argv)
    if not os.path.exists(REF):
        print("reference not present (this check runs in the build container only)")
        return 2
    import cv2
    import sklearn
    import PIL
    ns = load_reference_namespace()
    gold = np.load(os.path.join(HERE, "golden.npz"))
    lines = [f"pin of tests/golden/golden.npz to {REF} (Parts 1-3 executed unmodified, GUI defaults); "
             f"cv2 {cv2.__version__}, sklearn {sklearn.__version__}, numpy {np.__version__}, PIL {PIL.__version__}"]
    bad = 0
    for name in NAMES:
        ns["threshold_line"] = None
        r = run_reference(ns, os.path.join(REF_IMAGES, name + ".jpg"))
        errs = compare(name, r, gold, ns) + check_inputs(name, r)
        bad += bool(errs)
        n = len(np.asarray(r["circles"]).reshape(-1, 3)) if len(r["circles"]) else 0
        lines.append(f"{name:11s} {r['input_image_np'].shape[1]}x{r['input_image_np'].shape[0]} "
                     f"valid_grid={bool(r['valid_grid'])} board_ready={bool(r['board_ready'])} circles_after_filter={n}: "
                     + ("OK" if not errs else "MISMATCH " + "; ".join(errs)))
    lines.append(f"{len(NAMES) - bad}/{len(NAMES)} fixtures identical to the goldens")
    text = "\n".join(lines)
    print(text)
    if args.log:
        with open(args.log, "w") as f:
            f.write(text + "\n")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
