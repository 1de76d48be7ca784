"""Pins the C restatement against the real third-party calls the reference makes (cv2, sklearn,
numpy, PIL) on inputs OUTSIDE the committed goldens.  Skipped where those libraries are absent."""
import numpy as np
import pytest

cv = pytest.importorskip("cv2")
pytest.importorskip("sklearn")

from oracle import ref_replay as R  # noqa: E402


@pytest.mark.parametrize("seed", range(6))
def test_random_primitives(seed, oracle):
    rng = np.random.default_rng(100 + seed)
    h, w = int(rng.integers(20, 120)), int(rng.integers(20, 120))
    rgb = cv.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), float(rng.uniform(0.5, 3)))
    if seed % 2:
        rgb = (rgb // 16 * 16).astype(np.uint8)          # tie-heavy
    g = R.grey(rgb)
    np.testing.assert_array_equal(oracle.grey(rgb), g)
    np.testing.assert_array_equal(oracle.canny_rgb(rgb), R.edge_map(rgb))
    np.testing.assert_array_equal(oracle.canny_grey(g), cv.Canny(g, 50, 100))
    for b in (3, 5, 7):
        np.testing.assert_array_equal(oracle.median(g, b), cv.medianBlur(g, b))
        np.testing.assert_array_equal(oracle.gauss(g, b), cv.GaussianBlur(g, (b, b), b))


@pytest.mark.parametrize("seed", range(3))
def test_synthetic_pipeline(seed, oracle):
    from img2sgf_b200 import synth
    g, truth = synth.diagram(640, 30, 14, seed=seed, noise=2.0 if seed else 0.0, numbered=(seed == 2))
    rgb = synth.to_rgb(g)
    r = R.run(rgb, threshold=66)
    res, circles, masked = oracle.pipeline(rgb, 66)
    np.testing.assert_array_equal(circles, r.circles.reshape(-1, 3))
    np.testing.assert_array_equal(masked, r.masked)
    assert res.n_hlines == len(r.hlines) and res.n_vlines == len(r.vlines)
    assert bool(res.grid.valid) == r.grid.valid
    assert (r.board is not None) == bool(res.board_ready)
    if r.board is not None:
        np.testing.assert_array_equal(oracle.board_of(res), r.board)


@pytest.mark.parametrize("seed", range(4))
def test_random_sparse_lines(seed, oracle):
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(40, 200)), int(rng.integers(40, 200))
    img = (rng.random((h, w)) < 0.03).astype(np.uint8) * 255
    for y in rng.integers(0, h, 4):
        img[y, rng.integers(0, w // 2):] = 255
    for x in rng.integers(0, w, 4):
        img[rng.integers(0, h // 2):, x] = 255
    for d in (R.HORIZONTAL, R.VERTICAL):
        ref = R.find_lines(img, 20, d)
        got = oracle.find_lines(img, 20, d)
        np.testing.assert_array_equal(np.asarray(ref, np.float32).reshape(-1), got.reshape(-1))
        ref_c = R.cluster(ref) if len(ref) else []
        np.testing.assert_array_equal(np.asarray(ref_c, np.float64), oracle.cluster(got))


@pytest.mark.parametrize("sliders", [(70, 50), (35, 80), (95, 20), (50, 50), (10, 100), (100, 0)])
def test_enhance_vs_pil(sliders, oracle):
    """ImageEnhance.Contrast / Brightness at several slider settings (img2sgf.py:142-149), incl. factors
    inside [0,1] (PIL's interpolation branch) and far outside (its clipping branch)."""
    from PIL import Image, ImageEnhance
    cs, bs = sliders
    fc, fb = 102 / (101 - cs) - 1, 450 / (200 - bs) - 2
    rng = np.random.default_rng(cs * 1000 + bs)
    for k in range(3):
        h, w = int(rng.integers(5, 90)), int(rng.integers(5, 90))
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if k == 1:
            rgb = (rgb // 3 + 160).astype(np.uint8)               # bright, low contrast
        want = np.array(ImageEnhance.Brightness(ImageEnhance.Contrast(Image.fromarray(rgb)).enhance(fc)).enhance(fb))
        np.testing.assert_array_equal(oracle.enhance(rgb, fc, fb), want)


@pytest.mark.skipif(not __import__("os").path.exists("/root/reference/img2sgf.py"),
                    reason="the reference source exists in the build container only")
def test_goldens_pinned_to_reference_source():
    """The committed goldens equal what the reference's OWN code (img2sgf.py Parts 1-3, executed
    unmodified with GUI stand-ins) leaves in its globals for the 18 test images."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import check_against_reference as chk
    assert chk.main([]) == 0
