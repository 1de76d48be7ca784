"""The C restatement (oracle/img2sgf_oracle.c) against the golden vectors produced by the real
cv2/sklearn/numpy/PIL calls (tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest

from conftest import FIXTURES, load_input


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", FIXTURES)
def test_pipeline_matches_golden(name, golden, oracle):
    rgb = load_input(name)
    p = name + "/"
    assert tuple(golden[p + "shape"]) == rgb.shape[:2]
    thr = int(golden[p + "threshold"])
    assert oracle.choose_threshold(rgb.shape[1], rgb.shape[0]) == thr
    res, circles, masked = oracle.pipeline(rgb, thr)
    assert sha(oracle.grey(rgb)) == str(golden[p + "grey_sha"])
    assert sha(masked) == str(golden[p + "masked_sha"])
    np.testing.assert_array_equal(circles, golden[p + "circles"].reshape(-1, 3))
    assert res.n_hlines == len(golden[p + "hlines"]) and res.n_vlines == len(golden[p + "vlines"])
    assert res.n_hcentres == len(golden[p + "hcentres"]) and res.n_vcentres == len(golden[p + "vcentres"])
    assert bool(res.grid.valid) == bool(golden[p + "valid"])
    if res.grid.valid:
        assert [res.grid.hsize, res.grid.vsize] == list(golden[p + "sizes"])
        assert [res.grid.hspace, res.grid.vspace] == list(golden[p + "spaces"])
        np.testing.assert_array_equal(np.array(res.grid.hc[:res.grid.vsize]), golden[p + "hcentres_complete"])
        np.testing.assert_array_equal(np.array(res.grid.vc[:res.grid.hsize]), golden[p + "vcentres_complete"])
    assert bool(res.board_ready) == bool(golden[p + "board_ready"])
    if res.board_ready:
        np.testing.assert_array_equal(oracle.board_of(res), golden[p + "board"])


@pytest.mark.parametrize("name", ["ex9", "ex10", "ex7", "no_circles"])
def test_stages_match_golden(name, golden, oracle):
    rgb = load_input(name)
    p = name + "/"
    grey = oracle.grey(rgb)
    edges = oracle.canny_rgb(rgb)
    assert sha(edges) == str(golden[p + "edges_sha"])
    blurs = [grey, edges]
    for b in (1, 3, 5, 7):
        blurs += [oracle.median(grey, b), oracle.gauss(grey, b)]
    for k, img in enumerate(blurs):
        assert sha(img) == str(golden[p + f"blur{k}_sha"]), f"blur {k}"
        np.testing.assert_array_equal(oracle.hough_circles(img), golden[p + f"circles{k}"].reshape(-1, 3))
    masked = oracle.mask_circles(edges, golden[p + "circles"])
    assert sha(masked) == str(golden[p + "masked_sha"])
    thr = int(golden[p + "threshold"])
    hl, vl = oracle.find_lines(masked, thr, 1), oracle.find_lines(masked, thr, 2)
    np.testing.assert_array_equal(hl.reshape(-1), golden[p + "hlines"])
    np.testing.assert_array_equal(vl.reshape(-1), golden[p + "vlines"])
    np.testing.assert_array_equal(oracle.cluster(hl), golden[p + "hcentres"])
    np.testing.assert_array_equal(oracle.cluster(vl), golden[p + "vcentres"])


def test_primitives_on_random_vectors(random_vectors, oracle):
    rv = random_vectors
    tags = sorted({k.split("/")[0] for k in rv.files})
    assert len(tags) == 12
    for t in tags:
        rgb = rv[t + "/rgb"]
        g = oracle.grey(rgb)
        np.testing.assert_array_equal(g, rv[t + "/grey"], err_msg=t)
        np.testing.assert_array_equal(oracle.canny_rgb(rgb), rv[t + "/canny_rgb"], err_msg=t)
        np.testing.assert_array_equal(oracle.canny_grey(g), rv[t + "/canny_grey"], err_msg=t)
        for b in (3, 5, 7):
            np.testing.assert_array_equal(oracle.median(g, b), rv[t + f"/median{b}"], err_msg=t)
            np.testing.assert_array_equal(oracle.gauss(g, b), rv[t + f"/gauss{b}"], err_msg=t)
        np.testing.assert_array_equal(oracle.contrast(rgb, 102 / (101 - 70) - 1), rv[t + "/contrast"], err_msg=t)


def test_edge_cases(oracle):
    # cluster: fewer than two lines -> [] (AgglomerativeClustering.fit raises, img2sgf.py:273-278)
    assert len(oracle.cluster(np.zeros((0, 1), np.float32))) == 0
    assert len(oracle.cluster(np.array([[5.0]], np.float32))) == 0
    # gap 9 merges, gap 10 splits
    np.testing.assert_array_equal(oracle.cluster(np.array([0, 9, 19, 40], np.float32)), [4.5, 19.0, 40.0])
    # complete_grid: None for <2 lines, too-close lines, too many implied lines
    assert oracle.complete_grid([]) is None and oracle.complete_grid([3.0]) is None
    assert oracle.complete_grid([0.0, 5.0, 30.0]) is None
    assert oracle.complete_grid(np.arange(0, 2000, 20.0)[[0, 1, 99]]) is None
    # a gap of two spacings is filled by linear interpolation
    np.testing.assert_array_equal(oracle.complete_grid([0.0, 20.0, 60.0, 80.0]), [0, 20, 40, 60, 80])
    # blank image: nothing found anywhere
    blank = np.full((64, 80, 3), 255, np.uint8)
    res, circles, masked = oracle.pipeline(blank, 30)
    assert res.n_circles == 0 and res.n_hlines == 0 and not res.grid.valid and not masked.any()


def test_synthetic_truth(oracle):
    from img2sgf_b200 import synth
    g, truth = synth.diagram(640, 30, 14, seed=1)
    res, circles, _ = oracle.pipeline(synth.to_rgb(g), 80)
    assert res.board_ready and (res.grid.hsize, res.grid.vsize) == (19, 19)
    np.testing.assert_array_equal(oracle.board_of(res), truth)
