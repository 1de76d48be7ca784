"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed
golden vectors.  Bit-exact for every integer/byte/index output; circles, rho columns, cluster
centres, grid geometry and brightnesses are compared for exact equality too (the float32/float64
stages are IEEE-exact restatements, SURVEY.md Fact 4) -- tolerance 0."""
import numpy as np
import pytest

from conftest import FIXTURES, load_input

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from img2sgf_b200 import api as A, build
    build.build()
    return A


def mismatch(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return f"shape {a.shape} vs {b.shape}"
    n = int((a != b).sum())
    if n == 0:
        return ""
    idx = np.argwhere(a != b)[:5].tolist()
    return f"{n} of {a.size} differ, first at {idx}: got {[a[tuple(i)].item() for i in idx]} want {[b[tuple(i)].item() for i in idx]}"


def assert_same(got, want, what):
    m = mismatch(got, want)
    assert not m, f"{what}: {m}"


# ------------------------------------------------------------------ primitives on random vectors
def test_primitives_random_vectors(api, random_vectors):
    rv = random_vectors
    for t in sorted({k.split("/")[0] for k in rv.files}):
        rgb = rv[t + "/rgb"]
        g = rv[t + "/grey"]
        assert_same(api.grey_image(rgb), g, t + " grey")
        assert_same(api.edge_map(rgb), rv[t + "/canny_rgb"], t + " canny_rgb")
        assert_same(api.canny_grey(g), rv[t + "/canny_grey"], t + " canny_grey")
        g3, g5, g7 = api.gaussian_blurs(g)
        for b, got in ((3, g3), (5, g5), (7, g7)):
            assert_same(got, rv[t + f"/gauss{b}"], t + f" gauss{b}")
            assert_same(api.median_blur(g, b), rv[t + f"/median{b}"], t + f" median{b}")
        assert_same(api.contrast(rgb, 102 / (101 - 70) - 1), rv[t + "/contrast"], t + " contrast")


@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (5, 200), (130, 7), (64, 64), (129, 257), (300, 260)])
def test_primitives_odd_shapes(api, oracle, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    h, w = shape
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    # smooth a little so Canny has structure; keep pure numpy
    k = np.ones(3) / 3
    if h >= 3 and w >= 3:
        f = rgb.astype(np.float32)
        f[1:-1] = (f[:-2] + f[1:-1] + f[2:]) / 3
        f[:, 1:-1] = (f[:, :-2] + f[:, 1:-1] + f[:, 2:]) / 3
        rgb = f.astype(np.uint8)
    g = oracle.grey(rgb)
    assert_same(api.grey_image(rgb), g, "grey")
    assert_same(api.edge_map(rgb), oracle.canny_rgb(rgb), "canny_rgb")
    assert_same(api.canny_grey(g), oracle.canny_grey(g), "canny_grey")
    gs = api.gaussian_blurs(g)
    for b, got in zip((3, 5, 7), gs):
        assert_same(got, oracle.gauss(g, b), f"gauss{b}")
        assert_same(api.median_blur(g, b), oracle.median(g, b), f"median{b}")


@pytest.mark.parametrize("size,noise", [(400, 0.0), (333, 0.0), (512, 3.0)])
def test_primitives_saturated_diagrams(api, oracle, size, noise):
    """Diagram-like content (saturated paper / ink, anti-aliased rims): exercises the saturated-window
    shortcut of the 7x7 median on whole warps, partially resolved warps, and noisy content where it
    never applies; plus both Cannys and the Gaussians on the same arrays."""
    from img2sgf_b200 import synth
    g, _ = synth.diagram(size, 16, 7, seed=size, noise=noise)
    g = np.ascontiguousarray(g)
    inv = np.ascontiguousarray(255 - g)                       # ink-dominated windows (pixel == 0 branch)
    half = g.copy(); half[:, size // 2:] = 0                  # a long straight 255 | 0 boundary
    for name, img in (("diagram", g), ("inverted", inv), ("half", half)):
        for b in (3, 5, 7):
            assert_same(api.median_blur(img, b), oracle.median(img, b), f"{name} median{b}")
        for b, got in zip((3, 5, 7), api.gaussian_blurs(img)):
            assert_same(got, oracle.gauss(img, b), f"{name} gauss{b}")
        assert_same(api.canny_grey(img), oracle.canny_grey(img), f"{name} canny_grey")
    rgb = synth.to_rgb(g)
    assert_same(api.edge_map(rgb), oracle.canny_rgb(rgb), "canny_rgb")


def test_canny_long_chain_needs_many_passes(api, oracle):
    """A weak spiral seeded by one strong pixel crosses many 128-px tiles: exercises the
    cross-tile hysteresis passes and the status/retry path."""
    h, w = 700, 700
    img = np.full((h, w), 128, np.uint8)
    # serpentine of low-contrast steps (weak edges), one high-contrast spot (strong seed)
    for k, y in enumerate(range(20, 680, 24)):
        img[y:y + 12, 10:690] = 146
    img[20:32, 10:40] = 255
    assert_same(api.canny_grey(img, 50, 100), oracle.canny_grey(img, 50, 100), "serpentine")


# ------------------------------------------------------------------ HoughCircles stage
@pytest.mark.parametrize("name", ["ex9", "ex10", "ex7", "ex3", "no_circles"])
def test_hough_circles_fixture_calls(api, oracle, golden, name):
    rgb = load_input(name)
    grey = oracle.grey(rgb)
    edges = oracle.canny_rgb(rgb)
    blurs = [grey, edges] + [f(grey, b) for b in (3, 5, 7) for f in (oracle.median, oracle.gauss)]
    gidx = [0, 1, 4, 5, 6, 7, 8, 9]
    for k, img in zip(gidx, blurs):
        got = api.hough_circles(img)
        assert_same(got, golden[f"{name}/circles{k}"].reshape(-1, 3), f"{name} HoughCircles call {k}")


def test_mask_semantics(api, oracle):
    rng = np.random.default_rng(5)
    edges = (rng.random((120, 150)) < 0.3).astype(np.uint8) * 255
    circles = np.stack([rng.uniform(-5, 155, 300), rng.uniform(-5, 125, 300), rng.uniform(1, 30, 300)], 1).astype(np.float32)
    circles[:, :2] = np.floor(circles[:, :2]) + 0.5
    circles[:, 2] = np.round(circles[:, 2] * 20) / 20
    assert_same(api.mask_circles(edges, circles), oracle.mask_circles(edges, circles), "mask")
    assert_same(api.mask_circles(edges, circles[:0]), edges, "mask with no circles")


# ------------------------------------------------------------------ lines / clusters / grid / classify
def test_lines_and_clusters_random(api, oracle):
    for seed in range(4):
        rng = np.random.default_rng(seed)
        h, w = int(rng.integers(40, 400)), int(rng.integers(40, 400))
        img = (rng.random((h, w)) < 0.03).astype(np.uint8) * 255
        for y in rng.integers(0, h, 5):
            img[y, rng.integers(0, w // 2):] = 255
        for x in rng.integers(0, w, 5):
            img[rng.integers(0, h // 2):, x] = 255
        for d, od in ((api.Direction.H, 1), (api.Direction.V, 2)):
            want = oracle.find_lines(img, 20, od)
            got = api.find_lines(img, 20, d)
            assert_same(np.asarray(got, np.float32).reshape(-1), want.reshape(-1), f"lines seed {seed} dir {od}")
            assert_same(np.asarray(api.cluster(got), np.float64), oracle.cluster(want), f"cluster seed {seed} dir {od}")
    assert api.find_lines(np.zeros((50, 60), np.uint8), 20, api.Direction.H) == []
    assert api.cluster([]) == [] and api.cluster(np.array([[3.0]], np.float32)) == []


def test_validate_grid_cases(api, oracle):
    cases = [
        (np.arange(19) * 30.0 + 5, np.arange(19) * 31.0 + 7),
        (np.delete(np.arange(19) * 30.0, [3, 4, 9]), np.delete(np.arange(19) * 30.0, [17])),
        (np.arange(21) * 25.0, np.arange(20) * 25.0),
        (np.array([0.0, 5.0, 50.0]), np.arange(19) * 30.0),
        (np.array([10.0]), np.arange(19) * 30.0),
        (np.array([]), np.array([])),
        (np.arange(25) * 20.0, np.arange(19) * 20.0),
        (np.array([0.0, 40.0, 61.0, 100.5, 141.0, 400.0]), np.arange(9) * 33.3),
    ]
    for hc, vc in cases:
        og = oracle.validate_grid(hc, vc)
        res = api.validate_grid(hc, vc, [])
        assert bool(res[0]) == bool(og.valid), (hc, vc)
        if og.valid:
            assert (res[2], res[3]) == (og.vsize, og.hsize)
            assert_same(res[4], np.array(og.hc[:og.vsize]), "hcentres_complete")
            assert_same(res[5], np.array(og.vc[:og.hsize]), "vcentres_complete")
            assert (res[6], res[7]) == (og.hspace, og.vspace)


# ------------------------------------------------------------------ whole path on the reference fixtures
@pytest.mark.parametrize("name", FIXTURES)
def test_fixture_end_to_end(api, golden, name):
    rgb = load_input(name)
    p = name + "/"
    r = api.process_image(rgb)
    import hashlib
    sha = lambda a: hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert sha(r.grey_image_np) == str(golden[p + "grey_sha"]), "grey"
    assert sha(r.edge_detected_image_np) == str(golden[p + "edges_sha"]), "edges"
    assert_same(r.circles, golden[p + "circles"].reshape(-1, 3), "stacked circles")
    assert sha(r.circles_removed_image_np) == str(golden[p + "masked_sha"]), "masked"
    assert_same(np.asarray(r.hlines, np.float32).reshape(-1), golden[p + "hlines"], "hlines")
    assert_same(np.asarray(r.vlines, np.float32).reshape(-1), golden[p + "vlines"], "vlines")
    assert r.valid_grid == bool(golden[p + "valid"])
    if r.valid_grid:
        assert [r.hsize, r.vsize] == list(golden[p + "sizes"])
        assert [r.hspace, r.vspace] == list(golden[p + "spaces"])
        assert_same(r.hcentres_complete, golden[p + "hcentres_complete"], "hcentres_complete")
        assert_same(r.vcentres_complete, golden[p + "vcentres_complete"], "vcentres_complete")
    assert r.board_ready == bool(golden[p + "board_ready"])
    if r.board_ready:
        assert_same(r.detected_board.astype(np.uint8), golden[p + "board"], "board")
        assert r.num_black_stones == int((golden[p + "board"] == 1).sum())
        assert r.num_white_stones == int((golden[p + "board"] == 2).sum())


@pytest.mark.parametrize("name", ["ex1", "ex7", "ex9"])
def test_fixture_stage_entry_points(api, oracle, golden, name):
    """The reference-named entry points chained by hand give the same result as the pipeline."""
    rgb = load_input(name)
    p = name + "/"
    grey = api.grey_image(rgb)
    edges = api.edge_map(rgb)
    circles, masked = api.find_circles(grey, edges)
    assert_same(circles, golden[p + "circles"].reshape(-1, 3), "find_circles")
    thr = int(golden[p + "threshold"])
    hl, vl = api.find_lines(masked, thr, api.Direction.H), api.find_lines(masked, thr, api.Direction.V)
    hc, vc = api.cluster(hl), api.cluster(vl)
    assert_same(np.asarray(hc), golden[p + "hcentres"], "hcentres")
    assert_same(np.asarray(vc), golden[p + "vcentres"], "vcentres")
    valid, kept, vsize, hsize, hcc, vcc, hspace, vspace = api.validate_grid(hc, vc, circles)
    assert valid == bool(golden[p + "valid"])
    if valid and hsize <= 19 and vsize <= 19:
        board, br = api.classify_stones(grey, kept, hcc, vcc, hspace, vspace, 128)
        assert_same(board.astype(np.uint8), golden[p + "board"], "board")
        assert_same(br, golden[p + "brightness"], "stone_brightnesses")
        # black-threshold drag (img2sgf.py:762-765): classification alone with another threshold
        board2, _ = api.classify_stones(grey, kept, hcc, vcc, hspace, vspace, 60)
        want = np.zeros_like(board2)
        k = 0
        gb = golden[p + "board"]
        for i in range(hsize):
            for j in range(vsize):
                if gb[i, j]:
                    want[i, j] = 1 if golden[p + "brightness"][k] <= 60 else 2
                    k += 1
        assert_same(board2, want, "board at threshold 60")


# ------------------------------------------------------------------ batches, synthetic configs
def _synth_rgb(config, start, count, **kw):
    from img2sgf_b200 import synth
    g, t = synth.batch(config, start, count, **kw)
    return np.ascontiguousarray(np.repeat(g[..., None], 3, axis=-1)), t


def test_batch_equals_oracle_1024(api, oracle):
    """BASELINE.json configs[3] shape (1024x1024, threshold 150), a few seeds: the batched run
    matches the oracle image by image and recovers the generator's ground truth."""
    import torch
    from img2sgf_b200 import batch as B, synth
    rgb, truth = _synth_rgb("synth1024", 0, 5)
    rgb[4] = _synth_rgb("synth1024", 4, 1, noise=2.0, numbered=True)[0][0]      # config 5 flavour
    runner = B.BatchRunner(1024, 1024, chunk=2)
    rec = B.run_with_retry(runner, torch.from_numpy(rgb).cuda(), 150)
    for i in range(5):
        res, circles, _ = oracle.pipeline(rgb[i], 150)
        assert rec[i]["status"] == 0
        assert rec[i]["n_circles"] == res.n_circles, f"image {i}"
        assert bool(rec[i]["board_ready"]) == bool(res.board_ready)
        want = np.zeros((19, 19), np.uint8)
        if res.board_ready:
            b = oracle.board_of(res)
            want[:b.shape[0], :b.shape[1]] = b
        assert_same(rec[i]["board"].reshape(19, 19), want, f"board {i}")
        assert (rec[i]["n_black"], rec[i]["n_white"]) == (res.n_black, res.n_white)
        if i < 4:
            assert_same(rec[i]["board"].reshape(19, 19), truth[i], f"truth {i}")


def test_multistream_runner_equals_single_stream(api, oracle):
    """bench.py's default runner alternates chunks between several CUDA streams (one engine and
    workspace per stream): the records must not depend on it.  Noisy, numbered stones (config 5
    flavour) so that hysteresis has real weak-pixel work in every map."""
    import torch
    from img2sgf_b200 import batch as B, synth
    imgs = [synth.to_rgb(synth.diagram(512, 24, 11, seed=40 + k, noise=2.5 if k % 2 else 0.0, numbered=(k % 3 == 0))[0])
            for k in range(7)]
    rgb = np.stack(imgs)
    dev = torch.from_numpy(rgb).cuda()
    one = B.records_to_numpy(B.BatchRunner(512, 512, chunk=2, streams=1).run(dev, 60))
    many = B.records_to_numpy(B.BatchRunner(512, 512, chunk=2, streams=3).run(dev, 60))
    torch.cuda.synchronize()
    assert one.tobytes() == many.tobytes()
    for i in (0, 1, 3):
        res, circles, _ = oracle.pipeline(rgb[i], 60)
        assert one[i]["status"] == 0 and one[i]["n_circles"] == res.n_circles
        want = np.zeros((19, 19), np.uint8)
        if res.board_ready:
            b = oracle.board_of(res)
            want[:b.shape[0], :b.shape[1]] = b
        assert_same(one[i]["board"].reshape(19, 19), want, f"board {i}")


def test_full_size_properties_2048(api):
    """BASELINE.json configs[2] shape (2048x2048, s=60, r=28, threshold 176): size-independent
    properties instead of an oracle run -- the generator's ground truth is recovered exactly
    (the reference does), batch == single-image, results independent of chunking."""
    import torch
    from img2sgf_b200 import batch as B
    rgb, truth = _synth_rgb("synth2048", 100, 4)
    dev = torch.from_numpy(rgb).cuda()
    rec_a = B.run_with_retry(B.BatchRunner(2048, 2048, chunk=4), dev, 176)
    rec_b = B.run_with_retry(B.BatchRunner(2048, 2048, chunk=1), dev, 176)
    assert rec_a.tobytes() == rec_b.tobytes()
    for i in range(4):
        assert rec_a[i]["board_ready"] and (rec_a[i]["hsize"], rec_a[i]["vsize"]) == (19, 19)
        assert_same(rec_a[i]["board"].reshape(19, 19), truth[i], f"truth {i}")
    single = api.process_image(rgb[2], 176)
    assert single.record.tobytes() == rec_a[2].tobytes()
    # idempotence of masking: masking the masked image with the same circles changes nothing
    m2 = api.mask_circles(single.circles_removed_image_np, single.circles)
    assert_same(m2, single.circles_removed_image_np, "mask idempotence")


def test_ragged_batch_no_circles_and_blank(api, oracle):
    """Edge cases the reference guards: nothing found anywhere (no_circles.jpg, img2sgf.py:180-181),
    blank image, tiny image."""
    for rgb in (load_input("no_circles"), np.full((64, 80, 3), 255, np.uint8), np.zeros((3, 5, 3), np.uint8)):
        r = api.process_image(rgb)
        res, circles, masked = oracle.pipeline(rgb)
        assert len(r.circles) == res.n_circles
        assert_same(r.circles_removed_image_np, masked, "masked")
        assert r.valid_grid == bool(res.grid.valid) and r.board_ready == bool(res.board_ready)
        assert len(r.hlines) == res.n_hlines and len(r.vlines) == res.n_vlines
