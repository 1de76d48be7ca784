"""GPU parity tests of the batch drivers and of the round-2 additions: ragged batches (per-image sizes,
pitched planes), greyscale sources, the contrast/brightness prologue, pitched stage calls, the
headline configurations against the oracle at scale, and the engineered edge cases of
closest_index / average_intensity.  Tolerance 0 everywhere."""
import ctypes as C
import hashlib

import numpy as np
import pytest

from conftest import FIXTURES, load_input

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from img2sgf_b200 import api as A, build
    build.build()
    return A


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def want_board(oracle, res):
    want = np.zeros((19, 19), np.uint8)
    if res.board_ready:
        b = oracle.board_of(res)
        want[:b.shape[0], :b.shape[1]] = b
    return want


def check_record(oracle, rec, rgb, thr, what):
    res, circles, _ = oracle.pipeline(rgb, thr)
    assert rec["status"] == 0, what
    assert rec["n_circles"] == res.n_circles, f"{what}: circles {rec['n_circles']} vs {res.n_circles}"
    assert bool(rec["valid"]) == bool(res.grid.valid) and bool(rec["board_ready"]) == bool(res.board_ready), what
    assert np.array_equal(rec["board"].reshape(19, 19), want_board(oracle, res)), f"{what}: board"
    assert (rec["n_black"], rec["n_white"]) == (res.n_black, res.n_white), what


# ------------------------------------------------------------------ ragged batches
def _fixture_images(grey_sources_as_planes):
    imgs = []
    for name in FIXTURES:
        a = load_input(name)
        is_grey = (a[..., 0] == a[..., 1]).all() and (a[..., 1] == a[..., 2]).all()
        imgs.append(np.ascontiguousarray(a[..., 0]) if (is_grey and grey_sources_as_planes) else a)
    return imgs


@pytest.mark.parametrize("planes", [False, True])
def test_ragged_batch_of_fixtures_equals_goldens(api, golden, planes):
    """BASELINE.json configs[1]: the reference's test images (110x102 .. 1265x1245, no width a
    multiple of 16) as ONE ragged batch: every record equals the per-image goldens.  With
    planes=True the mode-"L" sources go in as single planes (a third of the bytes)."""
    from img2sgf_b200 import batch as B
    imgs = _fixture_images(planes)
    rec = B.RaggedRunner(streams=2).process_images(imgs)
    for k, name in enumerate(FIXTURES):
        p = name + "/"
        r = rec[k]
        assert r["status"] == 0, name
        assert r["n_circles"] == len(golden[p + "circles"].reshape(-1, 3)), name
        assert bool(r["valid"]) == bool(golden[p + "valid"]), name
        assert bool(r["board_ready"]) == bool(golden[p + "board_ready"]), name
        if r["valid"]:
            assert [r["hsize"], r["vsize"]] == list(golden[p + "sizes"]), name
        if r["board_ready"]:
            gb = golden[p + "board"]
            want = np.zeros((19, 19), np.uint8)
            want[:gb.shape[0], :gb.shape[1]] = gb
            assert np.array_equal(r["board"].reshape(19, 19), want), name
            assert (r["n_black"], r["n_white"]) == (int((gb == 1).sum()), int((gb == 2).sum())), name


def test_ragged_equals_single_and_order(api):
    """Records do not depend on grouping, group size, stream count or the order of the images."""
    from img2sgf_b200 import batch as B, synth
    rng = np.random.default_rng(3)
    imgs = []
    for k in range(11):
        size = int(rng.integers(180, 520))
        s = size // 22
        g, _ = synth.diagram(size, s, max(3, int(0.47 * s)), seed=200 + k, noise=2.0 if k % 3 == 0 else 0.0)
        crop = g[:size - int(rng.integers(0, 40)), :size - int(rng.integers(0, 40))]
        imgs.append(synth.to_rgb(np.ascontiguousarray(crop)))
    one = B.RaggedRunner(streams=1, max_group=3).process_images(imgs)
    two = B.RaggedRunner(streams=3, max_group=32).process_images(imgs[::-1])[::-1]
    assert one.tobytes() == two.tobytes()
    for k in (0, 4, 10):
        single = api.process_image(imgs[k])
        assert single.record.tobytes() == one[k].tobytes(), k


def test_grey_source_equals_rgb(api, oracle):
    """A mode-"L" source (R = G = B after convert('RGB'), img2sgf.py:651) processed as one plane gives
    byte-identical grey / edges / masked / circles / record."""
    rgb = load_input("ex1")
    assert (rgb[..., 0] == rgb[..., 1]).all() and (rgb[..., 1] == rgb[..., 2]).all()
    a = api.process_image(rgb)
    b = api.process_image(np.ascontiguousarray(rgb[..., 0]))
    assert np.array_equal(a.grey_image_np, b.grey_image_np) and np.array_equal(a.grey_image_np, rgb[..., 0])
    assert np.array_equal(a.edge_detected_image_np, b.edge_detected_image_np)
    assert np.array_equal(a.circles, b.circles) and np.array_equal(a.circles_removed_image_np, b.circles_removed_image_np)
    assert a.record.tobytes() == b.record.tobytes()
    assert np.array_equal(a.stone_brightnesses, b.stone_brightnesses)


# ------------------------------------------------------------------ prologue
@pytest.mark.parametrize("name", ["ex2", "ex9", "ex16", "no_circles"])
def test_enhance_equals_pil(api, name):
    """ImageEnhance.Contrast + Brightness (img2sgf.py:142-149) at three slider settings, against PIL itself."""
    from PIL import Image, ImageEnhance
    rgb = load_input(name)
    for cs, bs in ((70, 50), (35, 80), (95, 20)):
        fc, fb = api.scaled_contrast(cs), api.scaled_brightness(bs)
        want = np.array(ImageEnhance.Brightness(ImageEnhance.Contrast(Image.fromarray(rgb)).enhance(fc)).enhance(fb))
        got = api.enhance(rgb, fc, fb)
        assert np.array_equal(got, want), f"{name} contrast {cs} brightness {bs}: {int((got != want).sum())} bytes differ"


def test_pipeline_with_prologue(api):
    """process_image with slider values == process_image on the PIL-enhanced array."""
    from PIL import Image, ImageEnhance
    rgb = load_input("ex7")
    for cs, bs in ((70, 50), (60, 55)):
        enh = np.array(ImageEnhance.Brightness(ImageEnhance.Contrast(Image.fromarray(rgb)).enhance(
            api.scaled_contrast(cs))).enhance(api.scaled_brightness(bs)))
        a = api.process_image(rgb, contrast_slider=cs, brightness_slider=bs)
        b = api.process_image(enh)
        assert np.array_equal(a.grey_image_np, b.grey_image_np) and np.array_equal(a.circles, b.circles)
        assert a.record.tobytes() == b.record.tobytes()


# ------------------------------------------------------------------ pitched planes through the C ABI
def test_pitched_stage_calls(api, oracle):
    """Every stage entry point with an explicit pitch (rows padded, 16-byte aligned and deliberately
    odd) gives the tight-layout result."""
    import torch
    from img2sgf_b200 import _native as N
    lib = N.lib()
    rng = np.random.default_rng(9)
    h, w = 150, 203
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    f = rgb.astype(np.float32)
    f[1:-1] = (f[:-2] + f[1:-1] + f[2:]) / 3
    f[:, 1:-1] = (f[:, :-2] + f[:, 1:-1] + f[:, 2:]) / 3
    rgb = f.astype(np.uint8)
    grey = oracle.grey(rgb)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = lambda t: C.c_void_p(t.data_ptr())
    for pitch in (208, 256, 211):
        rp = 3 * w + (pitch - w)                         # RGB pitch with the same amount of padding
        drgb = torch.full((h, rp), 77, dtype=torch.uint8, device="cuda")
        drgb[:, :3 * w] = torch.from_numpy(rgb.reshape(h, 3 * w)).cuda()
        dgrey = torch.full((h, pitch), 99, dtype=torch.uint8, device="cuda")
        N.check(lib.i2s_grey(ptr(drgb), rp, ptr(dgrey), pitch, 1, h, w, st), "grey")
        assert np.array_equal(dgrey.cpu().numpy()[:, :w], grey), f"grey pitch {pitch}"
        outs = [torch.zeros((h, pitch), dtype=torch.uint8, device="cuda") for _ in range(3)]
        N.check(lib.i2s_gauss357(ptr(dgrey), ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), 1, h, w, pitch, st), "gauss")
        for b, o in zip((3, 5, 7), outs):
            assert np.array_equal(o.cpu().numpy()[:, :w], oracle.gauss(grey, b)), f"gauss{b} pitch {pitch}"
            m = torch.zeros((h, pitch), dtype=torch.uint8, device="cuda")
            N.check(lib.i2s_median(ptr(dgrey), ptr(m), 1, h, w, pitch, b, st), "median")
            assert np.array_equal(m.cpu().numpy()[:, :w], oracle.median(grey, b)), f"median{b} pitch {pitch}"
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        nb = lib.i2s_canny_workspace_bytes(1, h, w)
        ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
        e3 = torch.zeros((h, pitch), dtype=torch.uint8, device="cuda")
        N.check(lib.i2s_canny(ptr(drgb), 3, rp, ptr(e3), pitch, 1, h, w, 50, 200, 40, ptr(status), ptr(ws), nb, st), "canny3")
        assert np.array_equal(e3.cpu().numpy()[:, :w], oracle.canny_rgb(rgb)), f"canny rgb pitch {pitch}"
        e1 = torch.zeros((h, pitch), dtype=torch.uint8, device="cuda")
        N.check(lib.i2s_canny(ptr(dgrey), 1, pitch, ptr(e1), pitch, 1, h, w, 50, 100, 40, ptr(status), ptr(ws), nb, st), "canny1")
        assert np.array_equal(e1.cpu().numpy()[:, :w], oracle.canny_grey(grey)), f"canny grey pitch {pitch}"
        assert int(status.item()) == 0


# ------------------------------------------------------------------ the headline configurations at scale
def _synth_rgb(config, start, count, **kw):
    from img2sgf_b200 import synth
    g, t = synth.batch(config, start, count, **kw)
    return np.ascontiguousarray(np.repeat(g[..., None], 3, axis=-1)), t


def test_config3_64_seeds_vs_oracle(api, oracle):
    """BASELINE.json configs[3] (1024x1024, threshold 150): 64 seeds, every record against the oracle."""
    import torch
    from img2sgf_b200 import batch as B
    rgb, truth = _synth_rgb("synth1024", 1000, 64)
    rec = B.run_with_retry(B.BatchRunner(1024, 1024, chunk=16, streams=2), torch.from_numpy(rgb).cuda(), 150)
    for i in range(64):
        check_record(oracle, rec[i], rgb[i], 150, f"seed {1000 + i}")


def test_config2_8_images_vs_oracle(api, oracle):
    """BASELINE.json configs[2] (2048x2048, threshold 176): circles, masked image, lines and board of 8
    images against the oracle."""
    rgb, truth = _synth_rgb("synth2048", 500, 8)
    for i in range(8):
        res, circles, masked = oracle.pipeline(rgb[i], 176)
        r = api.process_image(rgb[i], 176)
        assert np.array_equal(r.circles, circles), f"circles {i}"
        assert sha(r.circles_removed_image_np) == sha(masked), f"masked {i}"
        assert len(r.hlines) == res.n_hlines and len(r.vlines) == res.n_vlines
        assert np.array_equal(r.record["board"].reshape(19, 19), want_board(oracle, res)), f"board {i}"
        assert np.array_equal(r.record["board"].reshape(19, 19), truth[i]), f"truth {i}"


def test_config5_noisy_numbered_vs_oracle(api, oracle):
    """BASELINE.json configs[4]: 64 numbered-stone diagrams with pixel noise (sigma 2): agreement with the
    oracle must be 100 % (the oracle itself may differ from the generator's truth)."""
    import torch
    from img2sgf_b200 import batch as B, synth
    imgs = np.stack([synth.to_rgb(synth.diagram(1024, 50, 24, seed=3000 + k, noise=2.0, numbered=True)[0]) for k in range(64)])
    rec = B.run_with_retry(B.BatchRunner(1024, 1024, chunk=16, streams=2), torch.from_numpy(imgs).cuda(), 150)
    for i in range(64):
        check_record(oracle, rec[i], imgs[i], 150, f"noisy numbered {i}")


def test_run_host_equals_run(api):
    """The pipelined host-buffer path (pinned host in, host records out) == the device-resident path."""
    import torch
    from img2sgf_b200 import batch as B
    rgb, _ = _synth_rgb("synth1024", 40, 13)
    runner = B.BatchRunner(1024, 1024, chunk=4, streams=3, copy_streams=2)
    dev = B.records_to_numpy(runner.run(torch.from_numpy(rgb).cuda(), 150))
    host = runner.run_host(torch.from_numpy(rgb).pin_memory(), 150).copy()
    again = runner.run_host(rgb, 150).copy()                 # numpy input: pinned on the fly
    assert dev.tobytes() == host.tobytes() == again.tobytes()


# ------------------------------------------------------------------ engineered edge cases
def test_closest_index_ties_and_empty_windows(api, oracle):
    """closest_index ties go to the LEFT line (img2sgf.py:459); an empty intensity window gives NaN and
    NaN counts as WHITE (:481, :541); circles far outside snap to the border intersections (:448-465)."""
    rng = np.random.default_rng(2)
    grey = rng.integers(0, 256, (200, 220), dtype=np.uint8)
    hc = np.array([20.0, 50.0, 80.0, 110.0, 140.0])          # y of horizontal lines
    vc = np.array([30.0, 60.0, 90.0, 120.0])                 # x of vertical lines
    circles = np.array([[45.0, 35.0, 12.0],                  # x tie between 30 and 60 -> 30; y tie 20|50 -> 20
                        [75.0, 65.0, 12.0],                  # ties again
                        [90.0, 110.0, 12.0],                 # exact hits
                        [500.0, -40.0, 12.0],                # far outside: snaps to (last x, first y)
                        [-10.0, 900.0, 12.0],                # far outside: (first x, last y)
                        [119.5, 139.5, 3.0]], np.float32)    # radius filtered out (:441-443)
    for hspace, vspace in ((30.0, 30.0), (0.4, 30.0), (30.0, 0.3), (44.0, 36.0)):
        g = oracle.Grid()
        g.valid, g.hsize, g.vsize, g.hspace, g.vspace = 1, len(vc), len(hc), hspace, vspace
        for k, v in enumerate(hc):
            g.hc[k] = v
        for k, v in enumerate(vc):
            g.vc[k] = v
        lo, hi = min(hspace, vspace) * 0.3, max(hspace, vspace) * 0.65
        kept = np.array([c for c in circles if lo < c[2] < hi], np.float32).reshape(-1, 3)
        want_b, want_br = oracle.classify(grey, kept, g, 128)
        got_b, got_br = api.classify_stones(grey, circles, hc, vc, hspace, vspace, 128)
        assert np.array_equal(got_b.astype(np.uint8), want_b), (hspace, vspace)
        assert np.array_equal(np.isnan(got_br), np.isnan(want_br)) and np.array_equal(np.nan_to_num(got_br), np.nan_to_num(want_br))
        if hspace < 1 or vspace < 1:
            assert np.isnan(got_br).all() and (got_b[got_b > 0] == 2).all()      # empty windows: NaN -> WHITE
    # a grid off the image: windows clipped to nothing
    got_b, got_br = api.classify_stones(grey, np.array([[400.0, 400.0, 12.0]], np.float32), np.array([380.0, 410.0]),
                                        np.array([390.0, 420.0]), 30.0, 30.0, 128)
    assert np.isnan(got_br).all() and got_b.sum() == 2


def test_candidate_capacity_16384(api, oracle):
    """The largest cand_cap the limits accept runs (working arrays move from shared memory to the
    workspace above 8192 candidates) and gives the reference's circles."""
    from img2sgf_b200 import synth, _native as N
    g, _ = synth.diagram(1024, 50, 24, seed=77, noise=3.0, numbered=True)
    want = oracle.hough_circles(g)
    for cap in (16384, 8192, 64):
        lim = N.Limits(cap, 4096, 1024, 5)
        got = api.hough_circles(g, limits=lim)
        assert np.array_equal(got, want), cap


def test_hysteresis_many_passes(api, oracle):
    """A pass budget beyond the 64-slot counter ring: the serpentine of weak edges crosses many tiles."""
    h, w = 700, 700
    img = np.full((h, w), 128, np.uint8)
    for y in range(20, 680, 24):
        img[y:y + 12, 10:690] = 146
    img[20:32, 10:40] = 255
    assert np.array_equal(api.canny_grey(img, 50, 100, hyst_passes=130), oracle.canny_grey(img, 50, 100))


# ------------------------------------------------------------------ crop as a view; records -> SGF
def test_selection_is_a_view(api, oracle):
    """crop_and_rotate_image (img2sgf.py:110-114) at rotation 0: process_image(selection=box) works on a view of
    the uploaded image (offset + full-image pitch) and equals processing the cropped copy and the oracle."""
    from img2sgf_b200 import synth
    g, _ = synth.diagram(700, 30, 14, seed=11, noise=1.5)
    rgb = synth.to_rgb(g)
    for box in ((37, 21, 660, 655), (0, 0, 700, 700), (101, 203, 452, 517)):
        x0, y0, x1, y1 = box
        crop = np.ascontiguousarray(rgb[y0:y1, x0:x1])
        a = api.process_image(rgb, selection=box)
        b = api.process_image(crop)
        assert a.record.tobytes() == b.record.tobytes(), box
        assert np.array_equal(a.circles, b.circles) and np.array_equal(a.circles_removed_image_np, b.circles_removed_image_np)
        res, circles, masked = oracle.pipeline(crop)
        assert np.array_equal(a.circles, circles) and np.array_equal(a.circles_removed_image_np, masked), box
    gv = api.process_image(np.ascontiguousarray(g), selection=(37, 21, 660, 655))       # greyscale source, cropped
    assert gv.record.tobytes() == api.process_image(rgb, selection=(37, 21, 660, 655)).record.tobytes()


def test_gpu_records_to_sgf(api, golden):
    """the records of the GPU path through align_board + to_SGF (img2sgf_b200/sgf.py) equal the
    reference's own output stage (replayed in oracle/ref_replay.py) applied to the golden boards."""
    from img2sgf_b200 import batch as B, sgf
    from oracle import ref_replay as R
    imgs = [load_input(n) for n in FIXTURES]
    rec = B.RaggedRunner().process_images(imgs)
    texts = sgf.records_to_sgf(rec)
    n = 0
    for k, name in enumerate(FIXTURES):
        if not bool(golden[name + "/board_ready"]):
            assert texts[k] is None, name
            continue
        part = golden[name + "/board"].astype(np.float64)
        hs, vs = part.shape
        stm = 1 if int((part == 1).sum()) <= int((part == 2).sum()) else 2        # img2sgf.py:528-534
        assert texts[k] == R.to_SGF(R.align_board(part, hs, vs), stm), name
        n += 1
    assert n >= 14
