"""CPU-side checks: the C-ABI library loads and exports every symbol include/img2sgf_b200.h
declares (no compute calls), struct layouts match, host logic (sharding, record gather over
gloo with world_size 2, thresholds, generator) behaves."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from img2sgf_b200 import build
    return build.build()


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "img2sgf_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(i2s_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = C.CDLL(built)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    from img2sgf_b200 import _native as N
    assert sorted(N.EXPORTS) == declared
    N.lib()     # sets argtypes for all of them


def test_struct_layouts(built):
    from img2sgf_b200 import _native as N
    assert N.RECORD_DTYPE.itemsize == 384
    assert N.RECORD_DTYPE.fields["status"][1] == 380
    assert N.GRID_DTYPE.fields["hcentres"][1] == 32
    lim = N.default_limits()
    assert lim.cand_cap >= 1024 and lim.circle_cap >= 1024 and lim.line_cap >= 256 and lim.hyst_passes >= 1
    assert N.lib().i2s_version() >= 100


def test_bad_arguments_fail_loudly(built):
    """Argument validation happens before any CUDA call, so it can be exercised without a GPU."""
    from img2sgf_b200 import _native as N
    rc = N.lib().i2s_grey(None, 0, None, 0, 1, 10, 10, None)
    assert rc == -1 and b"bad argument" in N.lib().i2s_last_error()
    with pytest.raises(N.NativeError):
        N.check(rc, "i2s_grey")
    lim = N.default_limits()
    assert N.lib().i2s_pipeline_workspace_bytes(4, 512, 512, C.byref(lim)) > 4 * 512 * 512 * 40
    # the masking kernel keeps 1 + circle index in 16 bits: a larger capacity is refused, not truncated
    buf = (C.c_uint8 * 64)()
    fl = (C.c_float * 16)()
    cnt = (C.c_int32 * 1)(0)
    rc = N.lib().i2s_mask_circles(buf, buf, 0, 1, 8, 8, fl, cnt, 65535, None)
    assert rc == -1 and b"bad argument" in N.lib().i2s_last_error()
    lim.circle_cap = 70000
    ws = (C.c_uint8 * 64)()
    st = (C.c_int32 * 1)()
    rc = N.lib().i2s_hough_circles(buf, 0, 1, 8, 8, fl, cnt, st, C.byref(lim), ws, 64, None)
    assert rc == -1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from img2sgf_b200 import api, _native as N
    with pytest.raises(N.NativeError):
        api.edge_map(np.zeros((16, 16, 3), np.uint8))
    with pytest.raises(N.NativeError):
        api.process_image(np.zeros((16, 16, 3), np.uint8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "img2sgf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f
                assert "import cv2" not in src and "cv2." not in src, f


def test_shard_range():
    from img2sgf_b200.batch import shard_range
    for total in (0, 1, 7, 8, 17, 8192):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_pixels_and_groups():
    from img2sgf_b200.batch import shard_by_pixels, group_by_size
    rng = np.random.default_rng(0)
    sizes = [(int(h), int(w)) for h, w in rng.integers(100, 1300, (40, 2))]
    for world in (1, 2, 3, 8):
        parts = shard_by_pixels(sizes, world)
        assert sorted(i for p in parts for i in p) == list(range(40))
        loads = [sum(sizes[i][0] * sizes[i][1] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(h * w for h, w in sizes)        # LPT bound
        assert parts == shard_by_pixels(sizes, world)                        # deterministic
    groups = group_by_size(sizes, max_group=8)
    assert sorted(i for g in groups for i in g) == list(range(40)) and max(len(g) for g in groups) <= 8
    for g in groups:
        ch, cw = max(sizes[i][0] for i in g), max(sizes[i][1] for i in g)
        assert sum(sizes[i][0] * sizes[i][1] for i in g) >= 0.6 * ch * cw * len(g) or len(g) == 1
    assert group_by_size([]) == [] and shard_by_pixels([], 2) == [[], []]


def test_ragged_pack_layout():
    """The packed layout RaggedRunner uploads: descriptors first, 16-byte aligned image starts,
    4-byte aligned row pitches, pixels where the descriptors say."""
    from img2sgf_b200.batch import RaggedRunner
    from img2sgf_b200 import _native as N
    rng = np.random.default_rng(1)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in ((5, 7), (9, 3), (4, 10))]
    nbytes, desc, ch, fill = RaggedRunner.pack(imgs, [2, 0, 1], thresholds=[11, 12, 13])
    buf = np.zeros(nbytes, np.uint8)
    fill(buf)
    assert ch == 3 and desc.dtype == N.IMAGE_DTYPE and list(desc["line_threshold"]) == [13, 11, 12]
    back = buf[:desc.nbytes].view(N.IMAGE_DTYPE)
    assert (back == desc).all()
    for k, i in enumerate([2, 0, 1]):
        h, w = imgs[i].shape[:2]
        o, p = int(desc[k]["offset"]), int(desc[k]["pitch"])
        assert o % 16 == 0 and p % 4 == 0 and p >= 3 * w and (desc[k]["h"], desc[k]["w"]) == (h, w)
        assert (buf[o:o + h * p].reshape(h, p)[:, :3 * w] == imgs[i].reshape(h, 3 * w)).all()


def test_choose_threshold():
    from img2sgf_b200.api import choose_threshold
    assert choose_threshold(750, 747) == 74 and choose_threshold(110, 102) == 23
    assert choose_threshold(10, 10) == 20 and choose_threshold(4000, 4000) == 200
    assert choose_threshold(2048, 2048) == 176


def test_synth_deterministic():
    from img2sgf_b200 import synth
    a, ta = synth.diagram(256, 12, 5, seed=7)
    b, tb = synth.diagram(256, 12, 5, seed=7)
    assert (a == b).all() and (ta == tb).all() and a.dtype == np.uint8 and set(np.unique(ta)) <= {0, 1, 2}
    imgs, truths = synth.batch("synth1024", 3, 2)
    assert imgs.shape == (2, 1024, 1024) and truths.shape == (2, 19, 19)


_GLOO_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from img2sgf_b200.batch import shard_range, gather_records, RECORD_BYTES
rank, world, total = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(sys.argv[2])
dist.init_process_group("gloo", rank=rank, world_size=world)
s, e = shard_range(total, rank, world)
local = torch.zeros((e - s, RECORD_BYTES), dtype=torch.uint8)
for i in range(s, e):
    local[i - s] = torch.from_numpy(np.random.default_rng(i).integers(0, 256, RECORD_BYTES, dtype=np.uint8))
full = gather_records(local, total)
# ragged batches: pixel-balanced assignment, records come back in input order on every rank
from img2sgf_b200.batch import shard_by_pixels, gather_ragged_records
from img2sgf_b200 import _native as N
sizes = [(10 + 7 * i, 20 + 3 * i) for i in range(total)]
assign = shard_by_pixels(sizes, world)
mine = np.zeros(len(assign[rank]), N.RECORD_DTYPE)
for k, i in enumerate(assign[rank]):
    mine[k]["n_circles"] = 1000 + i
    mine[k]["board"][:] = i % 3
allrec = gather_ragged_records(mine, assign)
assert list(allrec["n_circles"]) == [1000 + i for i in range(total)], allrec["n_circles"]
assert all((allrec[i]["board"] == i % 3).all() for i in range(total))
want = torch.stack([torch.from_numpy(np.random.default_rng(i).integers(0, 256, RECORD_BYTES, dtype=np.uint8))
                    for i in range(total)]) if total else torch.zeros((0, RECORD_BYTES), dtype=torch.uint8)
assert full.shape == want.shape and bool((full == want).all()), (rank, full.shape)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.parametrize("total", [5, 8])
def test_gather_records_gloo_world2(tmp_path, total):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    port = 29500 + os.getpid() % 1000 + total
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, str(total)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out
        assert "ok" in out


def test_find_lines_pair_runs_one_pass(monkeypatch):
    """find_lines(t, H) followed by find_lines(t, V) on the same image (img2sgf.py:259-261) costs one pass:
    host logic only, the device call is stubbed."""
    from img2sgf_b200 import api
    calls = []

    def fake_retrying(fn, lim=None):
        calls.append(1)
        return np.array([10.0, 20.0], np.float32), np.array([5.0], np.float32)

    monkeypatch.setattr(api, "_require_cuda", lambda: None)
    monkeypatch.setattr(api, "_dev", lambda a, dt=None: a)
    monkeypatch.setattr(api, "_retrying", fake_retrying)
    monkeypatch.setattr(api, "_lines_memo", None)
    img = np.zeros((32, 48), np.uint8)
    img[5, :] = 255
    h = api.find_lines(img, 40, api.Direction.H)
    v = api.find_lines(img, 40, api.Direction.V)
    assert len(calls) == 1 and h.shape == (2, 1) and v.shape == (1, 1)
    h[0, 0] = -1.0                                             # callers own what they get
    assert api.find_lines(img, 40, api.Direction.H)[0, 0] == 10.0 and len(calls) == 1
    api.find_lines(img, 41, api.Direction.H)                   # another threshold: a new pass
    img2 = img.copy(); img2[6, 3] = 255
    api.find_lines(img2, 41, api.Direction.H)                  # another image: a new pass
    assert len(calls) == 3


def test_failed_images_flags_overflows():
    from img2sgf_b200 import batch as B, _native as N
    rec = np.zeros(6, N.RECORD_DTYPE)
    rec["status"][1] = 1          # I2S_ST_CAND_OVERFLOW
    rec["status"][3] = 8          # I2S_ST_HYST_NOT_CONVERGED
    rec["status"][4] = N.ST_GRID_OVERFLOW
    assert list(B.failed_images(rec)) == [1, 3, 4]
    assert B.RETRY_BITS & N.ST_GRID_OVERFLOW == 0          # a larger limit cannot help a grid that does not fit
