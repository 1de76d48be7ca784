"""SGF emission: img2sgf_b200/sgf.py against the replay of the reference's
align_board / to_SGF (oracle/ref_replay.py), on the golden fixture boards and random part boards."""
import numpy as np
import pytest

from conftest import FIXTURES
from img2sgf_b200 import sgf, _native as N
from oracle import ref_replay as R


def _cases(golden):
    for name in FIXTURES:
        key = name + "/board"
        if key in golden.files:
            yield name, golden[key]
    rng = np.random.default_rng(3)
    for k in range(20):
        hs, vs = int(rng.integers(1, 20)), int(rng.integers(1, 20))
        yield f"rand{k}", rng.integers(0, 3, (hs, vs)).astype(np.uint8)
    yield "empty", np.zeros((19, 19), np.uint8)
    yield "only_white", np.full((3, 4), 2, np.uint8)


def test_align_and_to_sgf_match_replay(golden):
    n = 0
    for name, part in _cases(golden):
        hs, vs = part.shape
        for a in ((sgf.LEFT, sgf.TOP), (sgf.RIGHT, sgf.TOP), (sgf.LEFT, sgf.BOTTOM), (sgf.RIGHT, sgf.BOTTOM)):
            full = sgf.align_board(part, a)
            want = R.align_board(part.astype(np.float64), hs, vs, a)
            assert np.array_equal(full, want), (name, a)
            for stm in (1, 2):
                assert sgf.to_sgf(full, stm) == R.to_SGF(want, stm), (name, a, stm)
        n += 1
    assert n > 20


def test_records_to_sgf(golden):
    recs = np.zeros(3, N.RECORD_DTYPE)
    part = golden["ex1/board"] if "ex1/board" in golden.files else np.eye(19, dtype=np.uint8)
    hs, vs = part.shape
    b = np.zeros((19, 19), np.uint8); b[:hs, :vs] = part
    recs[0]["board"] = b.reshape(-1); recs[0]["board_ready"] = 1; recs[0]["hsize"] = hs; recs[0]["vsize"] = vs
    recs[0]["n_black"] = int((part == 1).sum()); recs[0]["n_white"] = int((part == 2).sum())
    recs[2]["board"][:] = 0; recs[2]["board_ready"] = 1; recs[2]["hsize"] = 5; recs[2]["vsize"] = 7
    out = sgf.records_to_sgf(recs)
    assert out[1] is None
    stm = 1 if recs[0]["n_black"] <= recs[0]["n_white"] else 2
    assert out[0] == R.to_SGF(R.align_board(part.astype(np.float64), hs, vs), stm)
    assert out[2] == "(;GM[1]FF[4]SZ[19]\nPL[B]\n\n\n)\n"
