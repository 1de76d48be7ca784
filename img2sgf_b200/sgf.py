"""Batched SGF emission from the per-image result records.

Host-side mirror of the reference's output stage: `align_board` (img2sgf.py:484-494), the
side-to-move guess at the end of `identify_board` (:528-534) and `to_SGF` (:781-810), with the Tk
variable `side_to_move` and the global `board_alignment` turned into arguments.  Pure string work
on the 384-byte records the GPU path returns -- no image data is touched here.
"""
from __future__ import annotations

import string

import numpy as np

BOARD_SIZE = 19
EMPTY, BLACK, WHITE = 0, 1, 2
TOP, BOTTOM, LEFT, RIGHT = range(4)          # Alignment, img2sgf.py:83-84
DEFAULT_ALIGNMENT = (LEFT, TOP)              # img2sgf.py:627
_LETTERS = string.ascii_lowercase


def align_board(part, alignment=DEFAULT_ALIGNMENT) -> np.ndarray:
    """Place a (hsize, vsize) part board in a 19x19 board; alignment = (LEFT|RIGHT, TOP|BOTTOM)."""
    part = np.asarray(part)
    hsize, vsize = part.shape
    full = np.zeros((BOARD_SIZE, BOARD_SIZE))
    xo = BOARD_SIZE - hsize if alignment[0] == RIGHT else 0
    yo = BOARD_SIZE - vsize if alignment[1] == BOTTOM else 0
    full[xo:xo + hsize, yo:yo + vsize] = part
    return full


def guess_side_to_move(num_black: int, num_white: int) -> int:
    """BLACK (1) when black has no more stones than white, else WHITE (2) -- img2sgf.py:528-534."""
    return BLACK if num_black <= num_white else WHITE


def _points(board, colour) -> str:
    ii, jj = np.nonzero(np.asarray(board) == colour)      # row-major: i outer, j inner, like the reference loops
    return "".join("[" + _LETTERS[i] + _LETTERS[j] + "]" for i, j in zip(ii.tolist(), jj.tolist()))


def to_sgf(full_board, side_to_move: int) -> str:
    """`to_SGF` (img2sgf.py:781-810): AB/AW lists in board order, the side to move listed first."""
    black = _points(full_board, BLACK)
    white = _points(full_board, WHITE)
    black = "AB" + black if black else ""
    white = "AW" + white if white else ""
    head = "(;GM[1]FF[4]SZ[" + str(BOARD_SIZE) + "]\n"
    if side_to_move == BLACK:
        return head + "PL[B]\n" + black + "\n" + white + "\n" + ")\n"
    return head + "PL[W]\n" + white + "\n" + black + "\n" + ")\n"


def records_to_sgf(records: np.ndarray, alignment=DEFAULT_ALIGNMENT, side_to_move=None) -> list:
    """One SGF string per record (None where no board was found).  `records` is the structured array
    of `_native.RECORD_DTYPE`; the record's board holds the part board at the top-left."""
    out = []
    for r in np.atleast_1d(records):
        if not r["board_ready"]:
            out.append(None)
            continue
        hs, vs = int(r["hsize"]), int(r["vsize"])
        part = r["board"].reshape(BOARD_SIZE, BOARD_SIZE)[:hs, :vs]
        stm = side_to_move if side_to_move is not None else guess_side_to_move(int(r["n_black"]), int(r["n_white"]))
        out.append(to_sgf(align_board(part, alignment), stm))
    return out
