"""The median kernel's bit-sliced saturated-window verdicts (img2sgf_b200/csrc/median_cores.cuh) run on the CPU
against brute-force window counts -- catches a logic slip before GPU time is spent."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_settle_words_equal_bruteforce(tmp_path):
    so = str(tmp_path / "libmedian_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "host", "median_host.cpp")])
    lib = C.CDLL(so)
    assert lib.mh_check(6) == 0
