import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
FIXTURES = [f"ex{i}" for i in range(1, 18)] + ["no_circles"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_input(name: str) -> np.ndarray:
    """Contrast-enhanced RGB array of a reference fixture (tests/golden/inputs, see make_golden.py)."""
    from PIL import Image
    a = np.array(Image.open(os.path.join(GOLDEN_DIR, "inputs", name + ".png")))
    if a.ndim == 2:
        a = np.repeat(a[..., None], 3, axis=-1)
    return np.ascontiguousarray(a[..., :3], np.uint8)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "golden.npz"))


@pytest.fixture(scope="session")
def random_vectors():
    return np.load(os.path.join(GOLDEN_DIR, "random_vectors.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O
