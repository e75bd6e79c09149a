import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (build container only)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def relerr(a, b):
    """max |a-b| / max |b|  — the 1e-12 relative measure used throughout (north_star)."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    scale = np.abs(b).max() if b.size else 0.0
    if scale == 0.0:
        return float(np.abs(a - b).max()) if a.size else 0.0
    return float(np.abs(a - b).max() / scale)


def entrywise(a, b, rel=1e-9, floor=1e-13):
    """Entry-wise check next to the norm-wise relerr: max over entries of |a - b| / (rel |b| + floor max|b|); <= 1 passes.
    A small-magnitude block that is wrong by O(1) relative fails here even when max-abs / max-abs is tiny."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    if b.size == 0:
        return 0.0
    return float((np.abs(a - b) / (rel * np.abs(b) + floor * np.abs(b).max() + 1e-300)).max())


@pytest.fixture(scope="session")
def has_cuda():
    import torch

    return torch.cuda.is_available()
