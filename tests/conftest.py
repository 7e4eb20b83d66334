"""pytest configuration: registers the `gpu` marker and shared golden-fixture helpers."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """Returns (meta dict, npz) for tests/golden/<name>.npz (written by oracle/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"])) if "meta" in z.files else None
    return meta, z


def rel_l2(a, b):
    import torch
    a = a.double().reshape(-1)
    b = b.double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="session")
def known_answers():
    return np.load(os.path.join(GOLDEN_DIR, "known_answers.npz"))
