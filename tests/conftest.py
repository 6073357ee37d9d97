import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def tables():
    with open(os.path.join(GOLDEN, "tables.json")) as f:
        return json.load(f)


def load_npz(name):
    z = np.load(os.path.join(GOLDEN, name))
    groups = {}
    for k in z.files:
        if "/" in k:
            g, rest = k.split("/", 1)
            groups.setdefault(g, {})[rest] = z[k]
        else:
            groups[k] = z[k]
    return groups


@pytest.fixture(scope="session")
def small_pair():
    return load_npz("small_pair.npz")


@pytest.fixture(scope="session")
def uni_pair():
    return load_npz("uni_pair.npz")


@pytest.fixture(scope="session")
def cfg1_seeded():
    return load_npz("cfg1_seeded.npz")


def as_lpl(raw):
    return [[tuple(t) for t in layer] for layer in raw]


def rel_err(a, b):
    """max |a-b| / max |b| (scale-relative max error)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="session")
def cdan_small():
    return np.load(os.path.join(GOLDEN, "cdan_small.npz"))
