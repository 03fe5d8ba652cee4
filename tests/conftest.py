import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """{case: {field: array}} from tests/golden/<name>.npz"""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for key in z.files:
        case, field = key.split("/")
        out.setdefault(case, {})[field] = z[key]
    return out


@pytest.fixture(scope="session")
def orb_golden():
    return load_golden("orb_cases")


@pytest.fixture(scope="session")
def sift_golden():
    return load_golden("sift_cases")


@pytest.fixture(scope="session")
def orb_set_golden():
    return load_golden("orb_set_3x1024")


@pytest.fixture(scope="session")
def matcher():
    """One FeatureMatcherGpu for the GPU session; fails loudly if the CUDA library is missing."""
    import eacham_b200
    m = eacham_b200.FeatureMatcherGpu(0.8)
    yield m
    m.close()
