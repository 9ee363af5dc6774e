import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


RULES_1000 = [lambda e: e.duration > 1000, lambda e: e.min > -0.5, lambda e: e.max < 110]


@pytest.fixture(scope="session")
def ctx():
    """One device context for the whole GPU session.  Fails loudly without the CUDA library / device."""
    from pypore_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()
