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


# FastStatSplit.best_single_split / score_samples fixture cases, shared by tests/golden/make_golden.py
# (which runs the real reference) and the tests: name -> (length, seed, tier, FastStatSplit kwargs)
SCORING_CASES = {
    "psps10": (2500, 11, "A", dict(min_width=100, window_width=10000, prior_segments_per_second=10)),
    "default_tierB": (1500, 12, "B", dict(min_width=100, window_width=10000)),
    "narrow_forced": (6000, 13, "A", dict(min_width=50, max_width=2500, window_width=1000,
                                          prior_segments_per_second=50)),
}


def sha(a):
    import hashlib
    import numpy as np
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
