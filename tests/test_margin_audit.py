"""SURVEY section 4, "margin audit": bit-exact split indices are a CHECKED precondition of the parity suites, not luck.

Two correct implementations of the reference's arithmetic (same IEEE operations in the same order) can only take
a different decision where their `log` differs -- glibc's on the CPU, CUDA's on the device, both within a couple of
ulp -- which moves a gain  n log V(s,e) - (n1 log V1 + n2 log V2)  by at most

    3 terms x n <= 1e4 samples x |log V| <= 10 x 4 ulp(2^-52)  ~=  3e-10.

The oracle records, for every window scan, how far the decision was from flipping: best gain minus the runner-up
(the second best candidate or min_gain, whichever is closer) for scans that split, min_gain minus the best gain
for scans that did not.  Every suite whose GPU result is compared bit for bit must keep its smallest margin well
above that bound.  (test_min_gain_exactly_at_a_decision_boundary puts min_gain ON a gain on purpose and is not part
of this audit; it checks the strict comparison itself.)"""
import numpy as np
import pytest

import oracle
from conftest import RULES_1000
from pypore_b200 import synth

LOG_ROUNDING_BOUND = 3e-10
REQUIRED = 10 * LOG_ROUNDING_BOUND


def audit(arrays, **kw):
    split, nosplit, scans = np.inf, np.inf, 0
    for a in arrays:
        s, n, c = oracle.margin_audit(np.asarray(a, np.float64), **kw)
        split, nosplit, scans = min(split, s), min(nosplit, n), scans + c
    return split, nosplit, scans


def events_of(x):
    x = x.astype(np.float64)
    s, l = oracle.events(x, 110, RULES_1000)
    return [x[a:a + n] for a, n in zip(s, l)]


@pytest.mark.parametrize("tier", ["A", "B"])
@pytest.mark.parametrize("kw", [dict(), dict(prior_segments_per_second=10),
                                dict(min_width=50, max_width=2500, window_width=1000, prior_segments_per_second=50),
                                dict(min_width=3, max_width=40, window_width=64),
                                dict(min_gain_per_sample=0.01, window_width=2000)])
def test_margins_of_the_split_suites(tier, kw):
    """The traces of tests/test_gpu_parity.py::test_split_events_bit_exact and of the C1 pipeline tests."""
    for n_events, seed in ((24, 31), (500, 0)):
        if n_events == 500 and kw not in (dict(), dict(prior_segments_per_second=10)):
            continue
        split, nosplit, scans = audit(events_of(synth.make_trace(n_events, seed=seed, tier=tier)), **kw)
        assert scans > 0 and split >= REQUIRED and nosplit >= REQUIRED, (n_events, split, nosplit)


def test_margins_of_the_bench_workload_and_long_events():
    """BASELINE configs[1] at full size (the trace bench.py hashes against the reference) and the long-event fixture."""
    split, nosplit, scans = audit(events_of(synth.make_trace(5000, seed=1, tier="A")))
    assert scans > 300000 and split >= REQUIRED and nosplit >= REQUIRED, (split, nosplit)
    x = synth.make_long_event(300000, seed=100, tier="A")
    for kw in (dict(min_width=100, max_width=20000, window_width=10000),
               dict(min_width=100, max_width=20000, window_width=10000, prior_segments_per_second=10),
               dict(min_width=100, max_width=15000, window_width=4000, min_gain_per_sample=2.0)):
        split, nosplit, scans = audit([x], **kw)
        assert split >= REQUIRED and nosplit >= REQUIRED, (kw, split, nosplit)
