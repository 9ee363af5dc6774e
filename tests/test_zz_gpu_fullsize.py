"""Parity at BASELINE's FULL long-event size (configs[3]: events of 10 M samples, max_width = 1e6, the spine kernel
on thread-block clusters followed by the work queue): 4 events in one trace.  The oracle result is too big to
store, so tests/golden/c4_full.npz (written by the REAL reference, tests/golden/make_golden.py --c4-only) holds
the event table, the segment counts and a SHA-256 of the (event, start, end) rows; the device tables must hash to
the same value.  The candidate counts are the oracle's (tests/test_oracle_golden.py checks those on the CPU)."""
import numpy as np
import pytest

from conftest import load_golden, sha
from pypore_b200 import synth
from pypore_b200.parsers import statsplit_min_gain

pytestmark = pytest.mark.gpu

C4_CANDIDATES = {"default": 300443705, "psps10": 188617127}
C4_SETTINGS = {"default": dict(min_width=100, max_width=1000000, window_width=10000),
               "psps10": dict(min_width=100, max_width=1000000, window_width=10000, prior_segments_per_second=10)}


def test_c4_full_size_tables_hash_to_the_references(ctx):
    g = load_golden("c4_full.npz")
    x = synth.make_long_trace(4, 10_000_000, seed0=100, tier="A")
    rules = dict(rule_mask=7, duration_gt=1000, duration_lt=0, min_gt=-0.5, max_lt=110.0)
    ctx.upload_trace(x)
    for name, kw in C4_SETTINGS.items():
        mw, MW, W, gain = statsplit_min_gain(**kw)
        r = ctx.pipeline(110.0, min_width=mw, max_width=MW, window_width=W, min_gain=gain, **rules)
        es, el = ctx.events(r["events"])
        assert np.array_equal(es, g["ev_start"]) and np.array_equal(el, g["ev_len"])
        seg = ctx.segments(r["segments"])
        rows = np.stack([seg["event"].astype(np.int64), seg["start"], seg["end"]], axis=1)
        assert r["segments"] == int(g[name + "_segments"])
        # size-independent properties first (they say more than a hash when something is off)
        for e in range(4):
            sel = rows[:, 0] == e
            assert rows[sel, 1][0] == 0 and rows[sel, 2][-1] == el[e]
            assert np.array_equal(rows[sel, 1][1:], rows[sel, 2][:-1])
            w = rows[sel, 2] - rows[sel, 1]
            assert w.min() >= mw and w.max() <= MW
        assert ctx.split_counters()["candidates"] == C4_CANDIDATES[name]
        assert sha(rows) == str(g[name + "_sha"])
        # statistics of a sample of the segments against numpy on the same samples (1e-9)
        pick = np.linspace(0, len(rows) - 1, 64).astype(int)
        for k in pick:
            v = x[es[rows[k, 0]] + rows[k, 1]:es[rows[k, 0]] + rows[k, 2]].astype(np.float64)
            assert abs(seg["mean"][k] - v.mean()) <= 1e-9 * abs(v.mean())
            assert abs(seg["std"][k] - v.std()) <= 1e-9 * v.std()
            assert seg["min"][k] == v.min() and seg["max"][k] == v.max()
