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


C2_CANDIDATES = {"default": 240399270, "psps10": 135656885}
C2_SETTINGS = {"default": dict(min_width=100, max_width=1000000, window_width=10000),
               "psps10": dict(min_width=100, max_width=1000000, window_width=10000, prior_segments_per_second=10)}


def test_c2_full_size_tables_hash_to_the_references(ctx):
    """BASELINE configs[1] at full size -- exactly bench.py's workload (60 M samples, 5000 events): the event and
    segment tables, resident and streamed from host memory, hash to what the REAL reference produced
    (tests/golden/c2_full.npz, make_golden.py --c2-only), and the device counted the oracle's candidates."""
    g = load_golden("c2_full.npz")
    x = synth.make_trace(5000, seed=1, tier="A")
    assert len(x) == int(g["samples"])
    rules = dict(rule_mask=7, duration_gt=1000, duration_lt=0, min_gt=-0.5, max_lt=110.0)
    ctx.upload_trace(x)
    for name, kw in C2_SETTINGS.items():
        mw, MW, W, gain = statsplit_min_gain(**kw)
        r = ctx.pipeline(110.0, min_width=mw, max_width=MW, window_width=W, min_gain=gain, **rules)
        assert r["events"] == int(g["events"]) and r["event_samples"] == int(g["event_samples"])
        es, el = ctx.events(r["events"])
        assert sha(np.stack([es, el], axis=1)) == str(g["events_sha"])
        seg = ctx.segments(r["segments"], stats=False)
        rows = np.stack([seg["event"].astype(np.int64), seg["start"], seg["end"]], axis=1)
        assert r["segments"] == int(g[name + "_segments"])
        assert ctx.split_counters()["candidates"] == C2_CANDIDATES[name]
        assert sha(rows) == str(g[name + "_sha"])
    # the public host-memory call (what bench.py's e2e times), tables delivered to page-locked host memory
    mw, MW, W, gain = statsplit_min_gain(**C2_SETTINGS["default"])
    r = ctx.pipeline(110.0, min_width=mw, max_width=MW, window_width=W, min_gain=gain, host_trace=x, export=True,
                     **rules)
    es, el = r["event_table"]
    t = r["segment_table"]
    assert sha(np.stack([np.asarray(es, np.int64), np.asarray(el, np.int64)], axis=1)) == str(g["events_sha"])
    rows = np.stack([np.asarray(t["event"]).astype(np.int64), np.asarray(t["start"], np.int64),
                     np.asarray(t["end"], np.int64)], axis=1)
    assert sha(rows) == str(g["default_sha"])


def test_c5_full_size_files_through_the_batch_driver():
    """BASELINE configs[4] at full file size: eight 250 kHz files of 2.49 M samples, Event.filter(1, 2000) +
    SpeedyStatSplit per event, through batch.FileBatch (grouped passes and one file per pass).  The tables hash to the
    oracle's (tests/golden/c5_files.npz, made by tests/golden/make_c5_files.py) for the tutorial gain AND for the default
    gain min_gain = -0.0, where every window of filtered samples splits down to 2 * min_width."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_c5_files import SETTINGS, table_hashes
    from batch_common import FILTER, FS, detector
    from pypore_b200.batch import FileBatch
    from pypore_b200.parsers import SpeedyStatSplit
    g = load_golden("c5_files.npz")
    files = [synth.make_trace(208, seed=900 + i, tier="A") for i in range(8)]
    assert sum(len(f) for f in files) == int(g["samples"])
    for group_samples in (1 << 26, 0):
        b = FileBatch(device=0, workers=2, group_samples=group_samples)
        try:
            for name, kw in SETTINGS.items():
                t = b.parse(files, 1000. / FS, detector(), SpeedyStatSplit(**kw), FILTER)
                ne, ns, he, hs = table_hashes(t)
                assert (ne, ns) == (int(g[name + "_events"]), int(g[name + "_segments"])), name
                assert he == str(g[name + "_events_sha"]) and hs == str(g[name + "_segments_sha"]), name
        finally:
            b.close()
