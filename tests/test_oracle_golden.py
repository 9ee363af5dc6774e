"""The oracle (CPU restatement) pinned against the reference itself.

* tests/golden/*.npz were produced by running the real Python reference
  (tests/golden/make_golden.py) in the build container;
* oracle/_ref holds the reference's cparsers.pyx compiled unmodified.
Nothing here needs a GPU.
"""
import hashlib
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, RULES_1000, load_golden
from pypore_b200 import synth


# same settings tests/golden/make_golden.py used
SPLIT_SETTINGS = {
    "default": dict(min_width=100, window_width=10000),
    "psps10": dict(min_width=100, window_width=10000, prior_segments_per_second=10),
    "narrow": dict(min_width=50, max_width=2500, window_width=1000, prior_segments_per_second=50),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["pipeline_tierA.npz", "pipeline_tierB.npz"])
def test_pipeline_matches_reference_fixture(name):
    g = load_golden(name)
    x32 = synth.make_trace(int(g["n_events"]), seed=int(g["seed"]), tier=str(g["tier"]))
    assert sha(x32) == str(g["input_sha256"]), "synthetic generator drifted from the fixture input"
    x = x32.astype(np.float64)
    second = float(g["second"])
    start, length = oracle.events(x, float(g["threshold"]), RULES_1000)
    assert np.array_equal(start / second, g["event_start_s"])
    assert np.array_equal((start + length) / second, g["event_end_s"])
    assert np.array_equal(length, g["event_n"])
    assert len(oracle.events(x, float(g["threshold"]))[0]) == int(g["n_events_default_rules"]) == 0
    runs = oracle.threshold_runs(x, float(g["threshold"]))
    idx = np.searchsorted(runs[0], start)
    assert np.array_equal(runs[2][idx], g["event_min"]) and np.array_equal(runs[3][idx], g["event_max"])
    for key, kw in SPLIT_SETTINGS.items():
        ev, st, en, _ = oracle.statsplit_events(x, start, length, **kw)
        assert np.array_equal(ev, g[key + "_event"]), key
        assert np.array_equal(st, g[key + "_start"]), key
        assert np.array_equal(en, g[key + "_end"]), key
        assert np.array_equal(st / second, g[key + "_start_s"]), key
        for e in range(len(start)):
            sel = ev == e
            m, s, mn, mx = oracle.segment_stats(x[start[e]:start[e] + length[e]], st[sel], en[sel])
            assert np.array_equal(m, g[key + "_mean"][sel]) and np.array_equal(s, g[key + "_std"][sel])
            assert np.array_equal(mn, g[key + "_min"][sel]) and np.array_equal(mx, g[key + "_max"][sel])


def test_long_event_forced_splits_match_reference_fixture():
    g = load_golden("long_event.npz")
    x32 = synth.make_long_event(300000, seed=100, tier="A")
    assert sha(x32) == str(g["input_sha256"])
    x = x32.astype(np.float64)
    for name, kw in {"long_default": dict(min_width=100, max_width=20000, window_width=10000),
                     "long_psps10": dict(min_width=100, max_width=20000, window_width=10000,
                                         prior_segments_per_second=10),
                     "long_highgain": dict(min_width=100, max_width=15000, window_width=4000,
                                           min_gain_per_sample=2.0)}.items():
        bp = oracle.statsplit(x, **kw)
        edges = np.concatenate(([0], bp, [len(x)]))
        assert np.array_equal(edges[:-1], g[name + "_start"]), name
        assert np.array_equal(edges[1:], g[name + "_end"]), name


@pytest.mark.parametrize("name", ["filter_o1_100k.npz", "filter_o1_250k.npz", "filter_o2_100k.npz",
                                  "filter_o4_100k.npz"])
def test_filter_matches_scipy_fixture(name):
    g = load_golden(name)
    x32 = synth.make_trace(3, seed=int(g["seed"]), tier="A")
    assert sha(x32) == str(g["input_sha256"])
    x = x32.astype(np.float64)
    fs = float(g["fs"])
    b, a = oracle.bessel_ba(int(g["order"]), float(g["cutoff"]), fs)
    assert np.allclose(b, g["b"], rtol=1e-13) and np.allclose(a, g["a"], rtol=1e-13)
    assert np.allclose(oracle.lfilter_zi(b, a), g["zi"], rtol=1e-10)
    start, length = oracle.events(x, 110, RULES_1000)
    for i in range(2):
        assert int(g["event%d_start" % i]) == start[i]
        y = oracle.event_filter(x[start[i]:start[i] + length[i]], fs, int(g["order"]), float(g["cutoff"]))
        ref = g["event%d_filtered" % i]
        assert y.shape == ref.shape
        # scipy's lfilter and this restatement are the same recurrence; allow last-bit noise
        assert np.max(np.abs(y - ref) / np.maximum(np.abs(ref), 1e-12)) < 1e-12
        kw = dict(min_width=100, window_width=10000, sampling_freq=fs, cutoff_freq=float(g["cutoff"]),
                  prior_segments_per_second=10)
        bp = oracle.statsplit(ref, **kw)
        edges = np.concatenate(([0], bp, [len(ref)]))
        assert np.array_equal(edges[:-1], g["event%d_seg_start" % i])
        assert np.array_equal(edges[1:], g["event%d_seg_end" % i])


def test_min_gain_known_answers():
    g = load_golden("params.npz")
    assert oracle.min_gain() == float(g["min_gain_default"]) and str(oracle.min_gain()) == "-0.0"
    assert oracle.min_gain(prior_segments_per_second=10) == float(g["min_gain_psps10"])
    assert oracle.min_gain(prior_segments_per_second=10, cutoff_freq=2000.) == float(g["min_gain_psps10_cut2000"])
    assert oracle.min_gain(min_gain_per_sample=0.5) == float(g["min_gain_per_sample_0p5"])
    assert oracle.min_gain(false_positive_rate=50., sampling_freq=2.5e5, prior_segments_per_second=25.) == \
        float(g["min_gain_fpr50_fs250k"])
    assert "Buffer dtype mismatch" in str(g["float32_error"])
    with pytest.raises(AssertionError):
        oracle.min_gain(min_width=100, max_width=50)
    with pytest.raises(AssertionError):
        oracle.min_gain(min_width=100, window_width=150)
    with pytest.raises(AssertionError):
        oracle.min_gain(cutoff_freq=60000.)


def test_survey_known_answers_c1():
    """config C1 seed 0."""
    for tier, n_default in (("B", 30632), ("A", 30643)):
        x = synth.make_trace(500, seed=0, tier=tier).astype(np.float64)
        assert len(x) == 6003609
        start, length = oracle.events(x, 110, RULES_1000)
        assert len(start) == 500 and int(length.sum()) == 3998546
        assert (int(start[0]), int(start[0] + length[0])) == (3684, 11728)
        assert len(oracle.statsplit_events(x, start, length, threads=8)[1]) == n_default
        ev, st, en, _ = oracle.statsplit_events(x, start, length, threads=8, prior_segments_per_second=10)
        assert len(st) == 2682
        assert list(zip(st[:4], en[:4])) == [(0, 786), (786, 3247), (3247, 5836), (5836, 7192)]
        if tier == "B":
            m, s, _, _ = oracle.segment_stats(x[start[0]:start[0] + length[0]], st[:1], en[:1])
            assert abs(m[0] - 52.3366245968651) < 1e-12 and abs(s[0] - 0.965207331829788) < 1e-12


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built and /root/reference absent")
def test_port_equals_compiled_reference():
    """The C restatement against the reference's own compiled cparsers.pyx, event by event."""
    cp = oracle.load_ref_cparsers()
    x = synth.make_trace(40, seed=11, tier="B").astype(np.float64)
    start, length = oracle.events(x, 110, RULES_1000)
    for kw in (dict(), dict(prior_segments_per_second=10), dict(min_width=30, max_width=900, window_width=400),
               dict(min_gain_per_sample=0.01, window_width=2000)):
        F = cp.FastStatSplit(**kw)
        assert F.min_gain == oracle.min_gain(**kw)
        for s, n in zip(start, length):
            segs = F.parse(x[s:s + n])
            bp = oracle.statsplit(x[s:s + n], **kw)
            assert [g.start for g in segs][1:] == list(bp)
            assert segs[-1].end == n
    # long event, forced max_width splits
    xl = synth.make_long_event(120000, seed=101, tier="B").astype(np.float64)
    kw = dict(min_width=100, max_width=9000, window_width=4000, prior_segments_per_second=10)
    assert [g.start for g in cp.FastStatSplit(**kw).parse(xl)][1:] == list(oracle.statsplit(xl, **kw))


def test_threshold_edge_cases():
    thr = 110.0
    # starts and ends below threshold; single-sample runs; NaN counts as above
    x = np.array([50, 50, 120, 50, 120, 120, np.nan, 50, 50], np.float64)
    start, length, mn, mx, below = oracle.threshold_runs(x, thr)
    assert list(start) == [0, 2, 3, 4, 7] and list(length) == [2, 1, 1, 3, 2]
    assert list(below) == [True, False, True, False, True]
    assert np.isnan(mn[3]) and np.isnan(mx[3]) and mn[0] == 50
    # no crossings at all
    s2 = oracle.threshold_runs(np.full(10, 120.0), thr)
    assert list(s2[0]) == [0] and list(s2[1]) == [10]
    # threshold not float32-representable: compare in double
    t = 110.1
    xf = np.array([np.float32(110.1)], np.float32)  # float32(110.1) > 110.1 as doubles? check both ways
    assert (oracle.threshold_runs(xf.astype(np.float64), t)[4][0]) == (float(xf[0]) < t)


def test_scoring_functions_match_reference_fixture():
    """best_single_split / score_samples (cparsers.pyx:120-155, 205-275) of the restatement against the
    compiled reference's outputs stored by tests/golden/make_golden.py."""
    from conftest import SCORING_CASES, sha
    g = load_golden("scoring.npz")
    for name, (length, seed, tier, kw) in SCORING_CASES.items():
        x = synth.make_long_event(length, seed=seed, tier=tier).astype(np.float64)
        assert sha(x) == str(g[name + "_input_sha256"])
        gain, idx = oracle.best_single_split(x)
        assert gain == float(g[name + "_best_gain"]) and idx == int(g[name + "_best_index"])
        assert np.array_equal(np.array(oracle.score_samples(x, no_split=True, **kw)), g[name + "_no_split"])
        sc = np.array(oracle.score_samples(x, **kw))
        assert sc.shape == g[name + "_scores"].shape and np.array_equal(sc, g[name + "_scores"])
    assert len(g["narrow_forced_scores"]) > 10
    for n in (0, 3, 5, 6, 7):
        x = synth.make_long_event(50, seed=14, tier="A").astype(np.float64)[:n]
        gain, idx = oracle.best_single_split(x)
        assert [gain, idx] == list(g["short%d_best" % n])


def test_oracle_reproduces_the_full_size_long_event_fixture():
    """BASELINE configs[3] at full size (4 events x 10 M samples, max_width=1e6): tests/golden/c4_full.npz was written by
    the real reference; the oracle must hash to the same tables.  Its candidate counts are what the GPU test
    (tests/test_zz_gpu_fullsize.py) expects of the device counters."""
    g = load_golden("c4_full.npz")
    x = synth.make_long_trace(4, 10_000_000, seed0=100, tier="A").astype(np.float64)
    assert sha(x) == str(g["input_sha256"])
    ws, wl = oracle.events(x, 110, RULES_1000)
    assert np.array_equal(ws, g["ev_start"]) and np.array_equal(wl, g["ev_len"])
    for name, kw, cand in (("default", dict(), 300443705), ("psps10", dict(prior_segments_per_second=10), 188617127)):
        oe, ost, oen, nc = oracle.statsplit_events(x, ws, wl, min_width=100, max_width=1000000, window_width=10000,
                                                   threads=8, **kw)
        rows = np.stack([oe, ost, oen], axis=1).astype(np.int64)
        assert len(rows) == int(g[name + "_segments"]) and nc == cand
        assert sha(rows) == str(g[name + "_sha"])


def test_oracle_reproduces_the_full_size_bench_workload_fixture():
    """BASELINE configs[1] at full size = bench.py's workload (make_trace(5000, seed=1), 59,883,057 samples):
    tests/golden/c2_full.npz was written by the real reference; the oracle hashes to the same event and segment
    tables, and its candidate counts are what the device counters must show (tests/test_zz_gpu_fullsize.py)."""
    g = load_golden("c2_full.npz")
    x = synth.make_trace(5000, seed=1, tier="A").astype(np.float64)
    assert len(x) == int(g["samples"]) and sha(x) == str(g["input_sha256"])
    ws, wl = oracle.events(x, 110, RULES_1000)
    assert len(ws) == int(g["events"]) and int(wl.sum()) == int(g["event_samples"])
    assert sha(np.stack([ws, wl], axis=1).astype(np.int64)) == str(g["events_sha"])
    for name, kw, cand in (("default", dict(), 240399270), ("psps10", dict(prior_segments_per_second=10), 135656885)):
        oe, ost, oen, nc = oracle.statsplit_events(x, ws, wl, min_width=100, window_width=10000, threads=8, **kw)
        assert len(oe) == int(g[name + "_segments"]) and nc == cand
        assert sha(np.stack([oe, ost, oen], axis=1).astype(np.int64)) == str(g[name + "_sha"])


def test_full_size_fixtures_equal_the_real_references_tables():
    """The full-size fixtures the oracle made (configs[2], configs[3] at 20 events, the 2 / 4 / 8-GPU traces) were
    re-derived with the REAL reference (tests/golden/make_reference_full_check.py, build container, minutes of CPU):
    its table hashes are committed next to the fixtures and must equal them -- the GPU runs that hash to these fixtures
    therefore hash to the reference's own output."""
    import json
    path = os.path.join(GOLDEN, "reference_full_check.json")
    chk = json.load(open(path))
    fix_b = np.load(os.path.join(GOLDEN, "bench_configs.npz"), allow_pickle=False)
    fix_s = np.load(os.path.join(GOLDEN, "sharded_full.npz"), allow_pickle=False)
    for name, fix, key in (("c3", fix_b, "c3_"), ("c4", fix_b, "c4_"), ("w2", fix_s, "w2_"), ("w4", fix_s, "w4_"),
                           ("w8", fix_s, "w8_")):
        r = chk[name]
        assert r["matches_fixture"] is True, name
        assert r["events_sha"] == str(fix[key + "events_sha"]) and r["segments_sha"] == str(fix[key + "segments_sha"]), name
        assert r["events"] == int(fix[key + "events"]) and r["segments"] == int(fix[key + "segments"]), name
        assert r["samples"] == int(fix[key + "samples"]), name
    fix_5 = np.load(os.path.join(GOLDEN, "c5_files.npz"), allow_pickle=False)
    for name in ("psps10", "default"):          # configs[4]: the reference's Event.filter is scipy's filtfilt
        r = chk["c5_" + name]
        assert r["matches_fixture"] is True and r["segments"] == int(fix_5[name + "_segments"]), name
        assert r["events_sha"] == str(fix_5[name + "_events_sha"]), name
        assert r["segments_sha"] == str(fix_5[name + "_segments_sha"]), name
