"""The drop-in surface on the GPU: these read like scripts written against PyPore
(File.parse / Event.filter / Event.parse, Segment attributes) and are checked against the
fixtures the real reference produced (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import oracle
from conftest import RULES_1000, load_golden
from pypore_b200 import synth
from pypore_b200.DataTypes import Event, File
from pypore_b200.core import MetaSegment, Segment
from pypore_b200.parsers import MemoryParse, RuleSet, SpeedyStatSplit, lambda_event_parser

pytestmark = pytest.mark.gpu


def fixture_file(name="pipeline_tierA.npz"):
    g = load_golden(name)
    # float64 like the reference's traces (read_abf.py:210); FastStatSplit rejects float32 events
    x = synth.make_trace(int(g["n_events"]), seed=int(g["seed"]), tier=str(g["tier"])).astype(np.float64)
    return g, x, File(current=x, timestep=1000. / float(g["fs"]))


def test_reference_style_script_matches_reference_output():
    g, x, f = fixture_file()
    assert f.second == float(g["second"])
    f.parse(parser=lambda_event_parser(threshold=110, rules=RULES_1000))   # arbitrary Python rules (host path)
    assert len(f.events) == len(g["event_n"]) and f.n == len(f.events)
    assert np.array_equal([e.start for e in f.events], g["event_start_s"])
    assert np.array_equal([e.end for e in f.events], g["event_end_s"])
    assert np.array_equal([e.duration for e in f.events], g["event_duration_s"])
    p = SpeedyStatSplit(min_width=100, window_width=10000, prior_segments_per_second=10)
    st, en, mean, std = [], [], [], []
    for event in f.events:
        assert isinstance(event, Event) and event.file is f and event.filtered is False
        event.parse(parser=p)
        assert event.state_parser is p and event.n == len(event.segments)
        for seg in event.segments:
            assert seg.event is event
            st.append(seg.start); en.append(seg.end); mean.append(seg.mean); std.append(seg.std)
            assert seg.duration == seg.end - seg.start or abs(seg.duration - (seg.end - seg.start)) < 1e-12
    assert np.array_equal(st, g["psps10_start_s"]) and np.array_equal(en, g["psps10_end_s"])
    assert np.allclose(mean, g["psps10_mean"], rtol=1e-9, atol=0) and np.allclose(std, g["psps10_std"], rtol=1e-9, atol=0)
    # event statistics are properties of the Event itself (it is a Segment)
    e0 = f.events[0]
    assert abs(e0.mean - g["event_mean"][0]) <= 1e-9 * abs(g["event_mean"][0])
    assert e0.min == g["event_min"][0] and e0.max == g["event_max"][0]


def test_default_rules_reject_everything_like_the_reference():
    g, x, f = fixture_file()
    f.parse(parser=lambda_event_parser(threshold=110))
    assert len(f.events) == int(g["n_events_default_rules"]) == 0


def test_raw_parser_protocol_objects():
    g, x, f = fixture_file("pipeline_tierB.npz")
    pieces = lambda_event_parser(threshold=110, rules=RuleSet(duration_gt=1000, min_gt=-0.5, max_lt=110)).parse(x)
    assert [int(p.duration) for p in pieces] == list(g["event_n"])
    assert all(isinstance(p.start, np.int64) and not hasattr(p, "end") for p in pieces)
    assert pieces[0].current.base is None or pieces[0].current.base is not x   # a copy, like parsers.py:152
    cur = pieces[0].current.astype(np.float64)
    segs = SpeedyStatSplit(min_width=100, window_width=10000).parse(cur)
    sel = g["default_event"] == 0
    assert [s.start for s in segs] == list(g["default_start"][sel]) and [s.end for s in segs] == list(g["default_end"][sel])
    assert all(s.duration == s.end - s.start and s.n == s.duration for s in segs)
    assert segs[3].current.base is cur                                     # views, like cparsers.pyx:115
    assert np.allclose([s.mean for s in segs], g["default_mean"][sel], rtol=1e-9, atol=0)
    assert np.allclose([s.min for s in segs], g["default_min"][sel], rtol=0, atol=0)
    d = segs[0].to_dict()
    assert d["name"] == "Segment" and d["start"] == 0 and set(d) >= {"mean", "std", "min", "max", "end", "duration"}
    segs[0].to_meta()
    assert isinstance(segs[0], MetaSegment) and not hasattr(segs[0], "current") and segs[0].mean == d["mean"]
    # MemoryParse replays stored boundaries (parsers.py:110-122)
    again = MemoryParse(g["default_start"][sel], g["default_end"][sel]).parse(cur)
    assert np.allclose([s.mean for s in again], g["default_mean"][sel], rtol=1e-9, atol=0)
    # an empty event gives the reference's single empty segment
    empty = SpeedyStatSplit().parse(np.zeros(0))
    assert len(empty) == 1 and empty[0].start == 0 and empty[0].end == 0


def test_user_made_segment_statistics_come_from_the_gpu():
    rng = np.random.RandomState(1)
    a = rng.normal(50, 3, 12345)
    s = Segment(current=a, start=10, duration=12345)
    assert abs(s.mean - np.mean(a)) < 1e-9 * abs(np.mean(a)) and abs(s.std - np.std(a)) < 1e-9 * np.std(a)
    assert s.min == a.min() and s.max == a.max() and s.n == 12345
    s.current = a[:100]                     # new samples invalidate the cached statistics
    assert abs(s.mean - np.mean(a[:100])) < 1e-9 * abs(np.mean(a[:100]))
    m = MetaSegment(current=a, start=0, duration=5)
    assert m.n == 12345 and m.end == 5 and abs(m.std - np.std(a)) < 1e-9 * np.std(a) and not hasattr(m, "current")


def test_device_resident_pipeline_equals_event_by_event_calls():
    g, x, f = fixture_file()
    seg = SpeedyStatSplit(min_width=50, max_width=2500, window_width=1000, prior_segments_per_second=50)
    f.parse(parser=lambda_event_parser(threshold=110, rules=RuleSet(duration_gt=1000, min_gt=-0.5, max_lt=110)),
            segmenter=seg)
    t = f.segment_table
    assert np.array_equal(t["event"], g["narrow_event"]) and np.array_equal(t["start"], g["narrow_start"])
    assert np.array_equal(t["end"], g["narrow_end"])
    st = [s.start for e in f.events for s in e.segments]
    en = [s.end for e in f.events for s in e.segments]
    assert np.array_equal(st, g["narrow_start_s"]) and np.array_equal(en, g["narrow_end_s"])
    ev = f.events[2]
    assert ev.state_parser is seg and ev.segments[1].event is ev and ev.n == int((g["narrow_event"] == 2).sum())
    k = int(np.searchsorted(g["narrow_event"], 2)) + 1
    assert abs(ev.segments[1].mean - g["narrow_mean"][k]) <= 1e-9 * abs(g["narrow_mean"][k])
    # host-rule variant of the same pipeline (rules evaluated in Python on the device run table)
    f2 = File(current=x, timestep=1000. / float(g["fs"]))
    f2.parse(parser=lambda_event_parser(threshold=110, rules=RULES_1000), segmenter=seg)
    assert all(np.array_equal(f2.segment_table[k], t[k]) for k in ("event", "start", "end", "mean", "std"))


def test_error_behaviour_on_device_paths():
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        SpeedyStatSplit().parse(np.zeros(500, np.float32))
    with pytest.raises(AssertionError):
        SpeedyStatSplit(min_width=100, window_width=150).parse(np.zeros(500))
    f = File(current=np.zeros(10, np.float32), timestep=0.01)
    with pytest.raises(SyntaxError):
        File()
    ev = Event(current=np.zeros(50), start=0, end=1, duration=1, second=1e5, file=f)
    ev.parse(parser=SpeedyStatSplit())                     # shorter than 2*min_width: one segment
    assert ev.n == 1 and ev.segments[0].end == 50 / 1e5
    with pytest.raises(AttributeError):                    # no .file, like DataTypes.py:289
        Event(current=np.zeros(300), start=0, second=1e5).parse(parser=SpeedyStatSplit())


def test_event_filter_then_parse_like_the_reference():
    g = load_golden("filter_o1_250k.npz")
    x = synth.make_trace(3, seed=int(g["seed"]), tier="A").astype(np.float64)
    f = File(current=x, timestep=1000. / float(g["fs"]))
    f.parse(parser=lambda_event_parser(threshold=110, rules=RULES_1000))
    p = SpeedyStatSplit(min_width=100, window_width=10000, sampling_freq=float(g["fs"]), cutoff_freq=2000.,
                        prior_segments_per_second=10)
    for i, event in enumerate(f.events[:2]):
        event.filter(order=1, cutoff=2000.)
        assert event.filtered and event.filter_order == 1 and event.filter_cutoff == 2000.
        assert event.current.dtype == np.float64
        ref = g["event%d_filtered" % i]
        assert np.max(np.abs(event.current - ref) / np.abs(ref)) < 1e-5
        event.parse(parser=p)
        assert [round(s.start * f.second) for s in event.segments] == list(g["event%d_seg_start" % i])
    with pytest.raises(ValueError, match="padlen"):
        Event(current=np.zeros(6), start=0, second=1e5, file=f).filter()
    # whole-file device-resident variant
    f2 = File(current=x, timestep=1000. / float(g["fs"]))
    f2.parse(parser=lambda_event_parser(threshold=110, rules=RuleSet(duration_gt=1000, min_gt=-0.5, max_lt=110)),
             segmenter=p, filter_params=(1, 2000.))
    ev = f2.events[1]
    assert ev.filtered and np.max(np.abs(ev.current - g["event1_filtered"]) / np.abs(g["event1_filtered"])) < 1e-5
    sel = f2.segment_table["event"] == 1
    assert np.array_equal(f2.segment_table["start"][sel], g["event1_seg_start"])


def test_best_single_split_and_score_samples_match_reference():
    """FastStatSplit.best_single_split / score_samples on the device, bit-equal to the
    compiled reference's fixture and to the oracle (exact arithmetic, recursion order, dense score arrays)."""
    from pypore_b200.cparsers import FastStatSplit
    from conftest import SCORING_CASES
    g = load_golden("scoring.npz")
    for name, (length, seed, tier, kw) in SCORING_CASES.items():
        x = synth.make_long_event(length, seed=seed, tier=tier).astype(np.float64)
        gain, idx = FastStatSplit(**kw).best_single_split(x)
        assert idx == int(g[name + "_best_index"]) and abs(gain - float(g[name + "_best_gain"])) <= 1e-9 * abs(gain)
        ns = FastStatSplit(**kw).score_samples(x, no_split=True)
        assert isinstance(ns, list) and len(ns) == length
        ref = g[name + "_no_split"]
        assert np.allclose(np.array(ns), ref, rtol=1e-9, atol=1e-9) and np.array_equal(np.array(ns) == 0, ref == 0)
        sc = np.array(FastStatSplit(**kw).score_samples(x))
        ref = g[name + "_scores"]
        assert sc.shape == ref.shape                       # same scans, same recursion order
        assert np.array_equal(sc == 0, ref == 0) and np.allclose(sc, ref, rtol=1e-9, atol=1e-9)
    for n in (0, 3, 5, 6, 7):
        x = synth.make_long_event(50, seed=14, tier="A").astype(np.float64)[:n]
        gain, idx = SpeedyStatSplit().best_single_split(x)
        assert idx == int(g["short%d_best" % n][1])
        if idx >= 0:
            assert abs(gain - g["short%d_best" % n][0]) <= 1e-9 * abs(gain)
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        SpeedyStatSplit().best_single_split(np.zeros(100, np.float32))
    with pytest.raises(AssertionError):
        FastStatSplit(min_width=100, max_width=50)


def _close(a, b, rtol):
    if isinstance(b, float):
        return a == b or abs(a - b) <= rtol * max(abs(a), abs(b))
    return a == b


def test_file_to_json_matches_the_reference_format():
    """File.parse -> Event.filter -> Event.parse -> File.to_json on the device against the
    JSON the real reference wrote for the same trace (structure and keys equal, indices exact, statistics
    1e-9, statistics of filtered events 1e-5 like the filtered current)."""
    import json
    import os
    from conftest import GOLDEN
    ref = json.loads(open(os.path.join(GOLDEN, "file_tierA.json")).read())
    x64 = synth.make_trace(4, seed=21, tier="A").astype(np.float64)
    f = File(current=x64, timestep=0.01)
    f.parse(parser=lambda_event_parser(threshold=110, rules=RULES_1000))
    seg = SpeedyStatSplit(min_width=100, window_width=10000, prior_segments_per_second=10, cutoff_freq=2000.)
    for i, event in enumerate(f.events):
        if i % 2 == 0:
            event.filter(1, 2000.)
        event.parse(parser=seg)
    ours = json.loads(f.to_json())
    assert set(ours) == set(ref) and ours["n"] == ref["n"] and ours["event_parser"] == ref["event_parser"]
    assert _close(ours["mean"], ref["mean"], 1e-9) and _close(ours["std"], ref["std"], 1e-9)
    for a, b in zip(ours["events"], ref["events"]):
        assert set(a) == set(b) and a["state_parser"] == b["state_parser"] and a["n"] == b["n"]
        rtol = 1e-5 if b["filtered"] else 1e-9
        for k in ("start", "end", "duration", "filtered", "filter_order", "filter_cutoff", "name"):
            assert a.get(k) == b.get(k), k
        for k in ("mean", "std", "min", "max"):
            assert _close(a[k], b[k], rtol), (k, a[k], b[k])
        assert len(a["segments"]) == len(b["segments"])
        for sa, sb in zip(a["segments"], b["segments"]):
            assert set(sa) == set(sb) and sa["name"] == sb["name"]
            assert (sa["start"], sa["end"], sa["duration"]) == (sb["start"], sb["end"], sb["duration"])
            assert all(_close(sa[k], sb[k], rtol) for k in ("mean", "std", "min", "max"))
    # reload: metadata objects, like the reference without the .abf file
    g = File.from_json(f.to_json())
    assert g.n == 4 and [e.n for e in g.events] == [e.n for e in f.events]


def test_experiment_parse_batch_driver(capsys):
    """Experiment.parse (DataTypes.py:956-988) over File objects -- one device-resident pass
    per file -- gives the events / segments of the reference's per-event loop; meta=True leaves metadata only."""
    from pypore_b200.DataTypes import Experiment, MetaEvent
    traces = [synth.make_trace(3, seed=31 + k, tier="A") for k in range(2)]
    det = lambda_event_parser(threshold=110, rules=RuleSet(duration_gt=1000, min_gt=-0.5, max_lt=110))
    seg = SpeedyStatSplit(prior_segments_per_second=10, cutoff_freq=2000.)
    exp = Experiment([File(current=t, timestep=0.01) for t in traces], name="batch")
    exp.parse(event_detector=det, segmenter=seg, filter_params=(1, 2000), verbose=True, meta=False)
    out = capsys.readouterr().out
    assert out.count("Opening") == 2 and "\tDetected 3 Events" in out and "\t\tEvent 1 has" in out
    assert exp.n == 2 and len(exp.events) == 6 and len(exp.segments) == sum(e.n for e in exp.events)
    # the same through the reference's own loop shape: File.parse, Event.filter, Event.parse
    for t, file in zip(traces, exp.files):
        f = File(current=t, timestep=0.01)
        f.parse(parser=det)
        for a, b in zip(f.events, file.events):
            a.filter(1, 2000)
            a.parse(parser=seg)
            assert b.filtered and (a.start, a.end) == (b.start, b.end)
            assert [(s.start, s.end) for s in a.segments] == [(s.start, s.end) for s in b.segments]
            assert np.allclose([s.mean for s in a.segments], [s.mean for s in b.segments], rtol=1e-9, atol=0)
    exp2 = Experiment([File(current=t, timestep=0.01) for t in traces])
    exp2.parse(event_detector=det, segmenter=seg, filter_params=None, verbose=False, meta=True)
    assert capsys.readouterr().out == ""
    assert all(isinstance(e, MetaEvent) and not hasattr(e, "current") for e in exp2.events)
    assert not hasattr(exp2.files[0], "current") and exp2.files[0].events[0].segments[0].mean > 0
    with pytest.raises(NotImplementedError):
        Experiment(["run1.abf"]).parse(verbose=False)


def test_filter_derivative_segmenter_matches_reference():
    """FilterDerivativeSegmenter (parsers.py:609-656) against what the real reference returned
    (tests/golden/make_golden.py: golden_fds), its quirks included; the filtfilt runs on the device."""
    from make_golden_cases import FDS_CASES
    from pypore_b200.parsers import FilterDerivativeSegmenter
    g = load_golden("fds.npz")
    for name, (length, seed, tier, kw) in FDS_CASES.items():
        x = synth.make_long_event(length, seed=seed, tier=tier).astype(np.float64)
        p = FilterDerivativeSegmenter(**kw)
        assert (p.low_threshold, p.high_threshold, p.cutoff_freq, p.sampling_freq) == \
            (kw["low_threshold"], kw["high_threshold"], kw["cutoff_freq"], kw["sampling_freq"])
        segs = p.parse(x)
        assert np.array_equal([s.start for s in segs], g[name + "_start"]), name
        assert np.array_equal([len(s.current) for s in segs], g[name + "_n"]), name
        assert np.allclose([s.mean for s in segs], g[name + "_mean"], rtol=1e-9, atol=0), name
        assert all(s.current.base is x or s.current is x for s in segs)      # views of the caller's array
    assert p.to_dict()["name"] == "FilterDerivativeSegmenter"
    with pytest.raises(ValueError):
        FilterDerivativeSegmenter().parse(np.zeros(5))                         # not longer than filtfilt's padlen


def test_float64_file_and_empty_trace_and_reassigned_rules():
    """The public API on the reference's native input: a float64 trace float32 cannot hold (int16 x scale,
    read_abf.py:208-210) through File.parse / lambda_event_parser.parse / Event.parse; an empty trace gives no
    events instead of an error; rules assigned after construction are the ones applied."""
    x32 = synth.make_trace(6, seed=13, tier="A")
    x = np.round(x32.astype(np.float64) / 0.0305).astype(np.int16) * 0.0305
    assert not np.array_equal(x.astype(np.float32).astype(np.float64), x)
    ws, wl = oracle.events(x, 110, [lambda e: e.duration > 1000, lambda e: e.min > -0.5, lambda e: e.max < 110])
    seg = SpeedyStatSplit(min_width=100, window_width=10000, prior_segments_per_second=10)
    f = File(current=x, timestep=0.01)
    f.parse(parser=lambda_event_parser(threshold=110, rules=RuleSet(duration_gt=1000, min_gt=-0.5, max_lt=110)),
            segmenter=seg)
    assert np.array_equal(f.event_table["start"], ws) and np.array_equal(f.event_table["length"], wl)
    oe, ost, oen, _ = oracle.statsplit_events(x, ws, wl, prior_segments_per_second=10)
    t = f.segment_table
    assert np.array_equal(t["event"], oe) and np.array_equal(t["start"], ost) and np.array_equal(t["end"], oen)
    # the reference's own call sequence, event by event, on the same doubles
    g = File(current=x, timestep=0.01)
    g.parse(parser=lambda_event_parser(threshold=110, rules=RULES_1000))
    assert [int(round(e.start * g.second)) for e in g.events] == list(ws)
    n0 = g.events[0].parse(parser=seg)
    assert n0 == g.events[0].n == int((oe == 0).sum())           # Event.parse returns the segment count
    assert np.array_equal(g.events[0].current, x[ws[0]:ws[0] + wl[0]])
    # integer counts are accepted like the reference accepts them
    h = File(current=np.round(x32.astype(np.float64)).astype(np.int16), timestep=0.01)
    h.parse(parser=lambda_event_parser(threshold=110, rules=RULES_1000))
    assert h.n == len(oracle.events(np.round(x32.astype(np.float64)), 110, RULES_1000)[0])
    # empty trace
    e = File(current=np.zeros(0, np.float32), timestep=0.01)
    e.parse(parser=lambda_event_parser(threshold=110, rules=RuleSet(duration_gt=1000)), segmenter=seg)
    assert e.n == 0 and len(e.segment_table["start"]) == 0 and json_ok(e)
    assert lambda_event_parser(threshold=110).parse(np.zeros(0)) == []
    # rules reassigned after construction
    p = lambda_event_parser(threshold=110)
    assert len(p.parse(x)) == 0                                   # the defaults want duration > 100000
    p.rules = [lambda ev: ev.duration > 1000, lambda ev: ev.max < 110]
    assert len(p.parse(x)) == len(oracle.events(x, 110, p.rules)[0]) > 0
    k = File(current=x, timestep=0.01)
    k.parse(parser=p, segmenter=seg)                              # host-evaluated rules through the resident path
    assert k.n == len(p.parse(x))


def json_ok(f):
    import json
    d = json.loads(f.to_json())
    return d["n"] == 0 and d["events"] == []
