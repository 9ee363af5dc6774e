"""Host-side logic that needs no GPU: parameter handling, error parity, the C ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from pypore_b200 import _lib, parsers
from pypore_b200.DataTypes import bessel_coefficients


def test_library_exports_every_declared_symbol():
    """libpypore_b200.so loads without a GPU and exports exactly what include/pypore_b200.h declares."""
    header = open(os.path.join(ROOT, "include", "pypore_b200.h")).read()
    declared = set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", header))
    declared -= {"pp_ctx", "pp_status", "pp_rule", "pp_prefix_mode", "pp_pipeline_params"}
    assert len(declared) >= 25
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), "missing export " + name
    assert declared == set(_lib.SIGNATURES), "ctypes binding and header disagree: %s" % (
        declared ^ set(_lib.SIGNATURES))
    assert L.pp_version() == 100


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.PyPoreCudaError):
        _lib.Context(0)
    with pytest.raises(_lib.PyPoreCudaError):
        parsers.SpeedyStatSplit().parse(np.zeros(1000))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pypore_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f


def test_min_gain_mirror_matches_reference_fixture():
    g = load_golden("params.npz")
    S = parsers.SpeedyStatSplit
    assert S().min_gain == float(g["min_gain_default"]) and str(S().min_gain) == "-0.0"
    assert S(prior_segments_per_second=10).min_gain == float(g["min_gain_psps10"])
    assert S(prior_segments_per_second=10, cutoff_freq=2000.).min_gain == float(g["min_gain_psps10_cut2000"])
    assert S(min_gain_per_sample=0.5).min_gain == float(g["min_gain_per_sample_0p5"])
    assert S(false_positive_rate=50., sampling_freq=2.5e5, prior_segments_per_second=25.).min_gain == \
        float(g["min_gain_fpr50_fs250k"])


def test_statsplit_argument_errors_match_reference():
    with pytest.raises(AssertionError):
        parsers.SpeedyStatSplit(min_width=100, max_width=50).parse(np.zeros(1000))
    with pytest.raises(AssertionError):
        parsers.SpeedyStatSplit(min_width=100, window_width=150).parse(np.zeros(1000))
    with pytest.raises(AssertionError):
        parsers.SpeedyStatSplit(cutoff_freq=60000.).parse(np.zeros(1000))
    # float32 input: the reference raises ValueError (Buffer dtype mismatch), before any device work
    with pytest.raises(ValueError, match="Buffer dtype mismatch, expected 'double' but got 'float'"):
        parsers.SpeedyStatSplit().parse(np.zeros(1000, np.float32))


def test_parser_json_round_trip():
    p = parsers.SpeedyStatSplit(min_width=50, prior_segments_per_second=10, cutoff_freq=2000.)
    d = p.to_dict()
    assert d["name"] == "SpeedyStatSplit" and d["min_width"] == 50 and d["cutoff_freq"] == 2000.
    for key in ("min_width", "max_width", "window_width", "min_gain_per_sample", "false_positive_rate",
                "prior_segments_per_second", "sampling_freq", "cutoff_freq"):
        assert key in d
    q = parsers.parser.from_json(p.to_json())
    assert isinstance(q, parsers.SpeedyStatSplit) and q.__dict__ == p.__dict__
    e = parsers.lambda_event_parser(threshold=110)
    assert e.to_dict() == {"threshold": 110, "name": "lambda_event_parser"}
    e2 = parsers.parser.from_json(e.to_json())
    assert isinstance(e2, parsers.lambda_event_parser) and e2.threshold == 110 and len(e2.rules) == 3


def test_ruleset_is_plain_callables_too():
    rs = parsers.RuleSet(duration_gt=1000, min_gt=-0.5, max_lt=110)

    class E(object):
        duration, min, max = 2000, 10.0, 90.0
    assert len(rs) == 3 and all(r(E()) for r in rs)
    E.max = 111.0
    assert not all(r(E()) for r in rs)
    a = rs.device_args()
    assert a["rule_mask"] == 7 and a["duration_gt"] == 1000 and a["min_gt"] == -0.5 and a["max_lt"] == 110.0
    assert parsers.RuleSet(duration_gt=10.5).device_args()["duration_gt"] == 10
    assert parsers.RuleSet(duration_lt=10.5).device_args()["duration_lt"] == 11


@pytest.mark.parametrize("name", ["filter_o1_100k.npz", "filter_o1_250k.npz", "filter_o2_100k.npz",
                                  "filter_o4_100k.npz"])
def test_bessel_coefficients_match_scipy_fixture(name):
    g = load_golden(name)
    b, a, zi = bessel_coefficients(int(g["order"]), float(g["cutoff"]), float(g["fs"]))
    assert np.allclose(b, g["b"], rtol=1e-13) and np.allclose(a, g["a"], rtol=1e-13)
    assert np.allclose(zi, g["zi"], rtol=1e-10)


def test_device_trace_dtype_choice():
    """float64 traces go up as float32 only when that is lossless; what float32 cannot hold (the reference's own
    int16 x float64-scale data, read_abf.py:208-210) stays float64; integers are exact."""
    x = np.array([1.0, 2.5, 120.03125])
    assert parsers._device_trace(x).dtype == np.float32
    assert parsers._device_trace(x.astype(np.float32)).dtype == np.float32
    y = parsers._device_trace(np.array([0.1, 120.0]))
    assert y.dtype == np.float64 and y[0] == 0.1
    counts = np.array([1, -2, 3000], np.int16)
    assert parsers._device_trace(counts).dtype == np.float32                       # small integers fit float32
    assert parsers._device_trace(counts * 0.030517578125).dtype == np.float32      # power-of-two scale: still exact
    assert parsers._device_trace(counts * 0.0305).dtype == np.float64              # a typical ADC scale is not
    assert parsers._device_trace(np.array([2**40 + 1], np.int64)).dtype == np.float64
    with pytest.raises(TypeError):
        parsers._device_trace(np.array(["a"]))


def test_file_json_round_trip_of_the_reference_format():
    """the JSON written by the REAL reference's File.to_json (tests/golden/file_tierA.json)
    is read by the mirror's File.from_json (MetaEvents / MetaSegments, the reference's own fallback without the
    .abf file) and written back identically, parsers included."""
    import json
    import os
    from conftest import GOLDEN
    from pypore_b200.DataTypes import File, MetaEvent
    from pypore_b200.core import MetaSegment
    from pypore_b200.parsers import SpeedyStatSplit, lambda_event_parser
    text = open(os.path.join(GOLDEN, "file_tierA.json")).read()
    ref = json.loads(text)
    f = File.from_json(text)
    assert f.n == ref["n"] == 4 and isinstance(f.event_parser, lambda_event_parser) and f.event_parser.threshold == 110
    assert all(isinstance(e, MetaEvent) for e in f.events)
    assert all(isinstance(s, MetaSegment) for e in f.events for s in e.segments)
    assert isinstance(f.events[0].state_parser, SpeedyStatSplit) and f.events[0].state_parser.cutoff_freq == 2000.0
    assert [e.filtered for e in f.events] == [True, False, True, False]
    back = json.loads(f.to_json())
    for key in ("filename", "n", "event_parser", "name"):
        assert back[key] == ref[key]
    assert len(back["events"]) == len(ref["events"])
    for a, b in zip(back["events"], ref["events"]):
        assert set(a) == set(b) - {"filtered"}      # MetaEvent.to_dict has no 'filtered' key (DataTypes.py:196-201)
        for k in a:
            if k == "name":
                assert a[k] == "MetaEvent" and b[k] == "Event"      # what the reference's reload yields too
            elif k == "segments":
                assert [{**s, "name": "Segment"} for s in a[k]] == b[k]
            else:
                assert a[k] == b[k], k
    # Event.to_json / from_json and the file-name form of from_json
    ev = f.events[1]
    again = MetaEvent.from_json(ev.to_json())
    assert again.start == ev.start and again.n == ev.n and len(again.segments) == len(ev.segments)


def _fake_tables(meta, with_segments=True, filtered=True):
    """Three events (one without segments) as the device would hand them back, in samples."""
    from pypore_b200 import wire
    rng = np.random.RandomState(5)
    ev = dict(start=np.array([100, 5000, 9000], np.int64), length=np.array([3000, 2000, 500], np.int64))
    for k in ("mean", "std", "min", "max"):
        ev[k] = rng.rand(3)
    sg = None
    if with_segments:
        sg = dict(event=np.array([0, 0, 0, 1, 1], np.int32), start=np.array([0, 700, 1800, 0, 900], np.int64),
                  end=np.array([700, 1800, 3000, 900, 2000], np.int64))
        for k in ("mean", "std", "min", "max"):
            sg[k] = rng.rand(5)
    seg = parsers.SpeedyStatSplit(prior_segments_per_second=10, cutoff_freq=2000.) if with_segments else None
    return wire.FileTables(1e5, ev, sg, (1, 2000.) if filtered else None, seg, meta=meta)


@pytest.mark.parametrize("meta", [False, True])
@pytest.mark.parametrize("with_segments", [False, True])
@pytest.mark.parametrize("filtered", [False, True])
def test_table_first_json_equals_the_object_walk(meta, with_segments, filtered):
    """File.to_json straight from the tables (no Event / Segment object built) gives the same tree as walking the
    objects the lazy view builds from the same tables -- for live files, metadata files, with and without segments."""
    import json
    from pypore_b200 import wire
    from pypore_b200.DataTypes import File, MetaEvent, _LazyList
    x = np.zeros(10000)
    f = File(current=x, timestep=0.01)
    f.event_parser = parsers.lambda_event_parser(threshold=110)
    f._attach(_fake_tables(meta, with_segments, filtered), host=x,
              filtered=np.zeros(5500) if filtered and not meta else None)
    del f.current   # (the file-level mean / std would otherwise ask the device)
    assert f._tables_current() and isinstance(f.events, _LazyList)
    counts = f._segment_counts()
    from_tables = json.loads(f.to_json())
    assert f._tables_current()                      # still no object handed out
    events = list(f.events)
    for e in events:
        list(e.segments)
    assert not f._tables_current()
    walked = json.loads(f.to_json())
    assert from_tables == walked
    assert counts == [e.n for e in events] == ([3, 2, 0] if with_segments else [0, 0, 0])
    assert from_tables["events"][0]["name"] == ("MetaEvent" if meta else "Event")
    assert ("segments" in from_tables["events"][0]) == with_segments
    assert all(isinstance(e, MetaEvent) == meta for e in events)
    if not meta:
        # to_meta on an untouched table-backed file flips the tables, on a touched one it converts the objects
        g = File(current=x, timestep=0.01)
        g.event_parser = f.event_parser
        g._attach(_fake_tables(False, with_segments, filtered), host=x, filtered=np.zeros(5500) if filtered else None)
        g.to_meta()
        f.to_meta()
        a, b = json.loads(g.to_json()), json.loads(f.to_json())
        assert a == b and not hasattr(g, "current")
        assert all(isinstance(e, MetaEvent) for e in list(g.events) + list(f.events))


def test_rules_assigned_after_construction_take_effect():
    """`parser.rules = [...]` after construction (supported by the reference, whose lambdas are looked up
    at call time) must not be shadowed by the device form of the default rules."""
    p = parsers.lambda_event_parser(threshold=110)
    assert p._device_rules().device_args()["duration_gt"] == 100000
    p.threshold = 95
    assert p._device_rules().max_lt == 95            # the default rule reads the threshold at call time
    p.rules = [lambda e: e.duration > 10]
    assert p._device_rules() is None                 # arbitrary callables: evaluated on the host
    p.rules = parsers.RuleSet(duration_gt=10)
    assert p._device_rules().device_args()["duration_gt"] == 10
    q = parsers.parser.from_json(p.to_json())        # the private list of defaults stays out of the JSON
    assert q.threshold == 95 and "_builtin_rules" not in p.to_dict()


def test_infinite_duration_rules_clamp():
    a = parsers.RuleSet(duration_gt=-np.inf, duration_lt=np.inf).device_args()
    assert a["duration_gt"] < -9e18 and a["duration_lt"] > 9e18
    assert -2**63 <= a["duration_gt"] and a["duration_lt"] < 2**63
    assert parsers.RuleSet(duration_gt=10.5, duration_lt=10.5).device_args()["duration_gt"] == 10
    with pytest.raises(ValueError):
        parsers.RuleSet(duration_gt=float("nan")).device_args()


def test_ignored_and_segment_semantics_without_a_device():
    from pypore_b200.core import MetaSegment, Segment, ignored
    with ignored(KeyError, AttributeError):
        {}["x"]
    with pytest.raises(ValueError):
        with ignored(KeyError):
            raise ValueError("not swallowed")
    s = Segment(np.arange(5.), start=10, duration=5, mean=99.0)     # statistics cannot be overridden
    assert "mean" not in s.__dict__ and s.n == 5 and len(s) == 5
    s._set_stats(2.0, 1.4, 0.0, 4.0)
    assert s.to_dict() == {"mean": 2.0, "std": 1.4, "min": 0.0, "max": 4.0, "start": 10, "duration": 5,
                           "name": "Segment"}
    s.scale(10.)                                                    # no `end`: stops after `start` (core.py:199-207)
    assert s.start == 1.0 and s.duration == 5
    s.to_meta()
    assert type(s) is MetaSegment and not hasattr(s, "current") and s.mean == 2.0
    m = MetaSegment(start=2, end=5)
    assert m.duration == 3 and MetaSegment(end=5, duration=3).start == 2 and MetaSegment(start=2, duration=3).end == 5
    back = MetaSegment.from_json(json=m.to_json())
    assert back.start == "2" and back.name == "MetaSegment"         # the flat reader keeps strings, like the reference


def test_fast_pretty_printer_is_json_dumps_byte_for_byte():
    """wire.pretty writes long lists of flat dicts from a template; the text must be exactly what the reference's
    json.dumps(indent=4, separators=(',', ' : ')) layout gives, including NaN / Infinity, escapes, empty containers,
    lists that are not homogeneous (those fall back) and '%' in keys."""
    import json
    from pypore_b200 import wire
    trees = [
        {"a": [{"x": 1.5, "y": float("nan"), "n": "Segment"}, {"x": -float("inf"), "y": 2, "n": "Segé\"q"}],
         "b": [], "c": [{}], "d": [{"k": [1]}], "e": [[{"z": 1}]], "f": {"g": [{"h": True, "i": None, "j": False}]}},
        [{"a": 1}, {"a": 2}], {"a": [{"x": 1}, {"y": 2}]}, {"a": [{"x": 1}, 3]}, {"a": [{"%s": 1.0, "100%": 2}]},
        {"a": [{"x": 1.0}, {"x": 2}]}, {"a": [{"x": True}, {"x": 2}]}, {"a": [{"x": np.float64(1.5)}, {"x": 2.0}]},
        {"events": [{"mean": 1.0, "segments": [{"mean": 0.1, "name": "Segment"}] * 3,
                     "state_parser": {"name": "p", "w": 1}}] * 2},
        {"a": [{"x": float(i) / 7, "s": "v%d" % (i % 3), "i": i} for i in range(200)]}, 5, "text", [], {},
    ]
    for tree in trees:
        assert wire.pretty(tree) == json.dumps(tree, **wire.LAYOUT)
