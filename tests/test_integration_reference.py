"""INTEGRATION.md section 1 for real: this repository's parsers handed to the REFERENCE's own containers.

The reference's plug-in boundary is duck-typed (`parser.parse(current) -> [Segment-like]`, PyPore/parsers.py:57-59):
`File.parse` reads `seg.current / seg.start / seg.duration` of what comes back (DataTypes.py:595-600), `Event.parse`
sets `segment.event` and calls `segment.scale(...)` (DataTypes.py:286-289), `File.to_json` asks the parsers for
`to_dict()` (DataTypes.py:729).  Here the reference's `File` / `Event` (loaded from /root/reference like the golden
fixtures were, tests/golden/make_golden.py) drive `pypore_b200.parsers.lambda_event_parser` and `SpeedyStatSplit`, and
the result is compared with the reference driving its OWN parsers on the same trace.

The reference only exists in the build container, the GPU only on the GPU box: the CPU variant runs the plug-ins over
an oracle-backed stand-in for the device context (tests/oracle_device.py) and checks the protocol end to end; the
`gpu` variant is the same test on the real device and runs wherever both are present."""
import json
import os
import sys

import numpy as np
import pytest

from pypore_b200 import synth

REFERENCE = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "PyPore")),
                                reason="the reference tree is only present in the build container")


@pytest.fixture(scope="module")
def reference():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    return make_golden.load_reference()


def run_both(reference):
    """(what the reference's File / Event produce with our plug-ins, what they produce with their own)."""
    from pypore_b200 import parsers as ours
    dt, ref_parsers, _ = reference
    x = synth.make_trace(4, seed=21, tier="A").astype(np.float64)
    rules = [lambda e: e.duration > 1000, lambda e: e.min > -0.5, lambda e: e.max < 110]
    kw = dict(min_width=100, window_width=10000, prior_segments_per_second=10, cutoff_freq=2000.)
    out = []
    for P in (ours, ref_parsers):
        f = dt.File(current=x, timestep=0.01)
        f.parse(parser=P.lambda_event_parser(threshold=110, rules=rules))
        for i, event in enumerate(f.events):
            if i % 2 == 0:
                event.filter(1, 2000.)       # the reference's own scipy filter: the parsers see float64 either way
            event.parse(parser=P.SpeedyStatSplit(**kw))
        out.append((f, json.loads(f.to_json())))
    return out


def compare(out, stat_rtol):
    (f_ours, j_ours), (f_ref, j_ref) = out
    assert j_ours["n"] == j_ref["n"] == 4 and j_ours["event_parser"] == j_ref["event_parser"]
    assert type(f_ours.events[0]).__module__ == type(f_ref.events[0]).__module__     # the reference's Event class
    for a, b in zip(j_ours["events"], j_ref["events"]):
        assert set(a) == set(b) and a["state_parser"] == b["state_parser"] and a["n"] == b["n"] > 1
        for k in ("start", "end", "duration", "filtered", "name"):
            assert a[k] == b[k], k
        for sa, sb in zip(a["segments"], b["segments"]):
            assert (sa["start"], sa["end"], sa["duration"], sa["name"]) == (sb["start"], sb["end"], sb["duration"],
                                                                           sb["name"])
            for k in ("mean", "std", "min", "max"):
                assert abs(sa[k] - sb[k]) <= stat_rtol * abs(sb[k]), k
    # our Segment objects inside the reference's Event: scaled to seconds, linked to the event, sample views intact
    ev = f_ours.events[1]
    seg = ev.segments[1]
    assert seg.event is ev and abs(seg.duration * f_ours.second - len(seg.current)) < 1e-6
    assert np.array_equal(seg.current, ev.current[int(round(seg.start * f_ours.second)):][:len(seg.current)])


def test_reference_containers_drive_our_parsers_oracle_device(reference, monkeypatch):
    from oracle_device import OracleDevice
    from pypore_b200 import _lib
    dev = OracleDevice()
    monkeypatch.setattr(_lib, "default_context", lambda device=None: dev)
    compare(run_both(reference), stat_rtol=1e-12)


@pytest.mark.gpu
def test_reference_containers_drive_our_parsers_on_the_device(reference):
    compare(run_both(reference), stat_rtol=1e-9)
