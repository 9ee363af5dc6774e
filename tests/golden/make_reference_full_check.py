"""tests/golden/reference_full_check.json: the REAL reference on every full-size bench workload whose fixture was
made by the CPU oracle -- BASELINE configs[2] (one 360 M-sample trace), configs[3] (20 events of 10 M samples,
max_width = 1e6), the multi-GPU workloads of bench.py --gpus 2 / 4 / 8 (120 M / 240 M / 480 M samples) and the eight
distinct files of configs[4] (Event.filter = scipy's filtfilt, then the split on the filtered current).

    python tests/golden/make_reference_full_check.py        # build container only; ~3 min on 8 cores, ~25 GB of RAM

The reference's own code does the work (loaded like tests/golden/make_golden.py does: File.parse with its
lambda_event_parser on the whole float64 trace, then its SpeedyStatSplit -- the compiled cparsers.pyx -- per event; the
per-event loop of Experiment.parse is spread over a fork pool, which changes nothing about what is computed).  The
SHA-256 of its event rows (start, length) and segment rows (event, start, end) is compared with the oracle-made
fixture and both are recorded; tests/test_oracle_golden.py asserts that every fixture hash equals the reference's.
So the full-size fixtures are pinned to the real reference, not only to its restatement."""
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from pypore_b200 import dist as ppdist  # noqa: E402
from pypore_b200 import synth  # noqa: E402

RULES = [lambda e: e.duration > 1000, lambda e: e.min > -0.5, lambda e: e.max < 110]
_STATE = {}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _split_range(job):
    lo, hi = job
    parsers, events, kw = _STATE["parsers"], _STATE["events"], _STATE["kw"]
    rows = []
    for k in range(lo, hi):
        for seg in parsers.SpeedyStatSplit(**kw).parse(events[k].current):
            rows.append((k, seg.start, seg.end))
    return np.array(rows, np.int64).reshape(-1, 3)


def through_reference(dt, parsers, x64, kw, grain):
    t0 = time.time()
    f = dt.File(current=x64, timestep=0.01)
    f.parse(parser=parsers.lambda_event_parser(threshold=110, rules=RULES))
    ev = np.array([(int(round(e.start * f.second)), len(e.current)) for e in f.events], np.int64).reshape(-1, 2)
    _STATE.update(parsers=parsers, events=f.events, kw=kw)
    jobs = [(a, min(a + grain, len(ev))) for a in range(0, len(ev), grain)]
    with mp.get_context("fork").Pool(os.cpu_count() or 1) as pool:      # forked: the events are shared, not pickled
        parts = pool.map(_split_range, jobs, chunksize=1)
    rows = np.concatenate(parts) if parts else np.zeros((0, 3), np.int64)
    _STATE.clear()
    return dict(samples=int(len(x64)), events=int(len(ev)), event_samples=int(ev[:, 1].sum()), segments=int(len(rows)),
                events_sha=sha(ev), segments_sha=sha(rows), seconds=round(time.time() - t0, 1))


def c5_through_reference(dt, parsers):
    """BASELINE configs[4]'s eight distinct 250 kHz files: File.parse, then per event the reference's own
    Event.filter(1, 2000) -- scipy's bessel + filtfilt -- and SpeedyStatSplit on the filtered current; both gains."""
    fs = 2.5e5
    settings = {"psps10": dict(min_width=100, window_width=10000, sampling_freq=fs, cutoff_freq=2000.,
                               prior_segments_per_second=10),
                "default": dict(min_width=100, window_width=10000)}
    t0 = time.time()
    ev_rows, seg_rows = [], {k: [] for k in settings}
    samples = 0
    for i in range(8):
        x64 = synth.make_trace(208, seed=900 + i, tier="A").astype(np.float64)
        samples += len(x64)
        f = dt.File(current=x64, timestep=1000. / fs)
        f.parse(parser=parsers.lambda_event_parser(threshold=110, rules=RULES))
        for k, event in enumerate(f.events):
            ev_rows.append((i, int(round(event.start * f.second)), len(event.current)))
            event.filter(order=1, cutoff=2000.)
            for name, kw in settings.items():
                for seg in parsers.SpeedyStatSplit(**kw).parse(event.current):
                    seg_rows[name].append((i, k, seg.start, seg.end))
    ev = np.array(ev_rows, np.int64).reshape(-1, 3)
    out = {}
    for name in settings:
        rows = np.array(seg_rows[name], np.int64).reshape(-1, 4)
        out[name] = dict(samples=int(samples), events=int(len(ev)), segments=int(len(rows)), events_sha=sha(ev),
                         segments_sha=sha(rows), seconds=round(time.time() - t0, 1))
    return out


def main():
    import make_golden
    dt, parsers, _ = make_golden.load_reference()
    fix_b = np.load(os.path.join(HERE, "bench_configs.npz"), allow_pickle=False)
    fix_s = np.load(os.path.join(HERE, "sharded_full.npz"), allow_pickle=False)
    default = dict(min_width=100, window_width=10000)
    work = [
        ("c4", fix_b, "c4_", lambda: synth.make_long_trace(20, 10_000_000, seed0=100, tier="A"),
         dict(min_width=100, max_width=1000000, window_width=10000), 1),
        ("c3", fix_b, "c3_", lambda: synth.make_trace(30000, seed=2, tier="A"), default, 250),
        ("w2", fix_s, "w2_", lambda: ppdist.synthetic_global(2, 5000, seed0=1), default, 250),
        ("w4", fix_s, "w4_", lambda: ppdist.synthetic_global(4, 5000, seed0=1), default, 250),
        ("w8", fix_s, "w8_", lambda: ppdist.synthetic_global(8, 5000, seed0=1), default, 250),
    ]
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    path = os.path.join(HERE, "reference_full_check.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for name, fix, key, make, kw, grain in work:
        if only and name not in only:
            continue
        x64 = np.asarray(make(), np.float64)
        r = through_reference(dt, parsers, x64, kw, grain)
        del x64
        r["fixture_events_sha"], r["fixture_segments_sha"] = str(fix[key + "events_sha"]), str(fix[key + "segments_sha"])
        r["matches_fixture"] = bool(r["events_sha"] == r["fixture_events_sha"] and
                                    r["segments_sha"] == r["fixture_segments_sha"] and
                                    r["events"] == int(fix[key + "events"]) and r["segments"] == int(fix[key + "segments"]))
        out[name] = r
        print(name, {k: v for k, v in r.items() if not k.endswith("sha")}, flush=True)
        json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    if not only or "c5" in only:
        fix = np.load(os.path.join(HERE, "c5_files.npz"), allow_pickle=False)
        for name, r in c5_through_reference(dt, parsers).items():
            r["fixture_events_sha"] = str(fix[name + "_events_sha"])
            r["fixture_segments_sha"] = str(fix[name + "_segments_sha"])
            r["matches_fixture"] = bool(r["events_sha"] == r["fixture_events_sha"] and
                                        r["segments_sha"] == r["fixture_segments_sha"])
            out["c5_" + name] = r
            print("c5", name, {k: v for k, v in r.items() if not k.endswith("sha")}, flush=True)
        json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    assert all(v["matches_fixture"] for v in out.values()), "the reference disagrees with a fixture"


if __name__ == "__main__":
    main()
