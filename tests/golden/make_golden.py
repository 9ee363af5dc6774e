"""Generate tests/golden/*.npz by running the REAL reference (/root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference is Python-2 source; it is exec'd from where it lies with the
in-memory textual shims of SURVEY.md App. B (iteritems->items, xrange->range,
izip->zip, print statements) and stub modules for its absent optional imports.
Nothing from the reference is written into this repository except the numeric
outputs stored in the fixtures.  Inputs are regenerated from
pypore_b200.synth by seed; a sha256 of the input bytes guards generator drift.
"""
import hashlib
import itertools
import os
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"

sys.path.insert(0, os.path.dirname(HERE))
from conftest import SCORING_CASES  # noqa: E402
from oracle import build_oracle  # noqa: E402
from pypore_b200 import synth  # noqa: E402


def _shim(src):
    src = src.replace("iteritems()", "items()").replace("xrange", "range")
    src = src.replace("from itertools import tee,izip,chain", "from itertools import tee,chain; izip=zip")
    src = src.replace("from itertools import chain, izip, tee, combinations",
                      "from itertools import chain, tee, combinations; izip=zip")
    src = src.replace("tics = map( int, tics )", "tics = list( map( int, tics ) )")   # Python-2 map returned a list
    src = re.sub(r'^(\s*)print "(.*)"\.format\((.*)\)\s*$', r'\1print("\2".format(\3))', src, flags=re.M)
    src = re.sub(r'^(\s*)print "(.*)"\s*$', r'\1print("\2")', src, flags=re.M)
    return src


def load_reference():
    """Returns the reference's PyPore.DataTypes module, loaded per SURVEY App. B."""
    so = build_oracle.build_ref()
    itertools.izip = zip
    pkg = types.ModuleType("PyPore")
    pkg.__path__ = [os.path.dirname(so)]
    sys.modules["PyPore"] = pkg

    def exec_into(name, filename, aliases=(), strip=()):
        with open(os.path.join(REFERENCE, "PyPore", filename)) as f:
            src = _shim(f.read())
        for s in strip:
            src = src.replace(s, "")
        mod = types.ModuleType(name)
        mod.__file__ = os.path.join(REFERENCE, "PyPore", filename)
        sys.modules[name] = mod
        for a in aliases:
            sys.modules[a] = mod
        exec(compile(src, mod.__file__, "exec"), mod.__dict__)
        return mod

    core = exec_into("core", "core.py", aliases=("PyPore.core",))
    pkg.core = core
    import PyPore.cparsers  # noqa: F401  (compiled reference from oracle/_ref)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "MySQLdb", "hmm", "database",
                 "alignment", "yahmm"):
        m = types.ModuleType(name)
        m.__all__ = []
        sys.modules[name] = m
    sys.modules["yahmm"].Model = type("Model", (), {})
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    ra = types.ModuleType("read_abf")

    def read_abf(*a, **k):
        raise IOError("no .abf data offline")
    ra.read_abf = read_abf
    sys.modules["read_abf"] = ra
    parsers = exec_into("PyPore.parsers", "parsers.py", aliases=("parsers",),
                        strip=("import pyximport\n",
                               "pyximport.install( setup_args={'include_dirs':np.get_include()})\n"))
    pkg.parsers = parsers
    dt = exec_into("PyPore.DataTypes", "DataTypes.py", aliases=("DataTypes",))
    pkg.DataTypes = dt
    return dt, parsers, core


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


SPLIT_SETTINGS = {
    "default": dict(min_width=100, window_width=10000),
    "psps10": dict(min_width=100, window_width=10000, prior_segments_per_second=10),
    "narrow": dict(min_width=50, max_width=2500, window_width=1000, prior_segments_per_second=50),
}


def golden_pipeline(dt, parsers, tier, seed, n_events, fs=1.e5):
    """File.parse -> Event.parse for several parser settings; records tables in samples."""
    x32 = synth.make_trace(n_events, seed=seed, tier=tier)
    out = {"input_sha256": np.array(sha(x32)), "seed": seed, "n_events": n_events, "fs": fs,
           "tier": np.array(tier), "threshold": 110.0}
    x64 = x32.astype(np.float64)
    f = dt.File(current=x64, timestep=1000. / fs)
    rules = [lambda e: e.duration > 1000, lambda e: e.min > -0.5, lambda e: e.max < 110]
    f.parse(parser=parsers.lambda_event_parser(threshold=110, rules=rules))
    second = f.second
    out["second"] = second
    out["event_start_s"] = np.array([e.start for e in f.events])
    out["event_end_s"] = np.array([e.end for e in f.events])
    out["event_duration_s"] = np.array([e.duration for e in f.events])
    out["event_n"] = np.array([len(e.current) for e in f.events], np.int64)
    out["event_min"] = np.array([e.min for e in f.events])
    out["event_max"] = np.array([e.max for e in f.events])
    out["event_mean"] = np.array([e.mean for e in f.events])
    out["event_std"] = np.array([e.std for e in f.events])
    # default rules reject everything in this regime
    f2 = dt.File(current=x64, timestep=1000. / fs)
    f2.parse(parser=parsers.lambda_event_parser(threshold=110))
    out["n_events_default_rules"] = len(f2.events)
    for name, kw in SPLIT_SETTINGS.items():
        ev_id, st, en, mean, std, mn, mx, st_s, en_s = [], [], [], [], [], [], [], [], []
        p = parsers.SpeedyStatSplit(**kw)
        for i, e in enumerate(f.events):
            raw = p.parse(e.current)  # samples, before Event.parse rescales
            for s in raw:
                ev_id.append(i); st.append(s.start); en.append(s.end)
                mean.append(s.mean); std.append(s.std); mn.append(s.min); mx.append(s.max)
            e.parse(parser=p)  # seconds
            for s in e.segments:
                st_s.append(s.start); en_s.append(s.end)
        out[name + "_event"] = np.array(ev_id, np.int64)
        out[name + "_start"] = np.array(st, np.int64)
        out[name + "_end"] = np.array(en, np.int64)
        out[name + "_mean"] = np.array(mean)
        out[name + "_std"] = np.array(std)
        out[name + "_min"] = np.array(mn)
        out[name + "_max"] = np.array(mx)
        out[name + "_start_s"] = np.array(st_s)
        out[name + "_end_s"] = np.array(en_s)
    return out


def golden_filter(dt, parsers, fs, order, cutoff, seed):
    """Event.filter then SpeedyStatSplit on the first two events of a small trace."""
    x32 = synth.make_trace(3, seed=seed, tier="A")
    x64 = x32.astype(np.float64)
    f = dt.File(current=x64, timestep=1000. / fs)
    rules = [lambda e: e.duration > 1000, lambda e: e.min > -0.5, lambda e: e.max < 110]
    f.parse(parser=parsers.lambda_event_parser(threshold=110, rules=rules))
    out = {"input_sha256": np.array(sha(x32)), "seed": seed, "fs": fs, "order": order, "cutoff": cutoff}
    from scipy import signal
    b, a = signal.bessel(order, cutoff / (f.second / 2.), btype='low', analog=0, output='ba')
    out["b"], out["a"], out["zi"] = b, a, signal.lfilter_zi(b, a)
    p = parsers.SpeedyStatSplit(min_width=100, window_width=10000, sampling_freq=fs, cutoff_freq=cutoff,
                                prior_segments_per_second=10)
    for i, e in enumerate(f.events[:2]):
        e.filter(order=order, cutoff=cutoff)
        assert e.filtered and e.current.dtype == np.float64
        out["event%d_start" % i] = int(round(e.start * f.second))
        out["event%d_filtered" % i] = e.current.copy()
        raw = p.parse(e.current)
        out["event%d_seg_start" % i] = np.array([s.start for s in raw], np.int64)
        out["event%d_seg_end" % i] = np.array([s.end for s in raw], np.int64)
        out["event%d_seg_mean" % i] = np.array([s.mean for s in raw])
        out["event%d_seg_std" % i] = np.array([s.std for s in raw])
    return out


def golden_long(parsers):
    """One 300k-sample event with forced max_width splits and the window chain (C4 regime, small)."""
    out = {}
    for name, kw in {"long_default": dict(min_width=100, max_width=20000, window_width=10000),
                     "long_psps10": dict(min_width=100, max_width=20000, window_width=10000,
                                         prior_segments_per_second=10),
                     "long_highgain": dict(min_width=100, max_width=15000, window_width=4000,
                                           min_gain_per_sample=2.0)}.items():
        x32 = synth.make_long_event(300000, seed=100, tier="A")
        raw = parsers.SpeedyStatSplit(**kw).parse(x32.astype(np.float64))
        out[name + "_start"] = np.array([s.start for s in raw], np.int64)
        out[name + "_end"] = np.array([s.end for s in raw], np.int64)
    out["input_sha256"] = np.array(sha(x32))
    return out


def golden_c4_full(dt, parsers):
    """BASELINE configs[3] at FULL size: 4 events of 10 M samples in one trace, SpeedyStatSplit with max_width=1e6,
    through the real reference (File.parse + the compiled cparsers per event).  Too big to store: the fixture holds
    the event table, the segment counts and a SHA-256 of the (event, start, end) int64 rows per setting."""
    x64 = synth.make_long_trace(4, 10_000_000, seed0=100, tier="A").astype(np.float64)
    f = dt.File(current=x64, timestep=0.01)
    f.parse(parser=parsers.lambda_event_parser(threshold=110, rules=[lambda e: e.duration > 1000,
                                                                      lambda e: e.min > -0.5,
                                                                      lambda e: e.max < 110]))
    out = dict(ev_start=np.array([int(round(e.start * f.second)) for e in f.events], np.int64),
               ev_len=np.array([len(e.current) for e in f.events], np.int64), input_sha256=np.array(sha(x64)))
    for name, kw in (("default", dict(min_width=100, max_width=1000000, window_width=10000)),
                     ("psps10", dict(min_width=100, max_width=1000000, window_width=10000,
                                     prior_segments_per_second=10))):
        rows = []
        for k, event in enumerate(f.events):
            for seg in parsers.SpeedyStatSplit(**kw).parse(event.current):
                rows.append((k, seg.start, seg.end))
        rows = np.array(rows, np.int64).reshape(-1, 3)
        out[name + "_segments"] = np.int64(len(rows))
        out[name + "_sha"] = np.array(sha(rows))
    return out


def golden_c2_full(dt, parsers):
    """BASELINE configs[1] at FULL size -- exactly bench.py's single-GPU workload, synth.make_trace(5000, seed=1),
    about 60 M samples -- through the real reference: File.parse(lambda_event_parser(threshold=110, rules=[duration >
    1000, min > -0.5, max < 110])) then SpeedyStatSplit(min_width=100, window_width=10000) per event.  Counts and
    SHA-256 of the event rows (start, length) and of the segment rows (event, start, end)."""
    x64 = synth.make_trace(5000, seed=1, tier="A").astype(np.float64)
    f = dt.File(current=x64, timestep=0.01)
    f.parse(parser=parsers.lambda_event_parser(threshold=110, rules=[lambda e: e.duration > 1000,
                                                                      lambda e: e.min > -0.5,
                                                                      lambda e: e.max < 110]))
    ev = np.array([(int(round(e.start * f.second)), len(e.current)) for e in f.events], np.int64).reshape(-1, 2)
    out = dict(samples=np.int64(len(x64)), events=np.int64(len(ev)), event_samples=np.int64(ev[:, 1].sum()),
               events_sha=np.array(sha(ev)), input_sha256=np.array(sha(x64)))
    for name, kw in (("default", dict(min_width=100, window_width=10000)),
                     ("psps10", dict(min_width=100, window_width=10000, prior_segments_per_second=10))):
        rows = []
        for k, event in enumerate(f.events):
            for seg in parsers.SpeedyStatSplit(**kw).parse(event.current):
                rows.append((k, seg.start, seg.end))
        rows = np.array(rows, np.int64).reshape(-1, 3)
        out[name + "_segments"] = np.int64(len(rows))
        out[name + "_sha"] = np.array(sha(rows))
    return out


def golden_params(parsers):
    """min_gain known answers and exception parity."""
    from PyPore.cparsers import FastStatSplit
    out = {
        "min_gain_default": FastStatSplit().min_gain,
        "min_gain_psps10": FastStatSplit(prior_segments_per_second=10).min_gain,
        "min_gain_psps10_cut2000": FastStatSplit(prior_segments_per_second=10, cutoff_freq=2000.).min_gain,
        "min_gain_per_sample_0p5": FastStatSplit(min_gain_per_sample=0.5).min_gain,
        "min_gain_fpr50_fs250k": FastStatSplit(false_positive_rate=50., sampling_freq=2.5e5,
                                               prior_segments_per_second=25.).min_gain,
    }
    try:
        FastStatSplit().parse(np.zeros(1000, np.float32))
        out["float32_error"] = np.array("")
    except ValueError as e:
        out["float32_error"] = np.array(str(e))
    return out


def golden_scoring():
    """FastStatSplit.best_single_split / score_samples of the compiled reference (cparsers.pyx:120-155,205-275)."""
    from PyPore.cparsers import FastStatSplit
    out = {}
    for name, (length, seed, tier, kw) in SCORING_CASES.items():
        x = synth.make_long_event(length, seed=seed, tier=tier).astype(np.float64)
        out[name + "_input_sha256"] = np.array(sha(x))
        g, i = FastStatSplit(**kw).best_single_split(x)
        out[name + "_best_gain"], out[name + "_best_index"] = g, i
        out[name + "_no_split"] = np.array(FastStatSplit(**kw).score_samples(x, no_split=True))
        out[name + "_scores"] = np.array(FastStatSplit(**kw).score_samples(x))
    for n in (0, 3, 5, 6, 7):   # too short for a candidate / the first lengths with one
        x = synth.make_long_event(50, seed=14, tier="A").astype(np.float64)[:n]
        with np.errstate(all="ignore"):
            g, i = FastStatSplit().best_single_split(x)
        out["short%d_best" % n] = np.array([g, i])
    return out


from make_golden_cases import FDS_CASES  # noqa: E402  (shared with tests/test_gpu_api.py)


def golden_fds(parsers):
    """FilterDerivativeSegmenter.parse (parsers.py:609-656) of the real reference, quirks included (np.argmax
    compared with high_threshold; a Segment for every SECOND pair of tics only)."""
    out = {}
    for name, (length, seed, tier, kw) in FDS_CASES.items():
        x = synth.make_long_event(length, seed=seed, tier=tier).astype(np.float64)
        segs = parsers.FilterDerivativeSegmenter(**kw).parse(x)
        out[name + "_start"] = np.array([s.start for s in segs], np.int64)
        out[name + "_n"] = np.array([len(s.current) for s in segs], np.int64)
        out[name + "_mean"] = np.array([s.mean for s in segs], np.float64)
    return out


def golden_json(dt, parsers):
    """File.to_json of the real reference after File.parse -> Event.filter -> Event.parse (DataTypes.py:708-738),
    the on-disk format of the result tables."""
    x64 = synth.make_trace(4, seed=21, tier="A").astype(np.float64)
    f = dt.File(current=x64, timestep=0.01)
    f.parse(parser=parsers.lambda_event_parser(threshold=110, rules=[lambda e: e.duration > 1000,
                                                                      lambda e: e.min > -0.5,
                                                                      lambda e: e.max < 110]))
    seg = parsers.SpeedyStatSplit(min_width=100, window_width=10000, prior_segments_per_second=10,
                                  cutoff_freq=2000.)
    for i, event in enumerate(f.events):
        if i % 2 == 0:
            event.filter(1, 2000.)
        event.parse(parser=seg)
    return f.to_json()


def golden_experiment(dt, parsers):
    """The real reference's batch driver, Experiment.parse(..., meta=True) (DataTypes.py:956-988), on the
    file-batch test set of tests/batch_common.py (250 kHz, Event.filter(1, 2000) + SpeedyStatSplit per event):
    every file's to_json after the run plus everything the run printed.  The only stand-ins are I/O:
    read_abf (no .abf data offline) hands out the in-memory traces, itertools.imap is Python 3's map."""
    import contextlib
    import io
    import json
    from batch_common import FS, TIMESTEP, edge_files, make_files
    traces = make_files(4) + edge_files()[1:4]
    names = ["batch%d.abf" % i for i in range(len(traces))]
    store = {n: x.astype(np.float64) for n, x in zip(names, traces)}
    dt.read_abf = lambda filename: (TIMESTEP, store[filename])
    itertools.imap = map
    exp = dt.Experiment(names)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        exp.parse(event_detector=parsers.lambda_event_parser(threshold=110, rules=[lambda e: e.duration > 1000,
                                                                                    lambda e: e.min > -0.5,
                                                                                    lambda e: e.max < 110]),
                  segmenter=parsers.SpeedyStatSplit(min_width=100, window_width=10000, sampling_freq=FS,
                                                    cutoff_freq=2000., prior_segments_per_second=10),
                  filter_params=(1, 2000.), verbose=True, meta=True)
    return json.dumps(dict(stdout=out.getvalue(), files=[json.loads(f.to_json()) for f in exp.files]), indent=1)


def main():
    dt, parsers, core = load_reference()
    if "--c2-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "c2_full.npz"), **golden_c2_full(dt, parsers))
        print("c2_full.npz", os.path.getsize(os.path.join(HERE, "c2_full.npz")))
        return
    if "--c4-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "c4_full.npz"), **golden_c4_full(dt, parsers))
        print("c4_full.npz", os.path.getsize(os.path.join(HERE, "c4_full.npz")))
        return
    if "--experiment-only" in sys.argv:
        with open(os.path.join(HERE, "experiment_meta.json"), "w") as out:
            out.write(golden_experiment(dt, parsers))
        print("experiment_meta.json", os.path.getsize(os.path.join(HERE, "experiment_meta.json")))
        return
    if "--json-only" in sys.argv:
        with open(os.path.join(HERE, "file_tierA.json"), "w") as out:
            out.write(golden_json(dt, parsers))
        print("file_tierA.json", os.path.getsize(os.path.join(HERE, "file_tierA.json")))
        return
    if "--fds-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "fds.npz"), **golden_fds(parsers))
        print("fds.npz", os.path.getsize(os.path.join(HERE, "fds.npz")))
        return
    if "--scoring-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "scoring.npz"), **golden_scoring())
        print("scoring.npz", os.path.getsize(os.path.join(HERE, "scoring.npz")))
        return
    np.savez_compressed(os.path.join(HERE, "scoring.npz"), **golden_scoring())
    with open(os.path.join(HERE, "file_tierA.json"), "w") as out:
        out.write(golden_json(dt, parsers))
    np.savez_compressed(os.path.join(HERE, "pipeline_tierA.npz"), **golden_pipeline(dt, parsers, "A", 3, 12))
    np.savez_compressed(os.path.join(HERE, "pipeline_tierB.npz"), **golden_pipeline(dt, parsers, "B", 4, 8))
    np.savez_compressed(os.path.join(HERE, "filter_o1_100k.npz"), **golden_filter(dt, parsers, 1.e5, 1, 2000., 5))
    np.savez_compressed(os.path.join(HERE, "filter_o1_250k.npz"), **golden_filter(dt, parsers, 2.5e5, 1, 2000., 6))
    np.savez_compressed(os.path.join(HERE, "filter_o2_100k.npz"), **golden_filter(dt, parsers, 1.e5, 2, 2000., 5))
    np.savez_compressed(os.path.join(HERE, "filter_o4_100k.npz"), **golden_filter(dt, parsers, 1.e5, 4, 5000., 5))
    np.savez_compressed(os.path.join(HERE, "long_event.npz"), **golden_long(parsers))
    np.savez_compressed(os.path.join(HERE, "params.npz"), **golden_params(parsers))
    np.savez_compressed(os.path.join(HERE, "fds.npz"), **golden_fds(parsers))
    with open(os.path.join(HERE, "experiment_meta.json"), "w") as out:
        out.write(golden_experiment(dt, parsers))
    np.savez_compressed(os.path.join(HERE, "c4_full.npz"), **golden_c4_full(dt, parsers))
    np.savez_compressed(os.path.join(HERE, "c2_full.npz"), **golden_c2_full(dt, parsers))
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
