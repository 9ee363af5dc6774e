"""Shared by the file-batch tests (CPU gloo / GPU): a small C5-style batch (BASELINE configs[4]: 250 kHz
files, Event.filter(1, 2000) then SpeedyStatSplit) and the oracle's version of one file's pass -- the
reference's Experiment.parse loop body (DataTypes.py:972-982) stated over tables."""
import numpy as np

import oracle
from pypore_b200 import synth
from pypore_b200.parsers import RuleSet, SpeedyStatSplit, lambda_event_parser

FS = 2.5e5
TIMESTEP = 1000. / FS
FILTER = (1, 2000.)
PYRULES = [lambda e: e.duration > 1000, lambda e: e.min > -0.5, lambda e: e.max < 110]


def detector():
    return lambda_event_parser(threshold=110, rules=RuleSet(duration_gt=1000, min_gt=-0.5, max_lt=110))


def segmenter():
    return SpeedyStatSplit(min_width=100, window_width=10000, sampling_freq=FS, cutoff_freq=2000.,
                           prior_segments_per_second=10)


def make_files(n_files, seed0=300):
    """Ragged batch: 2..7 events per file; one file without any event (open channel only)."""
    files = []
    for i in range(n_files):
        if i == 2:
            rng = np.random.RandomState(seed0 + i)
            files.append(synth.quantise(rng.normal(synth.OPEN_MEAN, synth.OPEN_STD, 7000)))
        else:
            files.append(synth.make_trace(2 + (i * 3) % 6, seed=seed0 + i, tier="A"))
    return files


def edge_files():
    """Files that start and end every way: inside an event, in the open channel, with an event cut to less than
    the duration rule by the file boundary, a file that is one long event, a file without any event."""
    from pypore_b200 import synth
    base = [synth.make_trace(3, seed=410 + i, tier="A") for i in range(6)]
    first_below = [int(np.argmax(x < 110)) for x in base]
    files = [base[0],
             base[1][first_below[1] + 2500:],                      # starts inside an event
             base[2][:first_below[2] + 3000],                      # ends inside an event
             base[3][first_below[3] + 500:first_below[3] + 4000],  # one run below the threshold, nothing else
             synth.quantise(np.random.RandomState(5).normal(120, 1.5, 5000)),   # open channel only
             base[4][first_below[4] + 5200:],                      # may start with a short tail of an event
             base[5][:first_below[5] + 600]]                       # ends with a piece shorter than the duration rule
    return files


def oracle_file_pass(ctx, current, second, event_detector, seg, filter_params):
    """Same contract as pypore_b200.batch.device_file_pass, computed by the oracle (tests only)."""
    x = np.asarray(current, np.float64)
    ws, wl = oracle.events(x, event_detector.threshold, PYRULES)
    mw, MW, W, gain = seg._params()
    ev_flt, seg_int, seg_flt = [], [], []
    for e, (s, n) in enumerate(zip(ws, wl)):
        cur = x[s:s + n]
        if filter_params is not None:
            cur = oracle.event_filter(cur, second, *filter_params)
        bp = oracle.statsplit(cur, min_width=mw, max_width=MW, window_width=W, gain=gain)
        edges = np.concatenate(([0], bp, [n])).astype(np.int64)
        m, sd, mn, mx = oracle.segment_stats(cur, edges[:-1], edges[1:])
        seg_int.append(np.stack([np.full(len(edges) - 1, e, np.int64), edges[:-1], edges[1:]], axis=1))
        seg_flt.append(np.stack([m, sd, mn, mx], axis=1))
        ev_flt.append([np.mean(cur), np.std(cur), np.min(cur), np.max(cur)])
    E = len(ws)
    return dict(ev_int=np.stack([ws, wl], axis=1).astype(np.int64).reshape(E, 2),
                ev_flt=np.asarray(ev_flt, np.float64).reshape(E, 4),
                seg_int=np.concatenate(seg_int, axis=0) if E else np.zeros((0, 3), np.int64),
                seg_flt=np.concatenate(seg_flt, axis=0) if E else np.zeros((0, 4)))


def oracle_tables(files):
    """The whole batch through the oracle on one process: what every rank must end up with."""
    from pypore_b200.batch import FileBatch
    b = FileBatch(workers=1, file_pass=oracle_file_pass)
    return b.parse(files, TIMESTEP, detector(), segmenter(), FILTER)


def assert_tables_match(t, want, rtol_stats=1e-9, exact=False):
    """Indices bit-exact; statistics within the north star's 1e-9 (1e-5 for filtered samples' extrema)."""
    for k in ("file", "start", "length"):
        assert np.array_equal(t.events[k], want.events[k]), "event column %s differs" % k
    for k in ("file", "event", "start", "end"):
        assert np.array_equal(t.segments[k], want.segments[k]), "segment column %s differs" % k
    for tab, wtab in ((t.events, want.events), (t.segments, want.segments)):
        for k in ("mean", "std", "min", "max"):
            if exact:
                assert np.array_equal(tab[k], wtab[k]), k
            else:
                tol = rtol_stats if k in ("mean", "std") else 1e-5
                assert np.allclose(tab[k], wtab[k], rtol=tol, atol=0), k


class OracleContext(object):
    """A stand-in for _lib.Context in the CPU tests of pypore_b200.batch.device_group_pass / device_file_pass:
    the same calls (trace upload / append by raw pointer, pipeline on the resident trace, table downloads),
    answered by the oracle.  It sees exactly what the device would: the concatenated trace WITH the +inf
    separators, so the grouping argument itself (no run below the threshold crosses a file boundary, runs
    above never become events) is what gets tested."""

    def __init__(self):
        self.trace = np.zeros(0, np.float32)

    def pinned_empty(self, n, dtype):
        return np.empty(n, dtype)

    def upload_trace_async(self, x32, extra_capacity=0):
        assert x32.dtype == np.float32
        self.trace = np.array(x32)
        self.cap = len(x32) + extra_capacity

    def append_trace(self, src_ptr, n, src_is_device):
        import ctypes
        assert not src_is_device and len(self.trace) + n <= self.cap
        src = np.ctypeslib.as_array((ctypes.c_float * int(n)).from_address(int(src_ptr)))
        self.trace = np.concatenate([self.trace, np.array(src)])

    @property
    def trace_len(self):
        return len(self.trace)

    def pipeline(self, threshold, rule_mask, duration_gt, duration_lt, min_gt, max_lt, min_width, max_width,
                 window_width, min_gain, filter_ba=None, with_stats=True, host_trace=None):
        if host_trace is not None:
            self.upload_trace_async(np.ascontiguousarray(host_trace, np.float32))
        x = self.trace.astype(np.float64)
        rules = []
        if rule_mask & 1:
            rules.append(lambda e: e.duration > duration_gt)
        if rule_mask & 8:
            rules.append(lambda e: e.duration < duration_lt)
        if rule_mask & 2:
            rules.append(lambda e: e.min > min_gt)
        if rule_mask & 4:
            rules.append(lambda e: e.max < max_lt)
        ws, wl = oracle.events(x, threshold, rules or [lambda e: True])
        self.ev = (np.asarray(ws, np.int64), np.asarray(wl, np.int64))
        est, rows, flt = [], [], []
        for e, (s, n) in enumerate(zip(ws, wl)):
            cur = x[s:s + n]
            if filter_ba is not None:
                cur = oracle.filtfilt(filter_ba[0], filter_ba[1], cur)
            bp = oracle.statsplit(cur, min_width=min_width, max_width=max_width, window_width=window_width,
                                  gain=min_gain)
            edges = np.concatenate(([0], bp, [n])).astype(np.int64)
            m, sd, mn, mx = oracle.segment_stats(cur, edges[:-1], edges[1:])
            rows.append(np.stack([np.full(len(edges) - 1, e, np.int64), edges[:-1], edges[1:]], axis=1))
            flt.append(np.stack([m, sd, mn, mx], axis=1))
            est.append([np.mean(cur), np.std(cur), np.min(cur), np.max(cur)])
        self.est = np.asarray(est, np.float64).reshape(len(ws), 4)
        self.rows = np.concatenate(rows, axis=0) if rows else np.zeros((0, 3), np.int64)
        self.flt = np.concatenate(flt, axis=0) if flt else np.zeros((0, 4))
        return dict(runs=0, events=len(ws), event_samples=int(np.sum(wl)), segments=len(self.rows))

    def events(self, n):
        return self.ev

    def event_stats(self, n):
        return {k: self.est[:, j] for j, k in enumerate(("mean", "std", "min", "max"))}

    def segments(self, n):
        out = dict(event=self.rows[:, 0].astype(np.int32), start=self.rows[:, 1], end=self.rows[:, 2])
        out.update({k: self.flt[:, j] for j, k in enumerate(("mean", "std", "min", "max"))})
        return out


def golden_experiment_inputs():
    """The traces tests/golden/make_golden.py::golden_experiment fed to the real reference's Experiment.parse."""
    traces = make_files(4) + edge_files()[1:4]
    return traces, ["batch%d" % i for i in range(len(traces))]


def assert_json_close(a, b, rtol, path="", time_rtol=1e-12):
    """Same keys, same structure, same strings / ints; statistics (mean / std / min / max) within rtol, every
    other float (start / end / duration in seconds, parser parameters) within time_rtol."""
    if isinstance(a, dict):
        assert isinstance(b, dict) and sorted(a) == sorted(b), "%s: keys %s != %s" % (path, sorted(a), sorted(b))
        for k in a:
            assert_json_close(a[k], b[k], rtol, path + "/" + str(k), time_rtol)
    elif isinstance(a, list):
        assert isinstance(b, list) and len(a) == len(b), "%s: %d != %d items" % (path, len(a), len(b))
        for i, (x, y) in enumerate(zip(a, b)):
            assert_json_close(x, y, rtol, "%s[%d]" % (path, i), time_rtol)
    elif isinstance(a, float) or isinstance(b, float):
        tol = rtol if path.rsplit("/", 1)[-1] in ("mean", "std", "min", "max") else time_rtol
        assert abs(a - b) <= tol * max(abs(a), abs(b)), "%s: %r != %r" % (path, a, b)
    else:
        assert a == b, "%s: %r != %r" % (path, a, b)


def experiment_through_batch(batch, capsys=None):
    """Experiment.parse(..., meta=True, batch=batch) on the golden inputs -> (list of per-file JSON dicts, stdout)."""
    import contextlib
    import io
    import json
    from pypore_b200.DataTypes import Experiment, File
    traces, names = golden_experiment_inputs()
    files = [File(current=x, timestep=TIMESTEP) for x in traces]
    for f, n in zip(files, names):
        f.filename = n
    exp = Experiment(files)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        exp.parse(event_detector=detector(), segmenter=segmenter(), filter_params=FILTER, verbose=True, meta=True,
                  batch=batch)
    return [json.loads(f.to_json()) for f in exp.files], out.getvalue()
