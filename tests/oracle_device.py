"""A stand-in for pypore_b200._lib.Context whose calls are answered by the CPU oracle (TEST INFRASTRUCTURE).

It lets the host layer -- the plug-in classes, their argument handling, what they hand back -- run where there is no
GPU: the build container, where the reference itself is available to call them (tests/test_integration_reference.py).
Only the calls the parsers make for one trace / a list of events are provided; the product never sees this class."""
import numpy as np

import oracle


class OracleDevice(object):
    def __init__(self):
        self.trace = None
        self.ev_arrays = None

    # -- trace / K1 ----------------------------------------------------------------------------------
    def upload_trace(self, x32, extra_capacity=0):
        assert x32.dtype == np.float32
        self.trace = x32.astype(np.float64)
        return len(x32)

    def upload_trace_f64(self, x64):
        assert x64.dtype == np.float64
        self.trace = np.array(x64)
        return len(x64)

    def threshold_scan(self, threshold, scan_len=-1):
        self.run_table = oracle.threshold_runs(self.trace, threshold)
        return len(self.run_table[0])

    def runs(self, n_runs):
        return self.run_table

    def select_events(self, rule_mask, duration_gt=0, duration_lt=0, min_gt=0.0, max_lt=0.0, skip_first=False,
                      skip_last=False):
        start, length, mn, mx, _ = self.run_table
        keep = np.ones(len(start), bool)
        if rule_mask & 1:
            keep &= length > duration_gt
        if rule_mask & 8:
            keep &= length < duration_lt
        if rule_mask & 2:
            keep &= mn > min_gt
        if rule_mask & 4:
            keep &= mx < max_lt
        self.set_events(start[keep], length[keep])
        return int(keep.sum()), int(length[keep].sum())

    def set_events(self, start, length):
        self.ev = (np.asarray(start, np.int64), np.asarray(length, np.int64))
        self.ev_arrays = [self.trace[s:s + n] for s, n in zip(*self.ev)]

    def events(self, n_events):
        return self.ev

    # -- events given as arrays / K2..K5 ---------------------------------------------------------------
    def upload_events_f64(self, arrays):
        self.ev_arrays = [np.asarray(a, np.float64) for a in arrays]
        lens = np.asarray([len(a) for a in arrays], np.int64)
        self.ev = (np.concatenate(([0], np.cumsum(lens)[:-1])), lens)
        return lens

    def filter_events(self, b, a, zi):
        self.ev_arrays = [oracle.filtfilt(np.asarray(b), np.asarray(a), x) for x in self.ev_arrays]

    def event_samples(self, n_samples):
        return np.concatenate(self.ev_arrays)

    def statsplit(self, min_width, max_width, window_width, min_gain, prefix_mode=0):
        rows = []
        for e, x in enumerate(self.ev_arrays):
            bp = oracle.statsplit(x, min_width=min_width, max_width=max_width, window_width=window_width, gain=min_gain)
            edges = np.concatenate(([0], bp, [len(x)])).astype(np.int64)
            rows.append(np.stack([np.full(len(edges) - 1, e, np.int64), edges[:-1], edges[1:]], axis=1))
        self.rows = np.concatenate(rows, axis=0) if rows else np.zeros((0, 3), np.int64)
        return len(self.rows)

    def segment_stats(self):
        pass

    def segments(self, n_segments, stats=True, pinned=False):
        out = dict(event=self.rows[:, 0].astype(np.int32), start=self.rows[:, 1], end=self.rows[:, 2])
        if stats:
            cols = [oracle.segment_stats(self.ev_arrays[e], [a], [b]) for e, a, b in self.rows]
            for j, k in enumerate(("mean", "std", "min", "max")):
                out[k] = np.asarray([c[j][0] for c in cols], np.float64)
        return out

    def event_stats(self, n_events):
        f = dict(mean=np.mean, std=np.std, min=np.min, max=np.max)
        return {k: np.asarray([fn(x) for x in self.ev_arrays], np.float64) for k, fn in f.items()}
