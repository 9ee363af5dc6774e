"""Batches of files: BASELINE configs[4] (Bessel-prefiltered pipeline over a batch of files at 1/2/4/8 GPUs).

SURVEY 8e: "Files are the unit for C5, with no halo."  The reference's caller is ``Experiment.parse``
(PyPore/DataTypes.py:956-988): a Python loop over files, then over events (``event.filter``,
``event.parse``), then -- with ``meta=True`` -- ``file.to_meta()``, which throws every sample away and keeps
the event / segment metadata (DataTypes.py:480-491,683-693; core.py:237-249).  Here

* a rank owns the files ``rank, rank + world, ...`` (``assign_files``);
* a pass is device-resident from the threshold scan to the tables (scan -> select -> Bessel filtfilt -> prefix
  sums -> split search -> compaction -> segment and event statistics) and serves SEVERAL files where the rules
  allow it (``groupable``: the files sit behind each other in one trace, a +inf sample between them), else one;
* a rank drives ``workers`` contexts, each with its own CUDA stream and its own host thread (ctypes releases
  the GIL), so the host->device copies of one pass, the table read-back of another and the kernels of a third
  overlap;
* nothing but the compact tables leaves the device: no filtered current is downloaded, because ``meta=True``
  discards it anyway;
* the ranks' tables are all-gathered (two ragged gathers: event rows, segment rows) and ordered by file, so
  every rank ends with the table one GPU would have produced, row for row.

``BatchTables.files()`` turns the tables into what the reference's loop leaves behind after ``meta=True``:
``File`` objects holding ``MetaEvent``s with ``MetaSegment``s, times in seconds.

There is no CPU fallback.  ``file_pass`` exists so that the CPU tests (gloo, world_size 2) can exercise the
assignment / gather / ordering logic with the oracle standing in for the device pass.
"""
import threading

import numpy as np

from . import _lib
from .parsers import SpeedyStatSplit, lambda_event_parser, _device_trace, _run_pipeline

STAT_KEYS = ("mean", "std", "min", "max")
SPLIT_WAVE = 148 * 7   # persistent CTAs of one full k3_split wave on a B200 (csrc/split.cuh: 7 CTAs per SM)


def assign_files(n_files, rank, world):
    """Indices of the files rank `rank` of `world` owns: round-robin, so that ragged file sizes and any
    trend along the list (longer recordings late in a session) spread over the ranks."""
    return list(range(int(rank), int(n_files), int(world)))


def device_file_pass(ctx, current, second, event_detector, segmenter, filter_params):
    """One file on one context.  Returns dict(ev_int [E,2] {start, length} in samples,
    ev_flt [E,4] {mean,std,min,max} of the (filtered) event, seg_int [S,3] {event, start, end} with
    event-relative sample indices, seg_flt [S,4])."""
    from .DataTypes import bessel_coefficients
    rs = event_detector._device_rules()
    if rs is None:
        raise TypeError("the batched pipeline needs device-evaluable rules (parsers.RuleSet)")
    if not isinstance(segmenter, SpeedyStatSplit):
        raise TypeError("the batched pipeline needs a pypore_b200 SpeedyStatSplit")
    mw, MW, W, gain = segmenter._params()
    filt = bessel_coefficients(filter_params[0], filter_params[1], second) if filter_params is not None else None
    if len(current) == 0:   # an empty recording has no events; it must not abort the batch
        return dict(ev_int=np.zeros((0, 2), np.int64), ev_flt=np.zeros((0, 4)), seg_int=np.zeros((0, 3), np.int64),
                    seg_flt=np.zeros((0, 4)))
    c = _run_pipeline(ctx, current, event_detector.threshold, min_width=mw, max_width=MW, window_width=W,
                      min_gain=gain, filter_ba=filt, with_stats=True, **rs.device_args())
    ne, ns = c["events"], c["segments"]
    ev_start, ev_len = ctx.events(ne)
    est = ctx.event_stats(ne)
    seg = ctx.segments(ns)
    return dict(ev_int=np.stack([ev_start, ev_len], axis=1).reshape(ne, 2),
                ev_flt=np.stack([est[k] for k in STAT_KEYS], axis=1).reshape(ne, 4),
                seg_int=np.stack([seg["event"].astype(np.int64), seg["start"], seg["end"]], axis=1).reshape(ns, 3),
                seg_flt=np.stack([seg[k] for k in STAT_KEYS], axis=1).reshape(ns, 4))


def groupable(event_detector):
    """Several files can share one resident pass when no run above the threshold can ever be an event: the
    files then sit behind each other in one device trace, separated by a single +inf sample.  A run below
    the threshold cannot cross the separator; runs above it may merge across files or pick up the separator,
    but `max < v` with v <= threshold rejects every one of them (max >= threshold, or NaN)."""
    rs = event_detector._device_rules()
    return (rs is not None and bool(rs.mask & _lib.RULE_MAX_LT)
            and float(rs.max_lt) <= float(event_detector.threshold))


def plan_groups(lengths, seconds, owned, group_samples, min_groups=1):
    """Consecutive owned files with the same sampling rate, at most `group_samples` samples per group (a file
    longer than that is a group of its own).  `min_groups` shrinks the budget so that every worker context
    gets work.  Returns a list of lists of file indices, in file order."""
    owned = list(owned)
    if group_samples <= 0:
        return [[i] for i in owned]
    total = sum(int(lengths[i]) for i in owned)
    budget = int(group_samples)
    if min_groups > 1 and total > 0:
        budget = max(1, min(budget, -(-total // int(min_groups))))
    groups, cur, used = [], [], 0
    for i in owned:
        n = int(lengths[i])
        if cur and (used + n + 1 > budget or seconds[i] != seconds[cur[0]] or n == 0 or lengths[cur[0]] == 0):
            groups.append(cur)
            cur, used = [], 0
        cur.append(i)
        used += n + 1
    if cur:
        groups.append(cur)
    return groups


def device_group_pass(ctx, currents, second, event_detector, segmenter, filter_params):
    """Several files in ONE device-resident pass (see `groupable`): the files are copied behind each other
    into the context's trace (pp_trace_upload + pp_trace_append, asynchronous from pinned memory), one launch of
    every stage serves all their events, and the tables are cut back into per-file results -- the same rows
    device_file_pass gives file by file, because every stage after the threshold scan works per event."""
    from .DataTypes import bessel_coefficients
    if not groupable(event_detector):
        raise TypeError("these rules could select a run above the threshold: files cannot share a pass")
    if not isinstance(segmenter, SpeedyStatSplit):
        raise TypeError("the batched pipeline needs a pypore_b200 SpeedyStatSplit")
    rs = event_detector._device_rules()
    mw, MW, W, gain = segmenter._params()
    filt = bessel_coefficients(filter_params[0], filter_params[1], second) if filter_params is not None else None
    xs = [_device_trace(c) for c in currents]
    if any(x.dtype != np.float32 for x in xs):
        # a float64 trace that float32 cannot hold goes up as float64 and runs on its own (the shared trace is float32)
        return [device_file_pass(ctx, c, second, event_detector, segmenter, filter_params) for c in currents]
    sep = getattr(ctx, "_batch_separator", None)
    if sep is None:
        sep = ctx._batch_separator = ctx.pinned_empty(1, np.float32)
        sep[0] = np.inf
    total = sum(x.shape[0] for x in xs) + len(xs) - 1
    ctx.upload_trace_async(xs[0], extra_capacity=total - xs[0].shape[0])
    offs = [0]
    for x in xs[1:]:
        ctx.append_trace(sep.ctypes.data, 1, False)
        offs.append(ctx.trace_len)
        ctx.append_trace(x.ctypes.data, x.shape[0], False)
    c = ctx.pipeline(event_detector.threshold, min_width=mw, max_width=MW, window_width=W, min_gain=gain,
                     filter_ba=filt, with_stats=True, **rs.device_args())     # ends with a sync: copies consumed
    del xs
    ne, ns = c["events"], c["segments"]
    ev_start, ev_len = ctx.events(ne)
    est = ctx.event_stats(ne)
    seg = ctx.segments(ns)
    offs = np.asarray(offs, np.int64)
    e_bounds = np.searchsorted(ev_start, np.concatenate((offs, [total + 1])))   # events are ordered by start
    s_bounds = np.searchsorted(seg["event"], e_bounds)
    ev_flt = np.stack([est[k] for k in STAT_KEYS], axis=1).reshape(ne, 4)
    seg_flt = np.stack([seg[k] for k in STAT_KEYS], axis=1).reshape(ns, 4)
    seg_event = seg["event"].astype(np.int64)
    out = []
    for k in range(len(currents)):
        e0, e1, s0, s1 = int(e_bounds[k]), int(e_bounds[k + 1]), int(s_bounds[k]), int(s_bounds[k + 1])
        out.append(dict(ev_int=np.stack([ev_start[e0:e1] - offs[k], ev_len[e0:e1]], axis=1).reshape(e1 - e0, 2),
                        ev_flt=ev_flt[e0:e1],
                        seg_int=np.stack([seg_event[s0:s1] - e0, seg["start"][s0:s1], seg["end"][s0:s1]],
                                         axis=1).reshape(s1 - s0, 3),
                        seg_flt=seg_flt[s0:s1]))
    return out


class BatchTables(object):
    """Event and segment tables of a batch of files.

    events:   dict(file, start, length [int64, samples], mean, std, min, max [float64])
    segments: dict(file, event [index inside the file], start, end [int64, event-relative samples],
              mean, std, min, max)
    Rows are ordered by (file, event start, segment start)."""

    def __init__(self, ev_int, ev_flt, seg_int, seg_flt, seconds, n_files):
        ev_int = np.asarray(ev_int, np.int64).reshape(-1, 3)
        seg_int = np.asarray(seg_int, np.int64).reshape(-1, 4)
        ev_flt = np.asarray(ev_flt, np.float64).reshape(-1, 4)
        seg_flt = np.asarray(seg_flt, np.float64).reshape(-1, 4)
        self.events = dict(file=ev_int[:, 0], start=ev_int[:, 1], length=ev_int[:, 2])
        self.segments = dict(file=seg_int[:, 0], event=seg_int[:, 1], start=seg_int[:, 2], end=seg_int[:, 3])
        for j, k in enumerate(STAT_KEYS):
            self.events[k] = ev_flt[:, j]
            self.segments[k] = seg_flt[:, j]
        self.seconds = np.asarray(seconds, np.float64)
        self.n_files = int(n_files)

    @property
    def n_events(self):
        return int(self.events["file"].shape[0])

    @property
    def n_segments(self):
        return int(self.segments["file"].shape[0])

    def file_rows(self, i):
        """(event row range, segment row range) of file i."""
        e = np.searchsorted(self.events["file"], [i, i + 1])
        s = np.searchsorted(self.segments["file"], [i, i + 1])
        return (int(e[0]), int(e[1])), (int(s[0]), int(s[1]))

    def files(self, filenames=None, event_detector=None, segmenter=None, filter_params=None):
        """The batch as the reference leaves it after Experiment.parse(..., meta=True): one File per file
        with MetaEvents / MetaSegments, start / end / duration in seconds (DataTypes.py:595-600,
        core.py:199-207), statistics from the device tables.  Every File is a lazy view on its rows of the
        tables (wire.FileTables): objects are built when somebody indexes them, to_json never builds any."""
        from . import wire
        from .DataTypes import File
        out = []
        ev, sg = self.events, self.segments
        for i in range(self.n_files):
            second = float(self.seconds[i])
            f = File(current=[], timestep=1000. / second)
            del f.current
            f.filename = filenames[i] if filenames is not None else ""
            (e0, e1), (s0, s1) = self.file_rows(i)
            rows_e = {k: ev[k][e0:e1] for k in ("start", "length") + STAT_KEYS}
            rows_s = {k: sg[k][s0:s1] for k in ("event", "start", "end") + STAT_KEYS}
            f._attach(wire.FileTables(second, rows_e, rows_s, filter_params, segmenter, meta=True))
            if event_detector is not None:
                f.event_parser = event_detector
            out.append(f)
        return out


class FileBatch(object):
    """A rank's share of a batch of files on one GPU, `workers` files in flight.

        batch = FileBatch(device=local_rank, workers=4, rank=rank, world=world)
        tables = batch.parse(traces, timestep, lambda_event_parser(...), SpeedyStatSplit(...), (1, 2000.))

    `traces` is the GLOBAL list (same length on every rank); a rank only touches the entries it owns, the
    others may be None.  `timestep` (ms per sample, like File(timestep=...)) is a scalar or one value per file.
    With world > 1 the result holds every rank's rows (torch.distributed must be initialised; NCCL moves the
    tables device to device, gloo -- the CPU tests -- host to host).
    """

    def __init__(self, device=0, workers=4, rank=0, world=1, group=None, file_pass=None, contexts=None,
                 share_split=True, group_samples=1 << 26, group_pass=None):
        self.rank, self.world, self.group = int(rank), int(world), group
        self.device = int(device)
        self.file_pass = file_pass or device_file_pass
        # files that may share a pass (`groupable`) are packed `group_samples` samples at a time; 0: one file per pass
        self.group_pass = group_pass or (device_group_pass if file_pass is None else None)
        self.group_samples = int(group_samples) if self.group_pass is not None else 0
        if contexts is not None:
            self.contexts = list(contexts)
        elif file_pass is None:
            self.contexts = [_lib.Context(self.device) for _ in range(max(1, int(workers)))]
        else:
            self.contexts = [None] * max(1, int(workers))
        self._owns_contexts = contexts is None and file_pass is None
        if self._owns_contexts and share_split and len(self.contexts) > 1 and self.group_samples <= 0:
            # a file holds a few hundred events: its split search cannot fill 1036 persistent CTAs, but a full wave
            # would keep every other context's kernels off the SMs until it ends.  Each context takes its share
            # of the wave (at least one CTA per SM), so the searches of the files in flight run side by side.
            for c in self.contexts:
                c.set_split_ctas(max(148, SPLIT_WAVE // len(self.contexts)))
        self.local = None

    def close(self):
        if self._owns_contexts:
            for c in self.contexts:
                c.close()
        self.contexts = []

    # -- this rank's files ------------------------------------------------------------------
    def parse_local(self, traces, timestep, event_detector=lambda_event_parser(threshold=90),
                    segmenter=SpeedyStatSplit(prior_segments_per_second=10, cutoff_freq=2000.),
                    filter_params=(1, 2000)):
        """The owned files through the worker contexts -- several files per pass where the rules allow it
        (`groupable`) -- returns {file index: per-file result}."""
        n_files = len(traces)
        seconds = self._seconds(timestep, n_files)
        mine = assign_files(n_files, self.rank, self.world)
        if self.group_samples > 0 and groupable(event_detector):
            lengths = {i: len(traces[i]) for i in mine}
            items = plan_groups(lengths, seconds, mine, self.group_samples, min_groups=2 * len(self.contexts))
        else:
            items = [[i] for i in mine]
        self.groups = items
        results, errors = {}, []
        lock = threading.Lock()
        cursor = [0]

        def work(ctx):
            while True:
                with lock:
                    if cursor[0] >= len(items) or errors:
                        return
                    item = items[cursor[0]]
                    cursor[0] += 1
                try:
                    if len(item) == 1:
                        rs = [self.file_pass(ctx, traces[item[0]], float(seconds[item[0]]), event_detector,
                                             segmenter, filter_params)]
                    else:
                        rs = self.group_pass(ctx, [traces[i] for i in item], float(seconds[item[0]]),
                                             event_detector, segmenter, filter_params)
                except BaseException as exc:  # surfaced on the calling thread
                    with lock:
                        errors.append(exc)
                    return
                with lock:
                    results.update(zip(item, rs))

        if len(self.contexts) == 1 or len(items) <= 1:
            work(self.contexts[0])
        else:
            threads = [threading.Thread(target=work, args=(c,)) for c in self.contexts[:len(items)]]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        if errors:
            raise errors[0]
        self.local = results
        return results

    @staticmethod
    def _seconds(timestep, n_files):
        ts = np.asarray(timestep, np.float64)
        if ts.ndim == 0:
            ts = np.full(n_files, float(ts))
        if ts.shape[0] != n_files:
            raise ValueError("timestep must be a scalar or hold one value per file")
        return 1000. / ts

    @staticmethod
    def _stack(results, order):
        """Per-file results -> contiguous rows with the file index in front."""
        ei, ef, si, sf = [], [], [], []
        for i in order:
            r = results[i]
            ne, ns = r["ev_int"].shape[0], r["seg_int"].shape[0]
            ei.append(np.concatenate([np.full((ne, 1), i, np.int64), r["ev_int"].astype(np.int64)], axis=1))
            si.append(np.concatenate([np.full((ns, 1), i, np.int64), r["seg_int"].astype(np.int64)], axis=1))
            ef.append(np.asarray(r["ev_flt"], np.float64).reshape(ne, 4))
            sf.append(np.asarray(r["seg_flt"], np.float64).reshape(ns, 4))

        def cat(parts, w, dt):
            return np.concatenate(parts, axis=0) if parts else np.zeros((0, w), dt)
        return cat(ei, 3, np.int64), cat(ef, 4, np.float64), cat(si, 4, np.int64), cat(sf, 4, np.float64)

    # -- the whole batch ----------------------------------------------------------------------
    def parse(self, traces, timestep, event_detector=lambda_event_parser(threshold=90),
              segmenter=SpeedyStatSplit(prior_segments_per_second=10, cutoff_freq=2000.),
              filter_params=(1, 2000)):
        n_files = len(traces)
        seconds = self._seconds(timestep, n_files)
        results = self.parse_local(traces, timestep, event_detector, segmenter, filter_params)
        ei, ef, si, sf = self._stack(results, sorted(results))
        if self.world > 1:
            ei, ef, si, sf = self._gather(ei, ef, si, sf)
        return BatchTables(ei, ef, si, sf, seconds, n_files)

    def _gather(self, ei, ef, si, sf):
        """Every rank's rows on every rank, ordered by file.  Two ragged all-gathers (dist.gather_tables):
        the row counts travel first, in one small collective."""
        import torch
        import torch.distributed as dist
        from .dist import gather_tables
        dev = torch.device("cuda", self.device) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
        mine = torch.tensor([ei.shape[0], si.shape[0]], dtype=torch.int64, device=dev)
        counts = torch.empty(2 * self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, mine, group=self.group)
        counts = counts.cpu().numpy().reshape(self.world, 2)
        out = []
        for ints, flts, col in ((ei, ef, 0), (si, sf, 1)):
            gi, gf = gather_tables(torch.from_numpy(np.ascontiguousarray(ints)).to(dev),
                                   torch.from_numpy(np.ascontiguousarray(flts)).to(dev),
                                   [int(c) for c in counts[:, col]], dist, self.group)
            out += list(_merge_by_file(gi.cpu().numpy(), gf.cpu().numpy()))
        return out[0], out[1], out[2], out[3]


def _merge_by_file(gi, gf):
    """Rank-major gathered rows -> file order.  Ranks hold interleaved files, but the rows of one file are
    contiguous and already ordered, so the merge moves whole per-file blocks (one memcpy each) instead of
    sorting rows: a 1000-file batch has 1000 blocks and possibly 10^7 rows."""
    n = gi.shape[0]
    if n == 0:
        return gi, gf
    col = gi[:, 0]
    starts = np.concatenate(([0], np.flatnonzero(col[1:] != col[:-1]) + 1))
    ends = np.concatenate((starts[1:], [n]))
    order = np.argsort(col[starts], kind="stable")
    if np.array_equal(order, np.arange(order.shape[0])):
        return gi, gf
    oi, of = np.empty_like(gi), np.empty_like(gf)
    pos = 0
    for b in order:
        a, e = int(starts[b]), int(ends[b])
        oi[pos:pos + e - a] = gi[a:e]
        of[pos:pos + e - a] = gf[a:e]
        pos += e - a
    return oi, of
