"""Container drivers: host-side mirror of the hot-path part of PyPore/DataTypes.py.

``File.parse(parser=...)``, ``Event.filter(order, cutoff)`` and
``Event.parse(parser=...)`` keep the reference's signatures and side effects
(DataTypes.py:258-289, 589-602).  On top of that, ``File.parse`` accepts
``segmenter=`` / ``filter_params=`` to run threshold -> [filter] -> split ->
statistics in one device-resident pass, returning the same objects lazily built
from the compact tables.  HMM-guided parsing, plotting, MySQL and .abf reading
are out of scope.
"""
import json
from functools import lru_cache, reduce

import numpy as np

from . import _lib
from .core import MetaSegment, Segment, ignored, _jsonable
from .parsers import RuleSet, SpeedyStatSplit, lambda_event_parser, parser, _as_float32_trace


class MetaEvent(MetaSegment):
    """Metadata of an event (DataTypes.py:49-80)."""

    def __init__(self, **kwargs):
        MetaSegment.__init__(self, **kwargs)

    def delete(self):
        with ignored(AttributeError):
            del self.state_parser
        for segment in self.segments:
            segment.delete()
        del self

    def to_dict(self):
        """DataTypes.py:196-201 (no 'filtered' key, unlike Event.to_dict)."""
        keys = ['mean', 'std', 'min', 'max', 'start', 'end', 'duration',
                'filter_order', 'filter_cutoff', 'n', 'state_parser', 'segments']
        d = {i: _jsonable(getattr(self, i)) for i in keys if hasattr(self, i)}
        d['name'] = self.__class__.__name__
        return d

    def to_json(self, filename=None):
        """DataTypes.py:203-216."""
        d = self.to_dict()
        with ignored(KeyError, AttributeError):
            d['segments'] = [seg.to_dict() for seg in d['segments']]
        with ignored(KeyError, AttributeError):
            d['state_parser'] = d['state_parser'].to_dict()
        _json = json.dumps(d, indent=4, separators=(',', ' : '))
        if filename:
            with open(filename, 'w') as out:
                out.write(_json)
        return _json

    @classmethod
    def from_json(cls, _json):
        """DataTypes.py:218-225."""
        if _json.endswith(".json"):
            with open(_json, 'r') as infile:
                _json = ''.join(line for line in infile)
        return cls(**json.loads(_json))

    @classmethod
    def from_segments(cls, segments):
        return cls(segments=segments)

    @property
    def n(self):
        try:
            return len(self.segments)
        except Exception:
            return 0


def bessel_coefficients(order, cutoff, second):
    """(b, a, zi) for Event.filter: scipy.signal.bessel exactly as DataTypes.py:268-270
    requests it, plus the lfilter_zi steady state filtfilt starts from.  The design costs ~0.3 ms of
    Python per call and every event of every file of a batch asks for the same one: memoised (read-only arrays)."""
    return _bessel_coefficients(order, float(cutoff), float(second))


@lru_cache(maxsize=64)
def _bessel_coefficients(order, cutoff, second):
    from scipy import signal
    nyquist = second / 2.
    b, a = signal.bessel(order, cutoff / nyquist, btype='low', analog=0, output='ba')
    b = np.atleast_1d(np.asarray(b, np.float64))
    a = np.atleast_1d(np.asarray(a, np.float64))
    n = max(len(a), len(b))
    b = np.r_[b, np.zeros(n - len(b))] / a[0]
    a = np.r_[a, np.zeros(n - len(a))] / a[0]
    # lfilter_zi: solve (I - A^T) zi = b[1:] - a[1:] b[0] for the companion matrix A of a
    comp = np.zeros((n - 1, n - 1))
    comp[0, :] = -a[1:]
    for i in range(1, n - 1):
        comp[i, i - 1] = 1.0
    zi = np.linalg.solve(np.eye(n - 1) - comp.T, b[1:] - a[1:] * b[0])
    for v in (b, a, zi):
        v.setflags(write=False)
    return b, a, zi


class Event(Segment):
    """An event: a stretch of the file holding useful data (DataTypes.py:241-565)."""

    def __init__(self, current, segments=[], **kwargs):
        if len(segments) > 0:
            try:
                current = np.concatenate([seg.current for seg in segments])
            except Exception:
                current = []
        Segment.__init__(self, current, filtered=False, segments=segments, **kwargs)

    def filter(self, order=1, cutoff=2000.):
        """Zero-phase Bessel low-pass of ``self.current`` (DataTypes.py:258-274)."""
        if type(self) != Event:
            raise TypeError("Cannot filter a metaevent. Must have the current.")
        b, a, zi = bessel_coefficients(order, cutoff, self.second)
        cur = np.ascontiguousarray(self.current, dtype=np.float64)
        padlen = 3 * len(b)
        if cur.shape[0] <= padlen:
            raise ValueError("The length of the input vector x must be greater than padlen, which is %d."
                             % padlen)
        ctx = _lib.default_context()
        ctx.upload_events_f64([cur])
        ctx.filter_events(b, a, zi)
        self.current = ctx.event_samples(cur.shape[0])
        self.filtered = True
        self.filter_order = order
        self.filter_cutoff = cutoff

    def parse(self, parser=SpeedyStatSplit(prior_segments_per_second=10), hmm=None):
        """Segment the event with a plug-in state parser (DataTypes.py:276-289,333)."""
        if hmm:
            raise NotImplementedError("HMM-guided segmentation needs yahmm and is out of scope")
        self.segments = parser.parse(self.current)
        for segment in self.segments:
            segment.event = self
            segment.scale(float(self.file.second))
        self.state_parser = parser

    def delete(self):
        with ignored(AttributeError):
            del self.current
        with ignored(AttributeError):
            del self.state_parser
        for segment in self.segments:
            segment.delete()
        del self

    def to_meta(self):
        for prop in ['mean', 'std', 'duration', 'start', 'min', 'max', 'end', 'start']:
            with ignored(AttributeError, KeyError):
                self.__dict__[prop] = getattr(self, prop)
        self.__dict__.pop('_stats', None)
        with ignored(AttributeError):
            del self.current
        for segment in self.segments:
            segment.to_meta()
        self.__class__ = type("MetaEvent", (MetaEvent,), self.__dict__)

    def to_dict(self):
        """DataTypes.py:493-498."""
        keys = ['mean', 'std', 'min', 'max', 'start', 'end', 'duration', 'filtered',
                'filter_order', 'filter_cutoff', 'n', 'state_parser', 'segments']
        d = {i: _jsonable(getattr(self, i)) for i in keys if hasattr(self, i)}
        d['name'] = self.__class__.__name__
        return d

    def to_json(self, filename=None):
        """DataTypes.py:500-513."""
        d = self.to_dict()
        with ignored(KeyError, AttributeError):
            d['segments'] = [seg.to_dict() for seg in d['segments']]
        with ignored(KeyError, AttributeError):
            d['state_parser'] = d['state_parser'].to_dict()
        _json = json.dumps(d, indent=4, separators=(',', ' : '))
        if filename:
            with open(filename, 'w') as out:
                out.write(_json)
        return _json

    @classmethod
    def from_json(cls, _json):
        """DataTypes.py:515-529: a JSON without 'current' gives a MetaEvent."""
        if _json.endswith(".json"):
            with open(_json, 'r') as infile:
                _json = ''.join(line for line in infile)
        d = json.loads(_json)
        event = MetaSegment()
        if 'current' not in d.keys():
            event.__class__ = type("MetaEvent", (MetaEvent,), d)
        else:
            event = cls(d['current'], start=d['start'])
        return event

    @property
    def n(self):
        return len(self.segments)


class _LazyList(object):
    """List-like whose items are built on first access (object creation is the
    dominant host cost at 10^5..10^6 segments)."""

    def __init__(self, n, factory):
        self._n, self._factory, self._cache = int(n), factory, {}

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        item = self._cache.get(i)
        if item is None:
            item = self._cache[i] = self._factory(i)
        return item

    def __iter__(self):
        for i in range(self._n):
            yield self[i]


class File(Segment):
    """A trace and the events detected in it (DataTypes.py:567-602).

    Only the ``current=`` / ``timestep=`` constructor is supported: there is no
    .abf data offline and the ABF reader is out of scope.
    """

    def __init__(self, filename=None, current=None, timestep=None, **kwargs):
        if current is not None and timestep is not None:
            filename = ""
        elif filename and current is None and timestep is None:
            raise NotImplementedError("reading .abf files is out of scope; pass current= and timestep=")
        else:
            raise SyntaxError("Must provide current and timestep, or filename "
                              "corresponding to a valid abf file.")
        Segment.__init__(self, current=current, filename=filename, second=1000. / timestep,
                         events=[], sample=None)

    def __getitem__(self, index):
        return self.events[index]

    @property
    def n(self):
        return len(self.events)

    def parse(self, parser=lambda_event_parser(threshold=90), segmenter=None, filter_params=None,
              context=None):
        """Detect events with a plug-in event parser (DataTypes.py:589-602).

        With ``segmenter=`` (a ``SpeedyStatSplit``) the whole pipeline runs on the
        resident trace -- threshold scan, optional ``Event.filter(*filter_params)``,
        split search, statistics -- and every event comes back with its
        ``segments`` already parsed, as after ``event.parse(parser=segmenter)``.
        """
        if segmenter is None and filter_params is None and not isinstance(parser, lambda_event_parser):
            # any duck-typed parser, exactly like the reference
            self.events = [Event(current=seg.current, start=seg.start / self.second,
                                 end=(seg.start + seg.duration) / self.second,
                                 duration=seg.duration / self.second, second=self.second, file=self)
                           for seg in parser.parse(self.current)]
            self.event_parser = parser
            return
        if not isinstance(parser, lambda_event_parser):
            raise TypeError("the device-resident pipeline needs a pypore_b200 lambda_event_parser")
        ctx = context or _lib.default_context()
        host = np.asarray(self.current)
        self._parse_resident(ctx, host, _as_float32_trace(host), parser, segmenter, filter_params)

    # ------------------------------------------------------------------
    def _parse_resident(self, ctx, host, x32, parser, segmenter, filter_params):
        second = self.second
        filt = None
        if filter_params is not None:
            order, cutoff = filter_params
            filt = bessel_coefficients(order, cutoff, second)
        rs = parser._device_rules()
        tables = None
        if segmenter is not None:
            mw, MW, W, gain = segmenter._params()
        if rs is not None and segmenter is not None:
            # one call from host memory: the copy is chunked and overlapped with the stages (pp_pipeline_host)
            counts = ctx.pipeline(parser.threshold, min_width=mw, max_width=MW, window_width=W, min_gain=gain,
                                  filter_ba=filt, with_stats=True, host_trace=x32, **rs.device_args())
            ev_start, ev_len = ctx.events(counts["events"])
            tables = ctx.segments(counts["segments"])
            r_start, _, r_min, r_max, _ = ctx.runs(counts["runs"])
            idx = np.searchsorted(r_start, ev_start)
            ev_min, ev_max = r_min[idx], r_max[idx]
        else:
            ctx.upload_trace(x32)
            ev_start, ev_len, ev_min, ev_max = parser._detect(ctx, host)
            if filt is not None and len(ev_start):
                ctx.filter_events(*filt)
            if segmenter is not None and len(ev_start):
                n_seg = ctx.statsplit(mw, MW, W, gain)
                ctx.segment_stats()
                tables = ctx.segments(n_seg)
        filtered = None
        if filt is not None and len(ev_start):
            filtered = ctx.event_samples(int(ev_len.sum()))
        self.event_table = dict(start=ev_start, length=ev_len, min=ev_min, max=ev_max)
        self.segment_table = tables
        ev_off = np.concatenate(([0], np.cumsum(ev_len)))
        seg_bounds = None
        if tables is not None:
            seg_bounds = np.searchsorted(tables["event"], np.arange(len(ev_start) + 1))

        def make_event(i):
            s, n = int(ev_start[i]), int(ev_len[i])
            if filtered is not None:
                cur = filtered[ev_off[i]:ev_off[i + 1]]
            else:
                cur = np.array(host[s:s + n])
            ev = Event(current=cur, start=s / second, end=(s + n) / second, duration=n / second,
                       second=second, file=self)
            if filtered is not None:
                ev.filtered = True
                ev.filter_order, ev.filter_cutoff = filter_params
            if tables is not None:
                lo, hi = int(seg_bounds[i]), int(seg_bounds[i + 1])

                def make_segment(k, cur=cur, ev=ev, lo=lo):
                    k += lo
                    a, b = int(tables["start"][k]), int(tables["end"][k])
                    seg = Segment(current=cur[a:b], start=a, duration=(b - a), end=b)
                    seg._set_stats(tables["mean"][k], tables["std"][k], tables["min"][k], tables["max"][k])
                    seg.event = ev
                    seg.scale(float(second))
                    return seg
                ev.segments = _LazyList(hi - lo, make_segment)
                ev.state_parser = segmenter
            return ev

        self.events = _LazyList(len(ev_start), make_event)
        self.event_parser = parser

    def to_meta(self):
        """Drop the ionic current of the file and of everything under it (DataTypes.py:683-693)."""
        with ignored(AttributeError):
            del self.current
        events = list(self.events)  # lazily built events are materialised once, then reduced to metadata
        for event in events:
            event.to_meta()
        self.events = events

    def to_dict(self):
        """DataTypes.py:695-706."""
        keys = ['filename', 'n', 'event_parser', 'mean', 'std', 'duration', 'start', 'end', 'events']
        if not hasattr(self, 'end') and (hasattr(self, 'start') and hasattr(self, 'duration')):
            setattr(self, 'end', self.start + self.duration)
        d = {i: _jsonable(getattr(self, i)) for i in keys if hasattr(self, i)}
        d['name'] = self.__class__.__name__
        return d

    def to_json(self, filename=None):
        """The file, its events and their segments as the reference's JSON (DataTypes.py:708-738)."""
        d = self.to_dict()
        devents = []
        for event in d['events']:
            devent = event.to_dict()
            try:
                devent['segments'] = [state.to_dict() for state in devent['segments']]
                devent['state_parser'] = devent['state_parser'].to_dict()
            except Exception:
                with ignored(KeyError, AttributeError):
                    del devent['segments']
                    del devent['state_parser']
            devents.append(devent)
        d['events'] = devents
        d['event_parser'] = d['event_parser'].to_dict()
        _json = json.dumps(d, indent=4, separators=(',', ' : '))
        if filename:
            with open(filename, 'w') as outfile:
                outfile.write(_json)
        return _json

    @classmethod
    def from_json(cls, _json):
        """Rebuild a file and its events from to_json's output (DataTypes.py:740-796).  Without the .abf
        file -- always the case here, reading .abf is out of scope -- the result holds MetaEvents /
        MetaSegments, exactly the reference's own fallback."""
        if _json.endswith(".json"):
            with open(_json, 'r') as infile:
                _json = ''.join(line for line in infile)
        d = json.loads(_json)
        if d['name'] != "File":
            raise TypeError("JSON does not encode a file")
        try:
            file = File(filename=d['filename'] + ".abf")
            meta = False
        except Exception:
            file = File(current=[], timestep=1)
            meta = True
        file.event_parser = parser.from_json(json.dumps(d['event_parser']))
        file.events = []
        for _json in d['events']:
            s, e = int(_json['start'] * file.second), int(_json['end'] * file.second)
            if meta:
                event = MetaEvent(**_json)
            else:
                current = file.current[s:e]
                event = Event(current=current, start=s / file.second, end=e / file.second,
                              duration=(e - s) / file.second, second=file.second, file=file)
            if _json['filtered']:
                if not meta:
                    event.filter(order=_json['filter_order'], cutoff=_json['filter_cutoff'])
            if meta:
                event.segments = [MetaSegment(**s_json) for s_json in _json['segments']]
            else:
                event.segments = [Segment(current=event.current[int(s_json['start'] * file.second):
                                                                int(s_json['end'] * file.second)],
                                          second=file.second, event=event, **s_json)
                                  for s_json in _json['segments']]
            event.state_parser = parser.from_json(json.dumps(_json['state_parser']))
            event.filtered = _json['filtered']
            file.events.append(event)
        return file

    def delete(self):
        with ignored(AttributeError):
            del self.current
        with ignored(AttributeError):
            del self.event_parser
        for event in self.events:
            event.delete()
        del self

    def close(self):
        self.delete()


class Experiment(object):
    """A series of files analysed together (DataTypes.py:938-1031).

    ``filenames`` may hold ``File`` objects (``File(current=..., timestep=...)``: the only way to get data in
    here, reading .abf is out of scope) next to file names; a name is handed to ``File(name)`` like the
    reference does.  ``parse`` keeps the reference's signature, defaults, printed messages and results; each
    file goes through ONE device-resident pass (threshold scan -> Event.filter -> segmenter -> statistics)
    instead of the reference's per-event Python loop.
    """

    def __init__(self, filenames, name=None):
        self.filenames = filenames
        self.name = name or "Experiment"
        self.files = []

    def parse(self, event_detector=lambda_event_parser(threshold=90),
              segmenter=SpeedyStatSplit(prior_segments_per_second=10, cutoff_freq=2000.),
              filter_params=(1, 2000), verbose=True, meta=False, batch=None):
        """DataTypes.py:956-988.  ``batch=`` (a ``pypore_b200.batch.FileBatch``; needs ``meta=True``, which
        discards the samples anyway) runs the files several at a time on the batch's worker contexts --
        and over its ranks, files being the unit -- and builds the same metadata objects from the gathered
        tables (kept as ``self.tables``)."""
        if batch is not None:
            return self._parse_batch(batch, event_detector, segmenter, filter_params, verbose, meta)
        for file in (f if isinstance(f, File) else File(f) for f in self.filenames):
            if verbose:
                print("Opening {}".format(file.filename))
            batched = (isinstance(event_detector, lambda_event_parser) and event_detector._device_rules() is not None
                       and (segmenter is None or isinstance(segmenter, SpeedyStatSplit)))
            if batched and (segmenter is not None or filter_params is not None):
                file.parse(parser=event_detector, segmenter=segmenter, filter_params=filter_params)
            else:
                # arbitrary Python rules or a foreign segmenter: the reference's own loop (DataTypes.py:972-982)
                file.parse(parser=event_detector)
                for event in file.events:
                    if filter_params is not None:
                        event.filter(*filter_params)
                    if segmenter is not None:
                        event.parse(parser=segmenter)
            if verbose:
                print("\tDetected {} Events".format(file.n))
                if segmenter is not None:
                    for i, event in enumerate(file.events):
                        print("\t\tEvent {} has {} segments".format(i + 1, event.n))
            if meta:
                file.to_meta()
            self.files.append(file)

    def _parse_batch(self, batch, event_detector, segmenter, filter_params, verbose, meta):
        if not meta:
            raise ValueError("batch= keeps only the event / segment tables on the host; use it with meta=True")
        if segmenter is None:
            raise ValueError("batch= runs the whole pipeline; it needs a segmenter")
        if not (isinstance(event_detector, lambda_event_parser) and event_detector._device_rules() is not None
                and isinstance(segmenter, SpeedyStatSplit)):
            raise TypeError("batch= needs a lambda_event_parser with device-evaluable rules (parsers.RuleSet or the "
                            "defaults) and a SpeedyStatSplit; other plug-ins go through the sequential loop")
        files = [f if isinstance(f, File) else File(f) for f in self.filenames]
        tables = batch.parse([f.current for f in files], [1000. / f.second for f in files], event_detector,
                             segmenter, filter_params)
        self.tables = tables
        for file in tables.files([f.filename for f in files], event_detector, segmenter, filter_params):
            if verbose:
                print("Opening {}".format(file.filename))
                print("\tDetected {} Events".format(file.n))
                for i, event in enumerate(file.events):
                    print("\t\tEvent {} has {} segments".format(i + 1, event.n))
            self.files.append(file)

    def delete(self):
        with ignored(AttributeError):
            del self.events
        with ignored(AttributeError):
            del self.segments
        for file in self.files:
            file.delete()
        del self

    @property
    def n(self):
        return len(self.files)

    @property
    def events(self):
        """All the events in all files."""
        try:
            return reduce(list.__add__, [list(file.events) for file in self.files])
        except Exception:
            return []

    @property
    def segments(self):
        """All segments of all events."""
        try:
            return reduce(list.__add__, [list(event.segments) for event in self.events])
        except Exception:
            return []


class Sample(object):
    """A container for events all suggested to be from the same substrate (DataTypes.py:1034-1040)."""

    def __init__(self, events=[], files=[], label=None):
        self.events = events
        self.files = files
        self.label = label
