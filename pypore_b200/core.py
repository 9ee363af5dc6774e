"""Core value types: host-side mirror of the reference's PyPore/core.py.

``Segment`` / ``MetaSegment`` keep the reference's attribute names and lazy
``mean/std/min/max/n`` properties (core.py:14-40, 115-223) so scripts written
against PyPore keep working.  The difference is where the statistics come
from: segments produced by the GPU parsers carry the values the segmented
reduction kernel computed (csrc/stats.cuh); a bare user-made ``Segment`` asks the
GPU for them on first access.  There is no NumPy fallback for the statistics.
"""
import json
import re
from contextlib import contextmanager

import numpy as np


@contextmanager
def ignored(*exceptions):
    """``try: ... except exceptions: pass`` as a context manager (core.py:251-261)."""
    try:
        yield
    except exceptions:
        pass


def _jsonable(v):
    if isinstance(v, np.generic):
        return v.item()
    return v


class MetaSegment(object):
    """Metadata of a stretch of ionic current without the samples (core.py:14-113)."""

    def __init__(self, **kwargs):
        for key, value in kwargs.items():
            with ignored(AttributeError):
                setattr(self, key, value)
        # If current is passed in, take the statistics from it and drop the array
        # (core.py:26-32); the numbers come from the GPU reduction.
        if hasattr(self, "current"):
            cur = self.current
            st = _gpu_stats(cur)
            self.n = len(cur)
            self.mean, self.std, self.min, self.max = st
            del self.current
        if hasattr(self, "start") and hasattr(self, "end") and not hasattr(self, "duration"):
            self.duration = self.end - self.start
        elif hasattr(self, "start") and hasattr(self, "duration") and not hasattr(self, "end"):
            self.end = self.start + self.duration
        elif hasattr(self, "end") and hasattr(self, "duration") and not hasattr(self, "start"):
            self.start = self.end - self.duration

    def __repr__(self):
        return self.to_json()

    def __len__(self):
        return self.n

    def delete(self):
        del self

    def to_meta(self):
        pass

    def to_dict(self):
        keys = ['mean', 'std', 'min', 'max', 'start', 'end', 'duration']
        d = {i: _jsonable(getattr(self, i)) for i in keys if hasattr(self, i)}
        d['name'] = self.__class__.__name__
        return d

    def to_json(self, filename=None):
        _json = json.dumps(self.to_dict(), indent=4, separators=(',', ' : '))
        if filename:
            with open(filename, 'w') as outfile:
                outfile.write(_json)
        return _json

    @classmethod
    def from_json(cls, filename=None, json=None):
        assert filename or json and not (filename and json)
        if filename:
            with open(filename, 'r') as infile:
                json = ''.join([line for line in infile])
        words = re.findall(r"\[[\w'.-]+\]|[\w'.-]+", json)
        attrs = {words[i]: words[i + 1] for i in range(0, len(words), 2)}
        return MetaSegment(**attrs)


def _gpu_stats(current):
    """(mean, std, min, max) of one array through the K4 reduction kernel."""
    from . import _lib
    cur = np.ascontiguousarray(current, dtype=np.float64)
    if cur.shape[0] == 0:
        nan = float('nan')
        return nan, nan, nan, nan
    ctx = _lib.default_context()
    ctx.upload_events_f64([cur])
    st = ctx.event_stats(1)
    return st["mean"][0], st["std"][0], st["min"][0], st["max"][0]


class Segment(object):
    """A stretch of ionic current with lazily evaluated statistics (core.py:115-249)."""

    def __init__(self, current, **kwargs):
        self.current = current
        for key, value in kwargs.items():
            if hasattr(self, key):
                continue
            with ignored(AttributeError):
                setattr(self, key, value)

    # statistics the GPU already computed for this segment (set by the parsers)
    def _set_stats(self, mean, std, mn, mx):
        self.__dict__['_stats'] = (mean, std, mn, mx)
        return self

    def _get_stats(self):
        st = self.__dict__.get('_stats')
        if st is None:
            st = _gpu_stats(self.current)
            self.__dict__['_stats'] = st
        return st

    def __setattr__(self, key, value):
        if key == 'current':
            self.__dict__.pop('_stats', None)  # new samples invalidate cached statistics
        object.__setattr__(self, key, value)

    def __repr__(self):
        return self.to_json()

    def __len__(self):
        return self.n

    def to_dict(self):
        keys = ['mean', 'std', 'min', 'max', 'start', 'end', 'duration']
        d = {i: _jsonable(getattr(self, i)) for i in keys if hasattr(self, i)}
        d['name'] = self.__class__.__name__
        return d

    def to_json(self, filename=None):
        _json = json.dumps(self.to_dict(), indent=4, separators=(',', ' : '))
        if filename:
            with open(filename, 'w') as outfile:
                outfile.write(_json)
        return _json

    def to_meta(self):
        for key in ['mean', 'std', 'min', 'max', 'end', 'start', 'duration']:
            with ignored(KeyError, AttributeError):
                self.__dict__[key] = getattr(self, key)
        self.__dict__.pop('_stats', None)
        del self.current
        self.__class__ = type("MetaSegment", (MetaSegment,), self.__dict__)

    def delete(self):
        with ignored(AttributeError):
            del self.current
        del self

    def scale(self, sampling_freq):
        """Rescale start/end/duration from samples to seconds (core.py:199-207)."""
        with ignored(AttributeError):
            self.start /= sampling_freq
            self.end /= sampling_freq
            self.duration /= sampling_freq

    @property
    def mean(self):
        return self._get_stats()[0]

    @property
    def std(self):
        return self._get_stats()[1]

    @property
    def min(self):
        return self._get_stats()[2]

    @property
    def max(self):
        return self._get_stats()[3]

    @property
    def n(self):
        return len(self.current)

    @classmethod
    def from_json(cls, filename=None, json=None):
        assert filename or json and not (filename and json)
        if filename:
            with open(filename, 'r') as infile:
                json = ''.join([line for line in infile])
        if 'current' not in json:
            return MetaSegment.from_json(json=json)
        words = re.findall(r"\[[\w\s'.-]+\]|[\w'.-]+", json)
        attrs = {words[i]: words[i + 1] for i in range(0, len(words), 2)}
        current = np.array([float(x) for x in attrs['current'][1:-1].split()])
        del attrs['current']
        return Segment(current, **attrs)
