"""Core value types behind the reference's names: ``Segment``, ``MetaSegment``, ``ignored``.

Scripts written against PyPore/core.py keep working -- same attribute names, the lazy ``mean/std/min/max/n`` of a
``Segment`` (core.py:209-223), the eager statistics of a ``MetaSegment`` (core.py:26-32), ``scale`` / ``to_meta`` /
``delete`` and the JSON methods -- but the design is this repository's:

* statistics never come from NumPy: a segment produced by the device parsers carries the row the segmented
  reduction kernel (csrc/stats.cuh) wrote for it, and a hand-made ``Segment`` asks the device on first access;
* the wire format is data (``wire.FIELDS``), shared by every class, instead of per-class method bodies.
"""
import numpy as np

from . import wire

_SPAN = ("start", "end", "duration")
_FROZEN = wire.FIELDS["MetaSegment"]   # what survives Segment.to_meta()


class ignored(object):
    """``with ignored(KeyError, AttributeError): ...`` swallows the listed exceptions (the reference exports this
    helper from core.py:251-261 and its users import it from here)."""

    def __init__(self, *exceptions):
        self.exceptions = exceptions

    def __enter__(self):
        return self

    def __exit__(self, kind, value, traceback):
        return kind is not None and issubclass(kind, self.exceptions)


def device_stats(current):
    """(mean, std, min, max) of one array through the K4 reduction -- population std, extrema on the samples,
    NaN for an empty array.  There is no host implementation of these."""
    from . import _lib
    samples = np.ascontiguousarray(current, dtype=np.float64)
    if samples.shape[0] == 0:
        return (float("nan"),) * 4
    ctx = _lib.default_context()
    ctx.upload_events_f64([samples])
    row = ctx.event_stats(1)
    return tuple(row[k][0] for k in ("mean", "std", "min", "max"))


_gpu_stats = device_stats  # older name, used by tests


def _adopt(obj, attrs, keep_existing):
    """Keyword arguments become attributes.  Read-only names (the statistics properties) are skipped silently,
    like the reference; `keep_existing` also skips anything the object already has (core.py:130-134)."""
    for name, value in attrs.items():
        if keep_existing and (name in obj.__dict__ or hasattr(type(obj), name)):
            continue   # (asked of the class, so that the lazy statistics are not computed just to be skipped)
        try:
            setattr(obj, name, value)
        except AttributeError:
            pass


def _complete_span(obj):
    """Two of start / end / duration determine the third (core.py:34-40)."""
    known = [k for k in _SPAN if hasattr(obj, k)]
    if len(known) != 2:
        return
    if "duration" not in known:
        obj.duration = obj.end - obj.start
    elif "end" not in known:
        obj.end = obj.start + obj.duration
    else:
        obj.start = obj.end - obj.duration


class _Stretch(object):
    """What every stretch of current has in common: the JSON face."""

    def to_dict(self):
        return wire.record(self)

    def to_json(self, filename=None):
        return wire.dumps(self.to_dict(), filename)

    __repr__ = lambda self: self.to_json()  # noqa: E731  (the representation IS the JSON, subclass overrides included)
    __len__ = lambda self: self.n           # noqa: E731


class MetaSegment(_Stretch):
    """A stretch of current reduced to its metadata (core.py:14-113): no samples, plain attributes."""

    def __init__(self, **kwargs):
        _adopt(self, kwargs, keep_existing=False)
        if "current" in self.__dict__:
            samples = self.__dict__.pop("current")
            self.n = len(samples)
            self.mean, self.std, self.min, self.max = device_stats(samples)
        _complete_span(self)

    def to_meta(self):
        """Already metadata."""

    def delete(self):
        """Nothing to release."""

    @classmethod
    def from_json(cls, filename=None, json=None):
        assert filename or json and not (filename and json)
        text = wire.read_text(filename) if filename else json
        return MetaSegment(**wire.loads_flat(text))


class Segment(_Stretch):
    """A stretch of current with its samples (core.py:115-249).  ``mean``, ``std``, ``min``, ``max`` are
    read-only and lazy; ``n`` is the sample count."""

    def __init__(self, current, **kwargs):
        self.current = current
        _adopt(self, kwargs, keep_existing=True)

    # -- statistics ------------------------------------------------------------------------------
    def _set_stats(self, mean, std, mn, mx):
        """The parsers attach the row the device already computed for this segment."""
        self.__dict__["_stats"] = (mean, std, mn, mx)
        return self

    def _get_stats(self):
        row = self.__dict__.get("_stats")
        if row is None:
            row = self.__dict__["_stats"] = device_stats(self.current)
        return row

    def __setattr__(self, name, value):
        if name == "current":
            self.__dict__.pop("_stats", None)   # other samples, other statistics
        object.__setattr__(self, name, value)

    mean = property(lambda self: self._get_stats()[0])
    std = property(lambda self: self._get_stats()[1])
    min = property(lambda self: self._get_stats()[2])
    max = property(lambda self: self._get_stats()[3])
    n = property(lambda self: len(self.current))

    # -- life cycle ------------------------------------------------------------------------------
    def _freeze(self, names, meta_class):
        """Turn into `meta_class` in place: the listed values become plain attributes, the samples go."""
        state = self.__dict__
        for name in names:
            try:
                state[name] = getattr(self, name)
            except (AttributeError, KeyError):
                pass
        state.pop("_stats", None)
        state.pop("current", None)
        self.__class__ = meta_class

    def to_meta(self):
        self._freeze(_FROZEN, MetaSegment)

    def delete(self):
        self.__dict__.pop("current", None)

    def scale(self, sampling_freq):
        """start / end / duration from samples to seconds (core.py:199-207); stops at the first one missing."""
        for name in _SPAN:
            if not hasattr(self, name):
                return
            setattr(self, name, getattr(self, name) / sampling_freq)

    @classmethod
    def from_json(cls, filename=None, json=None):
        assert filename or json and not (filename and json)
        text = wire.read_text(filename) if filename else json
        if "current" not in text:
            return MetaSegment.from_json(json=text)
        attrs = wire.loads_flat(text, bracketed_spaces=True)
        samples = np.array([float(v) for v in attrs.pop("current")[1:-1].split()])
        return Segment(samples, **attrs)
