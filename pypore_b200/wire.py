"""The reference's JSON wire format of the result tables, schema-driven and table-first.

PyPore stores a parsed file as nested JSON: File -> events -> segments, every level a dict of a fixed set of
attributes plus ``name`` (the class).  The reference produces it by walking Python objects
(core.py:72-113,152-173; DataTypes.py:196-225,493-529,695-738).  Here the format is DATA -- ``FIELDS`` below --
and there are two producers behind the same ``to_dict`` / ``to_json`` methods:

* ``record`` / ``expand``: any object of the mirror classes (or a user's subclass), attribute by attribute;
* ``file_tree``: a ``File`` that was parsed on the device is serialised straight from its event and segment
  TABLES (column ``tolist()`` calls, no ``Event`` / ``Segment`` object is ever built) -- at 10^5..10^6 segments the
  per-object Python cost is larger than the whole device pipeline.

``loads_*`` are the readers.  Only key names, their order and the text layout (``indent=4``, ``' : '``) are
the reference's; nothing under oracle/ or /root/reference is used at run time.
"""
import json
import re

import numpy as np

_STAT = ("mean", "std", "min", "max")
_SPAN = ("start", "end", "duration")

#: attributes written per class, in order; ``name`` (the class name) always follows.
FIELDS = {
    "Segment": _STAT + _SPAN,
    "MetaSegment": _STAT + _SPAN,
    "Event": _STAT + _SPAN + ("filtered", "filter_order", "filter_cutoff", "n", "state_parser", "segments"),
    "MetaEvent": _STAT + _SPAN + ("filter_order", "filter_cutoff", "n", "state_parser", "segments"),
    "File": ("filename", "n", "event_parser", "mean", "std", "duration", "start", "end", "events"),
}

LAYOUT = dict(indent=4, separators=(",", " : "))


def kind_of(obj):
    """The schema an object is written with: the first class of its MRO that has one."""
    for klass in type(obj).__mro__:
        if klass.__name__ in FIELDS:
            return klass.__name__
    raise TypeError("%s has no wire format" % type(obj).__name__)


def plain(value):
    """numpy scalars are not JSON-serialisable; everything else passes through."""
    return value.item() if isinstance(value, np.generic) else value


def record(obj, kind=None):
    """One level of the format: the schema's attributes that `obj` has, then its class name."""
    missing = object()
    out = {}
    for field in FIELDS[kind or kind_of(obj)]:
        value = getattr(obj, field, missing)
        if value is not missing:
            out[field] = plain(value)
    out["name"] = type(obj).__name__
    return out


def _parser_dict(p):
    return p.to_dict()


def expand_event(ev, strict):
    """An event with its nested levels as dicts.  strict=True is File.to_json's rule: if either the segments or
    the state parser cannot be written, both keys are dropped; strict=False is Event.to_json's: each on its own."""
    d = record(ev)
    if strict:
        try:
            d["segments"], d["state_parser"] = [record(s) for s in d["segments"]], _parser_dict(d["state_parser"])
        except Exception:
            d.pop("segments", None)
            d.pop("state_parser", None)
        return d
    for key, convert in (("segments", lambda v: [record(s) for s in v]), ("state_parser", _parser_dict)):
        try:
            d[key] = convert(d[key])
        except (KeyError, AttributeError):
            pass
    return d


_SCALAR_WORDS = {"nan": "NaN", "inf": "Infinity", "-inf": "-Infinity"}
_MARK = "\ue000%d\ue000"                               # stands for a list while the skeleton is encoded
_MARK_RE = re.compile(r'"\\ue000(\d+)\\ue000"')      # ... and how it looks in the encoder's (ASCII) output
_STEP = " " * LAYOUT["indent"]


class _NotFlat(Exception):
    pass


def _scalar_text(v, strings):
    t = type(v)
    if t is float:
        s = float.__repr__(v)
        return _SCALAR_WORDS.get(s, s)
    if t is int:
        return int.__repr__(v)
    if t is str:
        s = strings.get(v)
        if s is None:
            s = strings[v] = json.dumps(v)
        return s
    if v is True:
        return "true"
    if v is False:
        return "false"
    if v is None:
        return "null"
    raise _NotFlat()


def _column_text(col, strings):
    """JSON text of every value of one column (a tuple of scalars), whole-column calls where the types allow."""
    kinds = set(map(type, col))
    if kinds == {float}:
        out = list(map(float.__repr__, col))
        if "nan" in out or "inf" in out or "-inf" in out:
            out = [_SCALAR_WORDS.get(s, s) for s in out]
        return out
    if kinds == {int}:
        return list(map(int.__repr__, col))
    if kinds == {str}:
        for v in set(col):
            if v not in strings:
                strings[v] = json.dumps(v)
        return [strings[v] for v in col]
    return [_scalar_text(v, strings) for v in col]


def _flat_rows_text(rows, level, strings):
    """A non-empty list of dicts of scalars, all with the same keys, at indent level `level`: the text json.dumps
    would produce for it, made column by column and with one template per list instead of the pure-Python
    pretty-printer (which is what `indent=` costs: the C encoder only writes compact JSON)."""
    keys = tuple(rows[0])
    values = []
    for d in rows:
        if type(d) is not dict or tuple(d) != keys:
            raise _NotFlat()
        values.append(tuple(d.values()))
    inner, item = _STEP * (level + 2), _STEP * (level + 1)
    row_template = "{\n" + ",\n".join(inner + json.dumps(k).replace("%", "%%") + " : %s" for k in keys) + "\n" + item + "}"
    columns = [_column_text(c, strings) for c in zip(*values)]
    return "[\n" + item + (",\n" + item).join(map(row_template.__mod__, zip(*columns))) + "\n" + _STEP * level + "]"


def _skeleton(node, level, texts, strings):
    """Copy of `node` with every list of flat dicts replaced by a marker; their texts go to `texts`."""
    if type(node) is dict:
        return {k: _skeleton(v, level + 1, texts, strings) for k, v in node.items()}
    if type(node) is list and node:
        if type(node[0]) is dict and node[0]:
            try:
                text = _flat_rows_text(node, level, strings)
            except _NotFlat:
                pass
            else:
                texts.append(text)
                return _MARK % (len(texts) - 1)
        return [_skeleton(v, level + 1, texts, strings) for v in node]
    return node


def pretty(tree):
    """json.dumps(tree, **LAYOUT), byte for byte, but the long homogeneous lists (segment rows: 10^5..10^6 dicts of
    eight numbers) are written from a template.  Anything unusual falls back to json.dumps itself."""
    try:
        texts = []
        skeleton = _skeleton(tree, 0, texts, {})
        if not texts:
            return json.dumps(tree, **LAYOUT)
        return _MARK_RE.sub(lambda m: texts[int(m.group(1))], json.dumps(skeleton, **LAYOUT))
    except (_NotFlat, TypeError, ValueError, RecursionError):
        return json.dumps(tree, **LAYOUT)


def dumps(tree, filename=None):
    """Text of a tree in the reference's layout; also written to `filename` when given."""
    text = pretty(tree)
    if filename:
        with open(filename, "w") as fh:
            fh.write(text)
    return text


def read_text(filename):
    with open(filename, "r") as fh:
        return fh.read()


def source_text(arg):
    """`arg` is JSON text, or the name of a .json file holding it."""
    return read_text(arg) if arg.endswith(".json") else arg


def loads_flat(text, bracketed_spaces=False):
    """The flat reader of Segment / MetaSegment.from_json: the text is cut into words (a bracketed list counts as
    one) and consecutive words are key, value pairs -- values stay strings, as in the reference."""
    inner = r"[\w\s'.-]+" if bracketed_spaces else r"[\w'.-]+"
    words = re.findall(r"\[" + inner + r"\]|[\w'.-]+", text)
    return dict(zip(words[0::2], words[1::2]))


# ---------------------------------------------------------------------------------------------------
# table-first producer
# ---------------------------------------------------------------------------------------------------
class FileTables(object):
    """What the device handed back for one file, in samples: event rows (start, length, mean, std, min, max),
    segment rows (event, start, end, mean, std, min, max) and how the file was parsed."""

    def __init__(self, second, events, segments, filter_params, state_parser, meta=False):
        self.second = float(second)
        self.events = events          # dict of equally long arrays
        self.segments = segments      # dict of equally long arrays, or None (no segmenter ran)
        self.filter_params = filter_params
        self.state_parser = state_parser
        self.meta = meta              # rows describe MetaEvents / MetaSegments (file.to_meta() was called)

    def event_rows(self):
        """List of event dicts in schema order, built column-wise."""
        ev, sec = self.events, self.second
        start = np.asarray(ev["start"], np.int64)
        length = np.asarray(ev["length"], np.int64)
        n_ev = start.shape[0]
        kind = "MetaEvent" if self.meta else "Event"
        cols = {
            "mean": np.asarray(ev["mean"]).tolist(), "std": np.asarray(ev["std"]).tolist(),
            "min": np.asarray(ev["min"]).tolist(), "max": np.asarray(ev["max"]).tolist(),
            "start": (start / sec).tolist(), "end": ((start + length) / sec).tolist(),
            "duration": (length / sec).tolist(),
        }
        filtered = self.filter_params is not None
        seg_rows, bounds = None, None
        if self.segments is not None:
            seg_rows = self.segment_rows()
            bounds = np.searchsorted(np.asarray(self.segments["event"]), np.arange(n_ev + 1)).tolist()
            parser_d = _parser_dict(self.state_parser)
        rows = []
        for i in range(n_ev):
            d = {k: cols[k][i] for k in _STAT + _SPAN}
            if not self.meta:
                d["filtered"] = filtered
            if filtered:
                d["filter_order"], d["filter_cutoff"] = self.filter_params
            if seg_rows is not None:
                d["n"] = bounds[i + 1] - bounds[i]
                d["state_parser"] = dict(parser_d)
                d["segments"] = seg_rows[bounds[i]:bounds[i + 1]]
            else:
                d["n"] = 0   # no state parser to write: File.to_json drops 'segments' along with it
            d["name"] = kind
            rows.append(d)
        return rows

    def segment_rows(self):
        sg, sec = self.segments, self.second
        a = np.asarray(sg["start"], np.int64)
        b = np.asarray(sg["end"], np.int64)
        cols = [np.asarray(sg[k]).tolist() for k in _STAT] + [(a / sec).tolist(), (b / sec).tolist(),
                                                              ((b - a) / sec).tolist()]
        kind = "MetaSegment" if self.meta else "Segment"
        keys = _STAT + _SPAN
        return [dict(zip(keys, vals), name=kind) for vals in zip(*cols)]


def file_tree(f):
    """The nested dict File.to_json writes.  Uses the file's tables when they still describe it (nothing was
    handed out as objects that the caller could have changed), the objects otherwise."""
    d = record(f, "File")
    tables = getattr(f, "_tables", None)
    if tables is not None and f._tables_current():
        d["events"] = tables.event_rows()
    else:
        d["events"] = [expand_event(ev, strict=True) for ev in d["events"]]
    d["event_parser"] = _parser_dict(d["event_parser"])
    return d
