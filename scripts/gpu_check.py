"""Dev check (GPU box): C-ABI pipeline vs the oracle on a C1-like trace. Not a test, not a bench."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from pypore_b200 import synth, _lib

n_events = int(sys.argv[1]) if len(sys.argv) > 1 else 500
tier = sys.argv[2] if len(sys.argv) > 2 else "A"
x = synth.make_trace(n_events, seed=0, tier=tier)
x64 = x.astype(np.float64)
ctx = _lib.Context(0)
ctx.upload_trace(x)
for psps in (None, 10):
    mg = oracle.min_gain(prior_segments_per_second=psps)
    t = time.time()
    r = ctx.pipeline(110.0, 7, 1000, 0, -0.5, 110.0, 100, 1000000, 10000, mg)
    print("pipeline", r, "wall %.1f ms" % ((time.time() - t) * 1e3), ctx.stage_ms(), ctx.split_counters())
    t = time.time()
    r = ctx.pipeline(110.0, 7, 1000, 0, -0.5, 110.0, 100, 1000000, 10000, mg)
    print("pipeline again wall %.1f ms" % ((time.time() - t) * 1e3), ctx.stage_ms())
    es, el = ctx.events(r["events"])
    rules = [lambda e: e.duration > 1000, lambda e: e.min > -0.5, lambda e: e.max < 110]
    os_, ol = oracle.events(x64, 110, rules)
    print("events equal:", np.array_equal(es, os_) and np.array_equal(el, ol), len(es), len(os_))
    seg = ctx.segments(r["segments"])
    oe, ost, oen, nc = oracle.statsplit_events(x64, os_, ol, gain=mg, threads=8)
    same = len(oe) == r["segments"] and np.array_equal(seg["event"], oe) and np.array_equal(seg["start"], ost) and np.array_equal(seg["end"], oen)
    print("segments equal:", same, r["segments"], len(oe), "cand", nc)
    if not same:
        k = min(len(oe), r["segments"])
        bad = np.nonzero((seg["event"][:k] != oe[:k]) | (seg["start"][:k] != ost[:k]) | (seg["end"][:k] != oen[:k]))[0]
        print("first mismatches", bad[:10], [(seg["event"][i], seg["start"][i], seg["end"][i], oe[i], ost[i], oen[i]) for i in bad[:5]])
    else:
        m, s, mn, mx = [], [], [], []
        for e in range(min(len(os_), 40)):
            sel = oe == e
            a = oracle.segment_stats(x64[os_[e]:os_[e] + ol[e]], ost[sel], oen[sel])
            m.append(a[0]); s.append(a[1]); mn.append(a[2]); mx.append(a[3])
        k = sum(len(v) for v in m)
        for name, ref in (("mean", m), ("std", s), ("min", mn), ("max", mx)):
            ref = np.concatenate(ref)
            err = np.max(np.abs(seg[name][:k] - ref) / np.maximum(np.abs(ref), 1e-300))
            print(name, "max rel err", err)
    runs = ctx.runs(r["runs"])
    orun = oracle.threshold_runs(x64, 110)
    print("runs equal:", all(np.array_equal(a, b) for a, b in zip(runs, orun)))
