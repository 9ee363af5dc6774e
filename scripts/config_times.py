"""Dev: stage times of the other BASELINE configs on one GPU (C1, C4 scaled, C5 one file)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypore_b200 import _lib, synth
from pypore_b200.parsers import statsplit_min_gain
from pypore_b200.DataTypes import bessel_coefficients

ctx = _lib.Context(0)
rules = dict(rule_mask=7, duration_gt=1000, duration_lt=0, min_gt=-0.5, max_lt=110.0)


def run(name, x, reps=5, filt=None, **kw):
    mw, MW, W, gain = statsplit_min_gain(**kw)
    ctx.upload_trace(x)
    for _ in range(2):
        r = ctx.pipeline(110.0, min_width=mw, max_width=MW, window_width=W, min_gain=gain, filter_ba=filt, **rules)
    ms = []
    st = None
    for _ in range(reps):
        t0 = time.perf_counter()
        r = ctx.pipeline(110.0, min_width=mw, max_width=MW, window_width=W, min_gain=gain, filter_ba=filt, **rules)
        ms.append((time.perf_counter() - t0) * 1e3)
        st = ctx.stage_ms()
    c = ctx.split_counters()
    print("%-34s n=%9d ev=%5d seg=%7d cand=%10d exact=%8d  %.3f ms  %s" % (
        name, len(x), r["events"], r["segments"], c["candidates"], c["exact"], min(ms),
        {k: round(v, 3) for k, v in st.items()}), flush=True)


x1 = synth.make_trace(500, seed=0, tier="A")
run("C1 min_gain=0", x1, min_width=100, max_width=1000000, window_width=10000)
run("C1 psps=10", x1, min_width=100, max_width=1000000, window_width=10000, prior_segments_per_second=10)
x2 = synth.make_trace(5000, seed=1, tier="A")
run("C2 psps=10", x2, min_width=100, max_width=1000000, window_width=10000, prior_segments_per_second=10)
x1b = synth.make_trace(500, seed=0, tier="B")
run("C1 tier B (raw float32)", x1b, min_width=100, max_width=1000000, window_width=10000)
nlong = int(os.environ.get("NLONG", "4"))
x4 = synth.make_long_trace(nlong, 10_000_000, seed0=100, tier="A")
run("C4 %d x 10M, max_width=1e6" % nlong, x4, reps=2, min_width=100, max_width=1000000, window_width=10000)
run("C4 %d x 10M, psps=10" % nlong, x4, reps=2, min_width=100, max_width=1000000, window_width=10000,
    prior_segments_per_second=10)
# C5: one 10 s file at 250 kHz, Bessel order 1 / 2 kHz, then the split with the matching gain
rng = np.random.RandomState(1000)
x5 = synth.make_trace(208, seed=1000, tier="A")[:2_500_000]
filt = bessel_coefficients(1, 2000., 2.5e5)
run("C5 file, filter(1,2000) psps=10", x5, filt=filt, min_width=100, max_width=1000000, window_width=10000,
    sampling_freq=2.5e5, cutoff_freq=2000., prior_segments_per_second=10)
run("C5 file, filter(1,2000) gain=0", x5, filt=filt, min_width=100, max_width=1000000, window_width=10000)
