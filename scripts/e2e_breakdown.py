"""Dev: where the end-to-end step spends its time (GPU box)."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypore_b200 import _lib, synth
from pypore_b200.parsers import statsplit_min_gain

x = synth.make_trace(5000, seed=1, tier="A")
pinned = torch.from_numpy(x).pin_memory(); xp = pinned.numpy()
mw, MW, W, gain = statsplit_min_gain(min_width=100, max_width=1000000, window_width=10000)
ctx = _lib.Context(0)
rules = dict(rule_mask=7, duration_gt=1000, duration_lt=0, min_gt=-0.5, max_lt=110.0)

def t(fn, n=10):
    fn(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    ctx.sync()
    return (time.perf_counter() - t0) / n * 1e3

print("upload only            %.3f ms" % t(lambda: ctx.upload_trace_async(xp)))
r = ctx.pipeline(110.0, min_width=mw, max_width=MW, window_width=W, min_gain=gain, with_stats=True, **rules)
print("resident pipeline      %.3f ms" % t(lambda: ctx.pipeline(110.0, min_width=mw, max_width=MW, window_width=W, min_gain=gain, with_stats=True, **rules)))
for ch in (1 << 20, 2 << 20, 4 << 20, 8 << 20, 16 << 20):
    print("host pipeline chunk %3dM %.3f ms" % (ch >> 20, t(lambda: ctx.pipeline(110.0, min_width=mw, max_width=MW, window_width=W, min_gain=gain, with_stats=True, host_trace=xp, chunk_samples=ch, **rules))))
for ch in (4 << 20, 8 << 20, 12 << 20, 16 << 20, 20 << 20, 30 << 20):
    print("host pipeline + host tables, chunk %3dM %.3f ms" % (ch >> 20, t(lambda: ctx.pipeline(110.0, min_width=mw, max_width=MW, window_width=W, min_gain=gain, with_stats=True, host_trace=xp, chunk_samples=ch, export=True, **rules))))
print("events download        %.3f ms" % t(lambda: ctx.events(r["events"])))
print("segments download      %.3f ms" % t(lambda: ctx.segments(r["segments"])))
print("segments download (pinned) %.3f ms" % t(lambda: ctx.segments(r["segments"], pinned=True)))
