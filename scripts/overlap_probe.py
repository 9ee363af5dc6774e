"""Dev (1 GPU): does a 240 MB pinned H2D copy on one context's stream run concurrently with another context's pipeline?"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypore_b200 import _lib, synth
from pypore_b200.parsers import statsplit_min_gain

x = synth.make_trace(5000, seed=1, tier="A")
xp = torch.from_numpy(x).pin_memory().numpy()
A, B = _lib.Context(0), _lib.Context(0)
rules = dict(rule_mask=7, duration_gt=1000, duration_lt=0, min_gt=-0.5, max_lt=110.0)
mw, MW, W, gain = statsplit_min_gain(min_width=100, max_width=1000000, window_width=10000)
A.upload_trace(xp); B.upload_trace(xp)
for _ in range(3):
    A.pipeline(110.0, min_width=mw, max_width=MW, window_width=W, min_gain=gain, **rules)
def t_pipe():
    t0 = time.perf_counter(); A.pipeline(110.0, min_width=mw, max_width=MW, window_width=W, min_gain=gain, **rules); return (time.perf_counter() - t0) * 1e3
print("pipeline alone            %.2f ms" % min(t_pipe() for _ in range(5)))
t0 = time.perf_counter(); B.upload_trace_async(xp); B.sync(); print("upload alone              %.2f ms" % ((time.perf_counter() - t0) * 1e3))
for _ in range(3):
    t0 = time.perf_counter()
    B.upload_trace_async(xp)
    t1 = time.perf_counter()
    tp = t_pipe()
    t2 = time.perf_counter()
    B.sync()
    t3 = time.perf_counter()
    print("upload enqueue %.2f ms, pipeline under it %.2f ms, upload done after %.2f ms total" % ((t1 - t0) * 1e3, tp, (t3 - t0) * 1e3))
