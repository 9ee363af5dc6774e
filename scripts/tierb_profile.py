import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from pypore_b200 import _lib, synth
from pypore_b200.parsers import statsplit_min_gain
ctx = _lib.Context(0)
x = synth.make_trace(500, seed=0, tier="B")
ctx.upload_trace(x)
mw, MW, W, gain = statsplit_min_gain(min_width=100, max_width=1000000, window_width=10000)
for _ in range(3):
    r = ctx.pipeline(110.0, 7, 1000, 0, -0.5, 110.0, mw, MW, W, gain)
print(r, ctx.stage_ms(), ctx.split_counters())
