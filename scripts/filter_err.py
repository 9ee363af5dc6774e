import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from pypore_b200 import _lib
from pypore_b200.DataTypes import bessel_coefficients
ctx = _lib.Context(0)
rng = np.random.RandomState(3)
for order, fs in ((1, 1e5), (2, 1e5), (3, 2.5e5), (4, 1e5), (5, 1e5), (6, 1e5), (8, 1e5)):
    b, a, zi = bessel_coefficients(order, 2000., fs)
    pad = 3 * (order + 1)
    lens = [pad + 1, pad + 2, 100, 4095, 4096 - 2 * pad, 4097, 8192, 20011]
    evs = [60 + rng.normal(0, 2, n) for n in lens]
    ctx.upload_events_f64(evs)
    ctx.filter_events(b, a, zi)
    y = ctx.event_samples(sum(lens))
    k = 0
    errs = []
    for e in evs:
        ref = oracle.filtfilt(b, a, e)
        got = y[k:k + len(e)]
        errs.append(float(np.max(np.abs(got - ref) / np.abs(ref))))
        k += len(e)
    print(order, fs, ["%.1e" % v for v in errs])
