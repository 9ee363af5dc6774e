"""Synthetic nanopore traces (no .abf data exists offline).

The generator is the one SURVEY.md App. C.1 specifies, so the known answers of
App. C.3 / D stay valid: piecewise-constant levels with Gaussian noise inside
events, an open channel at 120 pA between them, float32 samples.  Tier A
quantises to a 2**-5 pA ADC grid, which mirrors the int16-ADC origin of real
traces (reference PyPore/read_abf.py:208-210) and makes every fp64 prefix sum
exact in any summation order.
"""
import numpy as np

OPEN_MEAN, OPEN_STD = 120.0, 1.5
ADC_STEP = 1.0 / 32.0


def quantise(x, step=ADC_STEP):
    """Snap to the ADC grid (tier A)."""
    return (np.round(np.asarray(x, np.float64) / step) * step).astype(np.float32)


def _event_body(rng, L, parts):
    used = 0
    while used < L:
        n = min(rng.randint(300, 3000), L - used)
        parts.append(rng.normal(rng.uniform(20, 90), 1.0, n))
        used += n


def make_trace(n_events, seed=0, tier="A", tail=True):
    """Config C1/C2/C3-style trace: ``n_events`` blockades separated by open channel.

    tier "A" = ADC-quantised (bit-exact suite), "B" = raw float32.
    ``tail=False`` omits the closing open-channel gap (the head of a neighbouring
    piece can then be appended, used by the multi-GPU straddle cases).
    """
    rng = np.random.RandomState(seed)
    parts = []
    for _ in range(n_events):
        parts.append(rng.normal(OPEN_MEAN, OPEN_STD, rng.randint(3000, 5000)))
        _event_body(rng, rng.randint(6000, 10000), parts)
    if tail:
        parts.append(rng.normal(OPEN_MEAN, OPEN_STD, rng.randint(3000, 5000)))
    x = np.concatenate(parts).astype(np.float32)
    return quantise(x) if tier == "A" else x


def make_long_event(length, seed=100, tier="A"):
    """Config C4 event body: the inner level loop only, ``length`` samples."""
    rng = np.random.RandomState(seed)
    parts = []
    _event_body(rng, int(length), parts)
    x = np.concatenate(parts).astype(np.float32)
    return quantise(x) if tier == "A" else x


def make_long_trace(n_events, length, seed0=100, tier="A", gap=4000):
    """C4 as one trace: long event bodies separated by short open-channel gaps."""
    rng = np.random.RandomState(seed0 + 7919)
    parts = []
    for e in range(n_events):
        g = rng.normal(OPEN_MEAN, OPEN_STD, gap).astype(np.float32)
        parts.append(quantise(g) if tier == "A" else g)
        parts.append(make_long_event(length, seed0 + e, tier))
    g = rng.normal(OPEN_MEAN, OPEN_STD, gap).astype(np.float32)
    parts.append(quantise(g) if tier == "A" else g)
    return np.concatenate(parts)


def add_subzero_spikes(trace, ev_start, ev_len, frac=0.02, seed=7, width=50, level=-20.0):
    """Plant a short sub-zero spike in a fraction of the events (exercises ``min > -0.5``)."""
    rng = np.random.RandomState(seed)
    out = trace.copy()
    hit = []
    for i, (s, n) in enumerate(zip(ev_start, ev_len)):
        if rng.uniform() < frac and n > 4 * width:
            p = int(s) + int(rng.randint(width, int(n) - 2 * width))
            out[p:p + width] = level
            hit.append(i)
    return out, np.asarray(hit, np.int64)
