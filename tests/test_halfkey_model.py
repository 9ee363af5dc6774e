"""A host model of the K3_CFG_HALFKEY development variant of the split search (pypore_b200/csrc/split.cuh,
DESIGN.md section 4 "planned next"): which half of a candidate's screening key a child window may inherit from
its parent's scan.  The model replays the kernel's level loop (k3_place / k3_resolve / the mode choice of
k3_screen_level) on real data, with the oracle taking the split decisions, and stores per candidate position the
ANCHOR each half-key was computed against instead of the key.  It asserts that

* every inherited half was stored against exactly the anchor the inheriting scan needs (L: the window start,
  R: the window end), for every candidate of the scan;
* scans of one level never read or write positions another scan of that level writes (they run concurrently);
* the replay is faithful: its breakpoints are the oracle's.

The variant itself has not run on a GPU yet; this pins its bookkeeping rules."""
import numpy as np
import pytest

import oracle
from oracle.oracle import _sequential_best, _window_gains
from pypore_b200 import synth


class Item(object):
    def __init__(self, s, e, ps, pad=0):
        self.s, self.e, self.ps, self.pad = s, e, ps, pad


def replay(x, mw, MW, W, gain, rng, p_full=0.0, p_away=0.0):
    c, c2 = oracle.cumsum(x)
    L = len(x)
    ancL = np.full(L + 1, -7, np.int64)    # anchor of the stored left half-key at candidate position j
    ancR = np.full(L + 1, -7, np.int64)
    bps, stats = [], dict(scans=0, inherited=0, sides=0)

    def worth(s, e):
        return (e - s > 2 * mw) or (e - s > MW)

    def forced(s, e):
        return min(s + MW, e - mw)

    def place(it, nxt):                     # k3_place
        stack = []
        while True:
            while True:
                if it.ps >= it.e - 2 * mw:
                    if it.e - it.s > MW:
                        xx = forced(it.s, it.e)
                        bps.append(xx)
                        l, r = Item(it.s, xx, it.s), Item(xx, it.e, xx)
                        wl, wr = worth(l.s, l.e), worth(r.s, r.e)
                        if wl and wr:
                            stack.append(r)
                            it = l
                            continue
                        if wl:
                            it = l
                            continue
                        if wr:
                            it = r
                            continue
                    break
                if it.ps > it.s + MW:
                    xx = forced(it.s, it.e)
                    bps.append(xx)
                    if not worth(xx, it.e):
                        break
                    it.s = it.ps = xx
                    it.pad = 0
                    continue
                pe = min(it.ps + W, it.e)
                if pe - it.ps <= 2 * mw:
                    it.ps = min(it.ps + W // 2, it.e)
                    continue
                nxt.append(it)
                break
            if not stack:
                break
            it = stack.pop()

    level = []
    place(Item(0, L, 0), level)
    while level:
        nxt, written, touched = [], {}, {}
        decisions = []
        for k, it in enumerate(level):      # SCREEN: all scans of a level run concurrently
            pe = min(it.ps + W, it.e)
            cand = np.arange(it.ps + mw, pe - mw + 1)
            full = rng.uniform() < p_full   # a window that ends up scanned exactly (K3_FULL_FLAG)
            screened = not (full and rng.uniform() < 0.5)    # ... with or without having been screened first
            mode = (it.pad & 3) if (it.ps == it.s and pe == it.e) else 0
            stats["scans"] += 1
            if screened:
                stats["sides"] += 2 * len(cand)
                if mode == 1:
                    assert np.all(ancL[cand] == it.ps), "left half inherited with the wrong anchor"
                    stats["inherited"] += len(cand)
                if mode == 2:
                    assert np.all(ancR[cand] == pe), "right half inherited with the wrong anchor"
                    stats["inherited"] += len(cand)
                for j in cand:              # reads and writes of this scan
                    assert touched.get(j, k) == k and written.get(j, k) == k, "scans of one level overlap"
                    touched[j] = k
                    written[j] = k
                store = pe - it.ps > 3 * mw                 # leaf scans (no child can be scanned) store nothing
                if mode != 1 and store:
                    ancL[cand] = -1 if full else it.ps      # a failed validity test leaves garbage behind
                if mode != 2 and store:
                    ancR[cand] = -1 if full else pe
            g = _window_gains(c, c2, it.ps, pe, it.ps + mw, pe - mw)
            decisions.append((it, pe, _sequential_best(g, gain, it.ps + mw)[1], not full))
        for it, pe, xx, reuse in decisions:  # DECIDE / resolve (k3_resolve)
            if xx >= 0:
                bps.append(xx)
                if worth(it.s, xx):
                    place(Item(it.s, xx, it.s, 1 if (reuse and it.ps == it.s) else 0), nxt)
                if worth(xx, it.e):
                    away = rng.uniform() < p_away        # handed to the global queue: a fresh task elsewhere
                    place(Item(xx, it.e, xx, 2 if (reuse and pe == it.e and not away) else 0), nxt)
            else:
                place(Item(it.s, it.e, min(it.ps + W // 2, it.e), 0), nxt)
        level = nxt
    return np.array(sorted(bps), np.int64), stats


@pytest.mark.parametrize("kw,length,seed", [
    (dict(min_width=100, max_width=1000000, window_width=10000), 9000, 1),          # the headline setting
    (dict(min_width=100, max_width=1000000, window_width=10000, prior_segments_per_second=10), 9000, 2),
    (dict(min_width=50, max_width=2500, window_width=1000, prior_segments_per_second=50), 12000, 3),  # windows, forced
    (dict(min_width=20, max_width=700, window_width=300), 5000, 4),
    (dict(min_width=100, max_width=1000000, window_width=200), 3000, 5),            # window == 2 * min_width
])
def test_inherited_halves_have_the_right_anchor(kw, length, seed):
    x = synth.make_long_event(length, seed=seed, tier="A").astype(np.float64)
    gain = oracle.min_gain(**kw)
    mw, MW, W = kw["min_width"], kw["max_width"], kw["window_width"]
    want = oracle.statsplit(x, **kw)
    for p_full, p_away in ((0.0, 0.0), (0.2, 0.2)):
        got, st = replay(x, mw, MW, W, gain, np.random.RandomState(seed), p_full, p_away)
        assert np.array_equal(got, want)
        if W >= length and p_full == 0.0 and MW >= length:
            # an event shorter than the window: every scan below the root inherits one half
            assert st["inherited"] > 0.3 * st["sides"]
