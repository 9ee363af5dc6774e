"""tests/golden/bench_configs.npz: the bench.py workloads of BASELINE configs[0], [2] and [3] at FULL size.

    python tests/golden/make_bench_configs.py       # ~3 min, ~8 GB of RAM

* c1 (configs[0]: synth.make_trace(500, seed=0), 6 M samples) goes through the REAL reference (File.parse + the
  compiled cparsers per event, loaded like tests/golden/make_golden.py does) -- it is the reference's own CPU-runnable case;
* c3 (configs[2]: synth.make_trace(30000, seed=2), 360 M samples) and c4 (configs[3]: 20 events of 10 M samples,
  synth.make_long_trace(20, 10_000_000, seed0=100), max_width = 1e6) go through the CPU oracle (the reference's Python
  layer needs tens of GB and minutes for these); the oracle hashes to the real reference's tables on configs[1] and on
  the 4-event version of configs[3] (tests/golden/c2_full.npz, c4_full.npz, tests/test_oracle_golden.py).
Stored: counts and SHA-256 of the event rows (start, length) and segment rows (event, start, end), int64."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import oracle  # noqa: E402
from pypore_b200 import synth  # noqa: E402

RULES = [lambda e: e.duration > 1000, lambda e: e.min > -0.5, lambda e: e.max < 110]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def through_oracle(x64, key, out, **kw):
    ws, wl = oracle.events(x64, 110, RULES)
    oe, ost, oen, nc = oracle.statsplit_events(x64, ws, wl, threads=os.cpu_count() or 1, **kw)
    out[key + "samples"], out[key + "events"] = np.int64(len(x64)), np.int64(len(ws))
    out[key + "event_samples"], out[key + "segments"] = np.int64(wl.sum()), np.int64(len(oe))
    out[key + "candidates"] = np.int64(nc)
    out[key + "events_sha"] = np.array(sha(np.stack([ws, wl], axis=1).astype(np.int64)))
    out[key + "segments_sha"] = np.array(sha(np.stack([oe, ost, oen], axis=1).astype(np.int64)))
    print(key, {n[len(key):]: v.item() for n, v in out.items() if n.startswith(key) and v.dtype.kind == "i"}, flush=True)


def main():
    out = {}
    import make_golden
    dt, parsers, _ = make_golden.load_reference()
    x64 = synth.make_trace(500, seed=0, tier="A").astype(np.float64)
    f = dt.File(current=x64, timestep=0.01)
    f.parse(parser=parsers.lambda_event_parser(threshold=110, rules=RULES))
    ev = np.array([(int(round(e.start * f.second)), len(e.current)) for e in f.events], np.int64).reshape(-1, 2)
    rows = []
    for k, event in enumerate(f.events):
        for seg in parsers.SpeedyStatSplit(min_width=100, window_width=10000).parse(event.current):
            rows.append((k, seg.start, seg.end))
    rows = np.array(rows, np.int64).reshape(-1, 3)
    out.update(c1_samples=np.int64(len(x64)), c1_events=np.int64(len(ev)), c1_event_samples=np.int64(ev[:, 1].sum()),
               c1_segments=np.int64(len(rows)), c1_events_sha=np.array(sha(ev)), c1_segments_sha=np.array(sha(rows)))
    print("c1 (real reference)", len(ev), len(rows), flush=True)
    chk = {}
    through_oracle(x64, "chk_", chk, min_width=100, max_width=1000000, window_width=10000)
    assert str(chk["chk_events_sha"]) == str(out["c1_events_sha"]) and str(chk["chk_segments_sha"]) == str(out["c1_segments_sha"])
    out["c1_candidates"] = chk["chk_candidates"]
    del x64, f
    through_oracle(synth.make_trace(30000, seed=2, tier="A").astype(np.float64), "c3_", out,
                   min_width=100, max_width=1000000, window_width=10000)
    through_oracle(synth.make_long_trace(20, 10_000_000, seed0=100, tier="A").astype(np.float64), "c4_", out,
                   min_width=100, max_width=1000000, window_width=10000)
    np.savez_compressed(os.path.join(HERE, "bench_configs.npz"), **out)


if __name__ == "__main__":
    main()
