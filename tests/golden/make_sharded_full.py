"""tests/golden/sharded_full.npz: the multi-GPU bench workloads at FULL size (bench.py --gpus 2 / 4 / 8: world x 5000
events, 120 M / 240 M / 480 M samples, pypore_b200.dist.synthetic_global(world, 5000, seed0=1)) through the CPU oracle:
counts and SHA-256 of the global event rows (start, length) and segment rows (global event id, start, end).

    python tests/golden/make_sharded_full.py        # ~70 s, ~12 GB of RAM at world 8

The oracle makes the fixture; tests/golden/make_reference_full_check.py then runs the REAL reference on the same traces
and records that its table hashes equal these (tests/golden/reference_full_check.json)."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from pypore_b200 import dist as ppdist  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    rules = [lambda e: e.duration > 1000, lambda e: e.min > -0.5, lambda e: e.max < 110]
    out = {}
    for world in (2, 4, 8):
        x = ppdist.synthetic_global(world, 5000, seed0=1).astype(np.float64)
        ws, wl = oracle.events(x, 110, rules)
        oe, ost, oen, nc = oracle.statsplit_events(x, ws, wl, min_width=100, max_width=1000000, window_width=10000,
                                                   threads=os.cpu_count() or 1)
        k = "w%d_" % world
        out[k + "samples"], out[k + "events"] = np.int64(len(x)), np.int64(len(ws))
        out[k + "event_samples"], out[k + "segments"] = np.int64(wl.sum()), np.int64(len(oe))
        out[k + "candidates"] = np.int64(nc)
        out[k + "events_sha"] = np.array(sha(np.stack([ws, wl], axis=1).astype(np.int64)))
        out[k + "segments_sha"] = np.array(sha(np.stack([oe, ost, oen], axis=1).astype(np.int64)))
        print(world, {n: v.item() for n, v in out.items() if n.startswith(k) and v.ndim == 0 and v.dtype != "<U64"})
        del x
    np.savez_compressed(os.path.join(HERE, "sharded_full.npz"), **out)


if __name__ == "__main__":
    main()
