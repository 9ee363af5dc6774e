"""tests/golden/c5_files.npz: BASELINE configs[4] at full FILE size -- eight 250 kHz files of 2.49 M samples
(synth.make_trace(208, seed=900 + i), the files scripts/c5_batch.py cycles through), Event.filter(1, 2000) then
SpeedyStatSplit per event -- through the CPU oracle (scipy-equivalent filtfilt restated in C, the restated split
search; tests/batch_common.py::oracle_file_pass): counts and SHA-256 of the batch's event rows (file, start, length) and
segment rows (file, event, start, end) for the tutorial gain (prior_segments_per_second=10) and the default gain.

    python tests/golden/make_c5_files.py            # ~25 s
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from batch_common import FILTER, FS, detector, oracle_file_pass  # noqa: E402
from pypore_b200 import synth  # noqa: E402
from pypore_b200.batch import FileBatch  # noqa: E402
from pypore_b200.parsers import SpeedyStatSplit  # noqa: E402

SETTINGS = {"psps10": dict(min_width=100, window_width=10000, sampling_freq=FS, cutoff_freq=2000.,
                           prior_segments_per_second=10),
            "default": dict(min_width=100, window_width=10000)}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def table_hashes(t):
    ev = np.stack([t.events[k] for k in ("file", "start", "length")], axis=1).astype(np.int64)
    sg = np.stack([t.segments[k] for k in ("file", "event", "start", "end")], axis=1).astype(np.int64)
    return len(ev), len(sg), sha(ev), sha(sg)


def main():
    files = [synth.make_trace(208, seed=900 + i, tier="A") for i in range(8)]
    out = dict(samples=np.int64(sum(len(f) for f in files)))
    for name, kw in SETTINGS.items():
        t = FileBatch(workers=4, file_pass=oracle_file_pass).parse(files, 1000. / FS, detector(), SpeedyStatSplit(**kw),
                                                                   FILTER)
        ne, ns, he, hs = table_hashes(t)
        out[name + "_events"], out[name + "_segments"] = np.int64(ne), np.int64(ns)
        out[name + "_events_sha"], out[name + "_segments_sha"] = np.array(he), np.array(hs)
        print(name, ne, ns)
    np.savez_compressed(os.path.join(HERE, "c5_files.npz"), **out)


if __name__ == "__main__":
    main()
