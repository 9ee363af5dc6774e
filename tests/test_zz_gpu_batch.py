"""File batches on the GPU (BASELINE configs[4], SURVEY 8e "files are the unit for C5, with no halo"):
pypore_b200.batch.FileBatch -- several worker contexts (streams + host threads) per GPU, files dealt over the
ranks, tables all-gathered -- against the oracle, against the one-file-at-a-time device path, and as
Experiment.parse(..., meta=True, batch=...).  Named test_zz_* so that it runs after the core parity tests."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from batch_common import (FILTER, TIMESTEP, assert_tables_match, detector, make_files, oracle_tables, segmenter)

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _single_file_tables(files):
    """The batch through File.parse, one file after the other, on the default context."""
    from pypore_b200.DataTypes import File
    rows = []
    for i, x in enumerate(files):
        f = File(current=x, timestep=TIMESTEP)
        f.parse(parser=detector(), segmenter=segmenter(), filter_params=FILTER)
        rows.append((i, f.event_table, f.segment_table))
    return rows


def test_file_batch_matches_oracle_and_single_file_path():
    from pypore_b200.batch import FileBatch
    from batch_common import edge_files
    files = make_files(9) + edge_files()        # ragged files; files that start / end inside events
    want = oracle_tables(files)
    # group_samples > 0 (default): several files per resident pass, +inf separators; 0: one file per pass
    for workers, group_samples in ((1, 1 << 26), (3, 1 << 26), (3, 0)):
        b = FileBatch(device=0, workers=workers, group_samples=group_samples)
        try:
            got = b.parse(files, TIMESTEP, detector(), segmenter(), FILTER)
            assert any(len(g) > 1 for g in b.groups) == (group_samples > 0)
            assert_tables_match(got, want)
            again = b.parse(files, TIMESTEP, detector(), segmenter(), FILTER)   # contexts reused across batches
            assert_tables_match(again, got, exact=True)
            if workers == 3 and group_samples:
                for i, et, st in _single_file_tables(files):
                    (e0, e1), (s0, s1) = got.file_rows(i)
                    assert np.array_equal(got.events["start"][e0:e1], et["start"])
                    assert np.array_equal(got.events["length"][e0:e1], et["length"])
                    if e1 > e0:
                        for k in ("event", "start", "end", "min", "max"):
                            assert np.array_equal(got.segments[k][s0:s1], st[k]), k
                        # the statistics kernel reads 16-byte vectors from the row's aligned start, so the order of
                        # its fp64 sums follows the row's position in the pass: same samples at another offset
                        # agree to rounding, not to the bit (the bar against the reference is 1e-9)
                        for k in ("mean", "std"):
                            assert np.allclose(got.segments[k][s0:s1], st[k], rtol=1e-12, atol=0), k
                # without the filter (float32 events, the parallel prefix path) and with default gains
                from pypore_b200.parsers import SpeedyStatSplit
                import oracle
                plain = b.parse(files, TIMESTEP, detector(), SpeedyStatSplit(min_width=100, window_width=10000), None)
                for i in (0, 5):
                    x = files[i].astype(np.float64)
                    (e0, e1), (s0, s1) = plain.file_rows(i)
                    ws, wl = plain.events["start"][e0:e1], plain.events["length"][e0:e1]
                    oe, ost, oen, _ = oracle.statsplit_events(x, ws, wl)
                    assert np.array_equal(plain.segments["event"][s0:s1], oe)
                    assert np.array_equal(plain.segments["start"][s0:s1], ost)
                    assert np.array_equal(plain.segments["end"][s0:s1], oen)
                    m = [np.mean(x[s:s + n]) for s, n in zip(ws, wl)]
                    assert np.allclose(plain.events["mean"][e0:e1], m, rtol=1e-9, atol=0)
        finally:
            b.close()


def _close(a, b, path=""):
    if isinstance(a, dict):
        assert isinstance(b, dict) and sorted(a) == sorted(b), path
        for k in a:
            _close(a[k], b[k], path + "/" + str(k))
    elif isinstance(a, list):
        assert isinstance(b, list) and len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _close(x, y, "%s[%d]" % (path, i))
    elif isinstance(a, float):
        assert abs(a - b) <= 1e-9 * max(abs(a), abs(b)), "%s: %r != %r" % (path, a, b)
    else:
        assert a == b, "%s: %r != %r" % (path, a, b)


def test_experiment_batch_meta_equals_the_sequential_loop(capsys):
    """Experiment.parse(meta=True) through the batch must leave the same metadata (and print the same lines)
    as the reference-shaped loop over files and events (DataTypes.py:956-988)."""
    from pypore_b200.batch import FileBatch
    from pypore_b200.DataTypes import Experiment, File
    files = make_files(5)
    mk = lambda: [File(current=x, timestep=TIMESTEP) for x in files]  # noqa: E731
    seq = Experiment(mk())
    seq.parse(event_detector=detector(), segmenter=segmenter(), filter_params=FILTER, verbose=True, meta=True)
    seq_out = capsys.readouterr().out
    b = FileBatch(device=0, workers=2)
    try:
        bat = Experiment(mk())
        bat.parse(event_detector=detector(), segmenter=segmenter(), filter_params=FILTER, verbose=True, meta=True,
                  batch=b)
        bat_out = capsys.readouterr().out
        with pytest.raises(ValueError):
            Experiment(mk()).parse(event_detector=detector(), segmenter=segmenter(), meta=False, batch=b)
    finally:
        b.close()
    assert bat_out == seq_out and "Detected" in seq_out
    assert bat.n == seq.n == 5 and len(bat.events) == len(seq.events) and len(bat.segments) == len(seq.segments)
    for fs, fb in zip(seq.files, bat.files):
        _close(json.loads(fs.to_json()), json.loads(fb.to_json()))
    assert bat.tables.n_events == len(seq.events)


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_file_batch_torchrun(world):
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29650 + world),
           os.path.join(ROOT, "tests", "batch_worker.py"), "11"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "BATCH OK world=%d" % world in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_split_cta_count_does_not_change_the_tables():
    """PP_OPT_SPLIT_CTAS: the split search is a work queue; a fraction of a wave (what FileBatch gives each of its
    contexts), a handful of CTAs or a single one must produce the tables of the full wave."""
    import oracle
    from pypore_b200 import _lib, synth
    x = synth.make_trace(40, seed=77, tier="A")
    rules = dict(rule_mask=7, duration_gt=1000, duration_lt=0, min_gt=-0.5, max_lt=110.0)
    c = _lib.Context(0)
    try:
        ref = None
        for n in (0, 259, 148, 5, 1, 1 << 20):
            c.set_split_ctas(n)
            r = c.pipeline(110.0, min_width=100, max_width=1000000, window_width=10000, min_gain=-0.0,
                           host_trace=x, **rules)
            seg = c.segments(r["segments"])
            if ref is None:
                ref = seg
                ws, wl = c.events(r["events"])
                oe, ost, oen, _ = oracle.statsplit_events(x.astype(np.float64), ws, wl)
                assert np.array_equal(seg["event"], oe) and np.array_equal(seg["start"], ost)
                assert np.array_equal(seg["end"], oen)
            else:
                assert all(np.array_equal(seg[k], ref[k]) for k in ref), n
        with pytest.raises(ValueError):
            c.set_split_ctas(-1)
    finally:
        c.close()


def test_experiment_batch_reproduces_the_real_references_run():
    """tests/golden/experiment_meta.json: what the REAL reference's Experiment.parse(..., meta=True) printed and left
    behind on the batch test set (generated by tests/golden/make_golden.py --experiment-only).  The device batch
    must reproduce it: names, counts, printed lines, event / segment times exactly (bit-exact indices), statistics of
    the FILTERED events within the north star's 1e-5 (every event is filtered here; measured filter error 2e-9)."""
    from batch_common import assert_json_close, experiment_through_batch
    from pypore_b200.batch import FileBatch
    with open(os.path.join(ROOT, "tests", "golden", "experiment_meta.json")) as f:
        want = json.load(f)
    for kw in (dict(workers=2), dict(workers=2, group_samples=0)):
        b = FileBatch(device=0, **kw)
        try:
            got, printed = experiment_through_batch(b)
        finally:
            b.close()
        assert printed == want["stdout"]
        assert_json_close(got, want["files"], 1e-5)
