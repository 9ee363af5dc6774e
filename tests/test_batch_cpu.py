"""File batches (BASELINE configs[4], SURVEY 8e "files are the unit") -- host logic on CPU: round-robin
ownership, worker threads, the ragged table all-gather over gloo (world_size 2 and 3) and the ordering of
the gathered rows.  The per-file device pass is played by the oracle (tests may use it); tests/
test_zz_gpu_batch.py runs the real one."""
import json
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from batch_common import (FILTER, TIMESTEP, assert_tables_match, detector, make_files, oracle_file_pass,
                          oracle_tables, segmenter)
from pypore_b200.batch import BatchTables, FileBatch, assign_files


def test_assign_files_partitions_the_batch():
    for n, world in ((0, 2), (1, 4), (7, 2), (8, 8), (1000, 8), (5, 3)):
        owned = [assign_files(n, r, world) for r in range(world)]
        assert sorted(i for o in owned for i in o) == list(range(n))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def test_worker_threads_give_the_single_thread_tables():
    files = make_files(7)
    want = oracle_tables(files)
    assert want.n_files == 7 and want.n_events > 10 and want.n_segments > want.n_events
    (e0, e1), (s0, s1) = want.file_rows(2)
    assert e0 == e1 and s0 == s1                          # the file without events contributes no rows
    got = FileBatch(workers=3, file_pass=oracle_file_pass).parse(files, [TIMESTEP] * 7, detector(), segmenter(),
                                                                 FILTER)
    assert_tables_match(got, want, exact=True)
    with pytest.raises(ValueError):
        FileBatch(workers=1, file_pass=oracle_file_pass).parse(files, [TIMESTEP] * 3, detector(), segmenter(), FILTER)

    def failing(ctx, current, *a):
        raise RuntimeError("boom")
    with pytest.raises(RuntimeError, match="boom"):       # a worker's exception reaches the caller
        FileBatch(workers=2, file_pass=failing).parse(files, TIMESTEP, detector(), segmenter(), FILTER)


def test_tables_to_meta_files_like_the_reference():
    """BatchTables.files(): what Experiment.parse(..., meta=True) leaves behind (DataTypes.py:956-988):
    Files of MetaEvents of MetaSegments, times in seconds, JSON-serialisable like the reference's."""
    files = make_files(4)
    t = oracle_tables(files)
    out = t.files(["f%d" % i for i in range(4)], detector(), segmenter(), FILTER)
    assert [f.filename for f in out] == ["f0", "f1", "f2", "f3"] and out[2].n == 0
    second = 1000. / TIMESTEP
    k = 0
    for i, f in enumerate(out):
        assert not hasattr(f, "current")
        for e, ev in enumerate(f.events):
            assert type(ev).__name__ == "MetaEvent" and ev.filtered and ev.filter_order == 1
            assert ev.start == t.events["start"][k] / second and ev.duration == t.events["length"][k] / second
            assert ev.mean == t.events["mean"][k] and ev.n == len(ev.segments) > 0
            assert ev.segments[0].start == 0.0 and abs(ev.segments[-1].end - ev.duration) < 1e-12
            assert all(type(s).__name__ == "MetaSegment" for s in ev.segments)
            k += 1
        d = json.loads(f.to_json())
        assert d["name"] == "File" and d["n"] == f.n and len(d["events"]) == f.n
        if f.n:
            assert d["events"][0]["name"] == "MetaEvent" and d["events"][0]["segments"][0]["name"] == "MetaSegment"
            assert d["events"][0]["state_parser"]["name"] == "SpeedyStatSplit"
    assert k == t.n_events


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_files, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        files = make_files(n_files)
        want = oracle_tables(files)
        owned = set(assign_files(n_files, rank, world))
        mine = [f if i in owned else None for i, f in enumerate(files)]   # a rank only holds its own files
        b = FileBatch(workers=2, rank=rank, world=world, file_pass=oracle_file_pass)
        got = b.parse(mine, TIMESTEP, detector(), segmenter(), FILTER)
        assert sorted(b.local) == sorted(owned)
        assert_tables_match(got, want, exact=True)        # row for row the one-process table, on every rank
        q.put((rank, "ok"))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_files", [(2, 5), (3, 4), (2, 1)])
def test_file_batch_gloo(world, n_files):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_files, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", "rank %d: %s" % (rank, msg)


def test_block_merge_equals_a_stable_row_sort():
    """The gathered rows are rank-major; FileBatch moves whole per-file blocks into file order.  Must equal a
    stable sort of the rows by file for any dealing of ragged (also empty) files over any number of ranks."""
    from pypore_b200.batch import _merge_by_file
    rng = np.random.RandomState(0)
    for _ in range(60):
        world, nf = rng.randint(1, 6), rng.randint(0, 14)
        rows = [(f, rng.randint(100)) for r in range(world) for f in range(r, nf, world)
                for _ in range(rng.randint(0, 4))]
        gi = np.array(rows, np.int64).reshape(-1, 2)
        gf = rng.rand(len(rows), 4)
        oi, of = _merge_by_file(gi, gf)
        o = np.argsort(gi[:, 0], kind="stable")
        assert np.array_equal(oi, gi[o]) and np.array_equal(of, gf[o])


def test_plan_groups_and_groupable():
    from pypore_b200.batch import groupable, plan_groups
    from pypore_b200.parsers import RuleSet, lambda_event_parser
    assert plan_groups({0: 10, 1: 10, 2: 10, 3: 0, 4: 50, 5: 5}, [1] * 6, range(6), 25) == [[0, 1], [2], [3], [4], [5]]
    assert plan_groups({0: 10, 2: 10, 4: 10}, [1, 1, 2, 2, 2], [0, 2, 4], 100) == [[0], [2, 4]]   # sampling rates
    assert plan_groups({i: 10 for i in range(10)}, [1] * 10, range(10), 1000, min_groups=4) == \
        [[0, 1], [2, 3], [4, 5], [6, 7], [8, 9]]
    assert plan_groups({0: 5, 1: 7}, [1, 1], [0, 1], 0) == [[0], [1]] and plan_groups({}, [], [], 10) == []
    assert groupable(lambda_event_parser(threshold=110, rules=RuleSet(duration_gt=1000, min_gt=-0.5, max_lt=110)))
    assert groupable(lambda_event_parser(threshold=110, rules=RuleSet(max_lt=90)))
    assert groupable(lambda_event_parser(threshold=110))                                    # reference defaults
    assert not groupable(lambda_event_parser(threshold=110, rules=RuleSet(duration_gt=1000, min_gt=-0.5)))
    assert not groupable(lambda_event_parser(threshold=110, rules=RuleSet(max_lt=111)))    # a run above could pass
    assert not groupable(lambda_event_parser(threshold=110, rules=[lambda e: e.max < 110]))  # opaque callables


def test_grouped_passes_equal_file_by_file_passes():
    """Several files behind each other in one trace with +inf separators (batch.device_group_pass), the
    oracle playing the device on the concatenated trace: per-file rows must equal file-by-file passes."""
    from batch_common import OracleContext
    from pypore_b200.batch import device_file_pass, device_group_pass
    from batch_common import edge_files
    files = edge_files()
    second = 1000. / TIMESTEP
    ctx = OracleContext()
    for filt in (FILTER, None):
        single = [device_file_pass(ctx, x, second, detector(), segmenter(), filt) for x in files]
        for cut in ([0, 7], [0, 3, 7], [0, 1, 2, 4, 7]):
            grouped = []
            for a, b in zip(cut[:-1], cut[1:]):
                grouped += (device_group_pass(ctx, files[a:b], second, detector(), segmenter(), filt) if b - a > 1
                            else [device_file_pass(ctx, files[a], second, detector(), segmenter(), filt)])
            assert len(grouped) == len(single)
            for g, w in zip(grouped, single):
                for k in ("ev_int", "ev_flt", "seg_int", "seg_flt"):
                    assert g[k].shape == w[k].shape and np.array_equal(g[k], w[k]), k
    assert sum(r["ev_int"].shape[0] for r in single) >= 10 and single[4]["ev_int"].shape[0] == 0
    from pypore_b200.parsers import RuleSet, lambda_event_parser
    with pytest.raises(TypeError):
        device_group_pass(ctx, files[:2], second, lambda_event_parser(threshold=110, rules=RuleSet(duration_gt=10)),
                          segmenter(), None)
    # FileBatch end to end with grouping on: same tables as without
    fb = lambda gs: FileBatch(workers=2, contexts=[OracleContext(), OracleContext()], group_samples=gs)  # noqa: E731
    want = fb(0).parse(files, TIMESTEP, detector(), segmenter(), FILTER)
    b = fb(40000)
    got = b.parse(files, TIMESTEP, detector(), segmenter(), FILTER)
    assert any(len(g) > 1 for g in b.groups) and sorted(i for g in b.groups for i in g) == list(range(7))
    assert_tables_match(got, want, exact=True)


def test_experiment_batch_reproduces_the_real_references_run():
    """tests/golden/experiment_meta.json is what the REAL reference's Experiment.parse(..., meta=True) printed and
    left behind (every file's to_json) on the batch test set.  Host side of the check: the batch driver with the
    oracle playing the device must reproduce it -- file names, event / segment times, parser dictionaries, counts
    and printed lines exactly, statistics to 1e-9."""
    from batch_common import OracleContext, assert_json_close, experiment_through_batch
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "experiment_meta.json")) as f:
        want = json.load(f)
    got, printed = experiment_through_batch(FileBatch(workers=2, contexts=[OracleContext(), OracleContext()],
                                                      group_samples=30000))
    assert printed == want["stdout"]
    assert_json_close(got, want["files"], 1e-9)


def test_experiment_batch_argument_errors():
    from batch_common import OracleContext
    from pypore_b200.DataTypes import Experiment, File
    from pypore_b200.parsers import lambda_event_parser
    b = FileBatch(workers=1, contexts=[OracleContext()])
    mk = lambda: Experiment([File(current=x, timestep=TIMESTEP) for x in make_files(2)])  # noqa: E731
    with pytest.raises(ValueError, match="meta=True"):
        mk().parse(event_detector=detector(), segmenter=segmenter(), meta=False, batch=b)
    with pytest.raises(ValueError, match="segmenter"):
        mk().parse(event_detector=detector(), segmenter=None, meta=True, batch=b)
    with pytest.raises(TypeError, match="device-evaluable"):      # opaque Python rules cannot run on the device
        mk().parse(event_detector=lambda_event_parser(threshold=110, rules=[lambda e: e.max < 110]),
                   segmenter=segmenter(), meta=True, batch=b)
