"""Multi-GPU host logic on CPU: world_size-2 (and 3) gloo processes exercise the boundary
planning, the halo exchange and the ragged table all-gather of pypore_b200.dist.  The local
threshold scan is played by the oracle here (tests may use it); on the GPU box the same
functions run over NCCL with the CUDA scan (tests/test_gpu_dist.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from pypore_b200 import dist as ppdist

RULES = dict(rule_mask=7, duration_gt=1000, duration_lt=0, min_gt=-0.5, max_lt=110.0)
PYRULES = [lambda e: e.duration > 1000, lambda e: e.min > -0.5, lambda e: e.max < 110]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def local_events(chunk, rank, world, group=None, device="cpu"):
    """What ShardedPipeline.step does up to event selection, with the oracle as the scanner."""
    runs = oracle.threshold_runs(chunk.astype(np.float64), 110.0)
    n_runs = len(runs[0])
    info = ppdist.boundary_info(len(chunk), n_runs, [a[0] for a in runs], [a[-1] for a in runs])
    infos = ppdist.gather_infos(info, dist, device, group)
    plan = ppdist.plan_boundaries(infos, RULES)[rank]
    halo = torch.empty(200000, dtype=torch.float32)
    got = ppdist.exchange_halo(plan, torch.from_numpy(chunk), halo, dist, group)
    ext = np.concatenate([chunk, halo[:got].numpy()])
    keep = []
    for i in range(n_runs):
        if (i == 0 and plan["skip_first"]) or (i == n_runs - 1 and plan["skip_last"]):
            continue
        if ppdist.rules_accept(RULES, runs[1][i], runs[2][i], runs[3][i]):
            keep.append((int(runs[0][i]), int(runs[1][i])))
    if plan["event"] is not None:
        keep.append(plan["event"])
    offsets = np.concatenate(([0], np.cumsum(infos[:, ppdist.I_N].astype(np.int64))))
    return keep, ext, offsets, plan


def _worker(rank, world, port, epr, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        chunk = ppdist.synthetic_chunk(rank, world, epr, seed0=50)
        keep, ext, offsets, plan = local_events(chunk, rank, world)
        glob = ppdist.synthetic_global(world, epr, seed0=50)
        assert offsets[-1] == len(glob)
        # every owned event's samples (chunk + halo) equal the global trace there
        for s, n in keep:
            assert s + n <= len(ext)
            assert np.array_equal(ext[s:s + n], glob[offsets[rank] + s:offsets[rank] + s + n])
        if rank < world - 1:
            assert plan["event"] is not None and plan["skip_last"]  # cut mid-event by construction
        if rank > 0:
            assert plan["skip_first"] and len(plan["send"]) == 1
        # ragged table all-gather: (global start, length) rows + a float column
        ints = torch.tensor([[offsets[rank] + s, n] for s, n in keep], dtype=torch.int64).reshape(-1, 2)
        flts = ints[:, :1].to(torch.float64) * 0.5
        counts = [None] * world
        dist.all_gather_object(counts, len(keep))
        gi, gf = ppdist.gather_tables(ints, flts, counts, dist)
        ws, wl = oracle.events(glob.astype(np.float64), 110, PYRULES)
        assert np.array_equal(gi[:, 0].numpy(), ws) and np.array_equal(gi[:, 1].numpy(), wl)
        assert np.array_equal(gf[:, 0].numpy(), ws * 0.5)
        # the packed single-collective gather used by ShardedPipeline (pp_pack_tables layout): events as above plus
        # segment rows {event id | start << 32, mean, std, min | max as float32}; ragged counts, one rank may be
        # longer; unpack_gathered rebuilds `end` from the next row / the event length
        k0 = sum(c for c in counts[:rank])
        seg_ev, seg_st, seg_en = [], [], []
        for j, (s0, n0) in enumerate(keep):
            cuts = [0, n0 // 3, n0 // 2 + rank, n0]
            for a, b in zip(cuts[:-1], cuts[1:]):
                seg_ev.append(k0 + j); seg_st.append(a); seg_en.append(b)
        nseg = len(seg_ev)
        mean = np.arange(nseg) * 0.25 - rank
        sd = np.arange(nseg) * 0.5 + 1
        mn = (np.arange(nseg) - 3.5).astype(np.float32)
        mx = (np.arange(nseg) + 0.125).astype(np.float32)
        if nseg:
            mn[0] = np.nan
        w0 = np.asarray(seg_ev, np.int64) | (np.asarray(seg_st, np.int64) << 32)
        w3 = mn.view(np.uint32).astype(np.int64) | (mx.view(np.uint32).astype(np.int64) << 32)
        rows = np.stack([w0, mean.view(np.int64), sd.view(np.int64), w3], axis=1) if nseg else np.zeros((0, 4), np.int64)
        words = torch.cat([ints.reshape(-1), torch.from_numpy(rows.reshape(-1)),
                           torch.zeros(5 * rank, dtype=torch.int64)])
        pcounts = [None] * world
        dist.all_gather_object(pcounts, (len(keep), nseg))
        m = max(max(2 * e + ppdist.SEG_WORDS * s for e, s in pcounts), 1)
        t = ppdist.unpack_gathered(ppdist.gather_packed_raw(words, m, dist), pcounts)
        assert np.array_equal(t["events"].numpy()[:, 0], ws) and np.array_equal(t["events"].numpy()[:, 1], wl)
        lo = sum(c[1] for c in pcounts[:rank])
        assert t["seg_int"].shape == (sum(c[1] for c in pcounts), 3) and t["seg_flt"].dtype == torch.float64
        mine_i = t["seg_int"][lo:lo + nseg].numpy()
        mine_f = t["seg_flt"][lo:lo + nseg].numpy()
        assert np.array_equal(mine_i[:, 0], seg_ev) and np.array_equal(mine_i[:, 1], seg_st)
        assert np.array_equal(mine_i[:, 2], seg_en)
        assert np.array_equal(mine_f[:, 0], mean) and np.array_equal(mine_f[:, 1], sd)
        assert np.array_equal(mine_f[:, 2], mn.astype(np.float64), equal_nan=True)
        assert np.array_equal(mine_f[:, 3], mx.astype(np.float64))
        # segments of the straddling event, split from chunk+halo, equal the oracle's on the global trace
        if plan["event"] is not None:
            s, n = plan["event"]
            a = oracle.statsplit(ext[s:s + n].astype(np.float64), prior_segments_per_second=10)
            b = oracle.statsplit(glob[offsets[rank] + s:offsets[rank] + s + n].astype(np.float64),
                                 prior_segments_per_second=10)
            assert np.array_equal(a, b)
        q.put((rank, "ok"))
    except Exception as e:  # surface the failure to the parent
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_event_detection_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 4, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", "rank %d: %s" % (rank, msg)


def _info(n, runs):
    """runs: list of (below, length, min, max) covering a chunk of n samples."""
    starts = np.cumsum([0] + [r[1] for r in runs])[:-1]
    first = (starts[0], runs[0][1], runs[0][2], runs[0][3], runs[0][0])
    last = (starts[-1], runs[-1][1], runs[-1][2], runs[-1][3], runs[-1][0])
    return ppdist.boundary_info(n, len(runs), first, last)


def test_plan_event_spanning_whole_chunks():
    # rank0 ends inside an event that covers ALL of rank1 and ends inside rank2
    infos = [_info(5000, [(False, 3000, 118, 125), (True, 2000, 30, 80)]),
             _info(4000, [(True, 4000, 25, 85)]),
             _info(6000, [(True, 1500, 40, 70), (False, 4500, 117, 126)])]
    plans = ppdist.plan_boundaries(infos, RULES)
    assert plans[0]["event"] == (3000, 2000 + 4000 + 1500) and plans[0]["recv"] == [(1, 4000), (2, 1500)]
    assert plans[0]["skip_last"] and not plans[0]["skip_first"]
    assert plans[1]["skip_first"] and plans[1]["skip_last"] and plans[1]["event"] is None
    assert plans[1]["send"] == [(0, 4000)] and plans[2]["send"] == [(0, 1500)]
    assert plans[2]["skip_first"] and not plans[2]["skip_last"]


def test_plan_rejected_and_non_straddling_runs():
    # open channel straddles 0|1 (rejected by max<thr: no halo moves); an edge sits exactly on 1|2
    infos = [_info(5000, [(True, 3000, 30, 80), (False, 2000, 118, 125)]),
             _info(4000, [(False, 1000, 117, 124), (True, 3000, 20, 85)]),
             _info(3000, [(False, 3000, 118, 126)])]
    plans = ppdist.plan_boundaries(infos, RULES)
    assert plans[0]["skip_last"] and plans[0]["event"] is None and plans[0]["recv"] == []
    assert plans[1]["skip_first"] and plans[1]["send"] == [] and not plans[1]["skip_last"]
    assert not plans[2]["skip_first"]
    # a straddling event with a sub-zero spike in its continuation is rejected as a whole (min > -0.5)
    infos = [_info(5000, [(False, 3000, 118, 125), (True, 2000, 30, 80)]),
             _info(4000, [(True, 900, -20, 85), (False, 3100, 117, 126)])]
    plans = ppdist.plan_boundaries(infos, RULES)
    assert plans[0]["event"] is None and plans[0]["skip_last"] and plans[1]["skip_first"]
    # too short on either side alone, long enough merged (duration > 1000)
    infos = [_info(5000, [(False, 4400, 118, 125), (True, 600, 30, 80)]),
             _info(4000, [(True, 600, 35, 85), (False, 3400, 117, 126)])]
    plans = ppdist.plan_boundaries(infos, RULES)
    assert plans[0]["event"] == (4400, 1200)
    # NaN in the continuation poisons min/max like np.min / np.max: rejected
    infos = [_info(5000, [(False, 3000, 118, 125), (True, 2000, 30, 80)]),
             _info(4000, [(True, 900, np.nan, np.nan), (False, 3100, 117, 126)])]
    assert ppdist.plan_boundaries(infos, RULES)[0]["event"] is None


def test_plan_boundaries_property_random_cuts():
    """Random traces cut at random places (inside events, exactly on run edges, several cuts inside one run,
    one-sample chunks): the events the ranks own under plan_boundaries -- their own accepted runs minus the
    skipped first / last run, plus the straddling event each rank keeps -- must be exactly the events of the
    uncut trace, each owned once, with recv counts that deliver exactly the continuation samples."""
    rng = np.random.RandomState(11)
    rules = dict(rule_mask=7, duration_gt=6, duration_lt=0, min_gt=-0.5, max_lt=110.0)
    pyrules = [lambda e: e.duration > 6, lambda e: e.min > -0.5, lambda e: e.max < 110]
    n_straddle = n_multi = 0
    for case in range(300):
        parts = []
        for _ in range(rng.randint(1, 9)):
            parts.append(rng.normal(120, 1.5, rng.randint(1, 12)))
            body = rng.normal(rng.uniform(20, 90), 1.0, rng.randint(1, 25))
            if rng.uniform() < 0.15:
                body[rng.randint(len(body))] = -20.0          # sub-zero spike: the whole run is rejected
            parts.append(body)
        if rng.uniform() < 0.5:
            parts.append(rng.normal(120, 1.5, rng.randint(1, 12)))
        if rng.uniform() < 0.3:
            parts = parts[1:]                                   # trace starts inside an event
        x = np.concatenate(parts)
        n = len(x)
        world = int(rng.randint(2, 7))
        if n < world:
            continue
        cuts = np.sort(rng.choice(np.arange(1, n), size=world - 1, replace=False))
        bounds = np.concatenate(([0], cuts, [n]))
        ws, wl = oracle.events(x, 110.0, pyrules)
        want = sorted(zip(ws.tolist(), wl.tolist()))
        infos, runs = [], []
        for r in range(world):
            chunk = x[bounds[r]:bounds[r + 1]]
            rr = oracle.threshold_runs(chunk, 110.0)
            runs.append(rr)
            infos.append(ppdist.boundary_info(len(chunk), len(rr[0]), [a[0] for a in rr], [a[-1] for a in rr]))
        plans = ppdist.plan_boundaries(infos, rules)
        got = []
        for r in range(world):
            rr, plan = runs[r], plans[r]
            k = len(rr[0])
            for i in range(k):
                if (i == 0 and plan["skip_first"]) or (i == k - 1 and plan["skip_last"]):
                    continue
                if ppdist.rules_accept(rules, rr[1][i], rr[2][i], rr[3][i]):
                    got.append((int(bounds[r] + rr[0][i]), int(rr[1][i])))
            if plan["event"] is not None:
                s, length = plan["event"]
                got.append((int(bounds[r] + s), int(length)))
                n_straddle += 1
                n_multi += len(plan["recv"]) > 1
                # the halo pieces are the heads of the following chunks, in order, and end where the event ends
                assert [q for q, _ in plan["recv"]] == list(range(r + 1, r + 1 + len(plan["recv"])))
                assert bounds[r] + s + length == bounds[plan["recv"][-1][0]] + plan["recv"][-1][1]
                for q, cnt in plan["recv"]:
                    assert (r, cnt) in plans[q]["send"] and cnt <= bounds[q + 1] - bounds[q]
        assert sorted(got) == want, "case %d world %d" % (case, world)
    assert n_straddle > 50 and n_multi > 5      # the generator does produce the interesting cases


def test_device_dealing_and_proportional_cuts():
    """Ranks of a job smaller than the node alternate between the two halves of the board; chunk boundaries follow
    the measured upload rates."""
    from pypore_b200.dist import device_for_rank, proportional_cuts
    assert [device_for_rank(r, 8) for r in range(8)] == [0, 4, 1, 5, 2, 6, 3, 7]
    assert [device_for_rank(r, 8) for r in range(2)] == [0, 4]
    assert device_for_rank(0, 1) == 0 and [device_for_rank(r, 3) for r in range(3)] == [0, 1, 2]
    assert sorted(device_for_rank(r, 4) for r in range(4)) == [0, 1, 2, 3]
    cuts = proportional_cuts(1000, [23.0, 23.0, 35.0, 35.0])
    assert cuts[0] == 0 and cuts[-1] == 1000 and (np.diff(cuts) > 0).all()
    assert abs(int(np.diff(cuts)[0]) - 198) <= 1 and abs(int(np.diff(cuts)[3]) - 302) <= 1
    assert list(proportional_cuts(12, [1, 1, 1])) == [0, 4, 8, 12]


def test_recut_chunks_partition_the_same_global_trace():
    """The chunks of a cut that follows upload rates are a partition of the very trace the even cut partitions."""
    world, epr = 3, 4
    chunks = [ppdist.synthetic_chunk(r, world, epr, seed0=5) for r in range(world)]
    whole = np.concatenate(chunks)
    assert np.array_equal(whole, ppdist.synthetic_global(world, epr, seed0=5))
    lens = [len(c) for c in chunks]
    for rates in ([1.0, 1.0, 1.0], [20.6, 36.0, 20.6], [1.0, 5.0, 0.2], [3.0, 0.01, 3.0]):
        cuts = ppdist.proportional_cuts(len(whole), rates)
        parts = [ppdist.recut_chunk(r, cuts, lens, chunks[r], lambda q: chunks[q]) for r in range(world)]
        assert [len(p) for p in parts] == list(np.diff(cuts))
        assert np.array_equal(np.concatenate(parts), whole)
