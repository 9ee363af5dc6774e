"""Multi-GPU parity on real devices: torchrun over NCCL, every rank checks the all-gathered event and
segment tables against the oracle run on the whole trace (straddling events at every chunk boundary)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_pipeline_matches_oracle(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "tests", "dist_worker.py"), "24"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


# ---------------------------------------------------------------------------------------------
# The device-planned step on ONE GPU: several contexts play the ranks, the collectives are plain
# device copies.  Covers k_shard_plan, the in-place halo and pp_shard_finish_planned against the
# oracle and against the host plan (pypore_b200.dist.plan_boundaries) without a multi-GPU box.
# ---------------------------------------------------------------------------------------------
RULES = dict(rule_mask=7, duration_gt=1000, duration_lt=0, min_gt=-0.5, max_lt=110.0)


def _random_infos(rng, world):
    """Boundary records with every shape the plan has to handle: joined / unjoined boundaries, chunks
    that are one single run, NaN extrema, runs that the rules reject."""
    import numpy as np
    from pypore_b200 import dist as ppdist
    infos = np.zeros((world, ppdist.INFO_LEN))
    for r in range(world):
        nruns = int(rng.choice([1, 1, 2, 3, 50]))
        n = int(rng.randint(2000, 30000))
        fb = int(rng.randint(0, 2))
        lb = fb if nruns % 2 == 1 else 1 - fb
        first_len = n if nruns == 1 else int(rng.randint(1, n // 2))
        last_len = n if nruns == 1 else int(rng.randint(1, n - first_len + 1))

        def ext(below):
            lo = rng.uniform(-2, 90) if below else rng.uniform(110, 115)
            v = [lo, lo + rng.uniform(0, 15)]
            if rng.uniform() < 0.1:
                v[int(rng.randint(0, 2))] = np.nan
            return v
        fmn, fmx = ext(fb)
        lmn, lmx = (fmn, fmx) if nruns == 1 else ext(lb)
        infos[r] = [n, nruns, fb, first_len, fmn, fmx, lb, n - last_len, last_len, lmn, lmx, 0]
    return infos


def test_device_plan_matches_host_plan(ctx):
    import numpy as np
    import torch
    from pypore_b200 import dist as ppdist
    rng = np.random.RandomState(5)
    plan_dev = torch.zeros(8, dtype=torch.int64, device="cuda")
    n_event = n_redo = 0
    for trial in range(300):
        world = int(rng.randint(1, 7))
        infos = _random_infos(rng, world)
        if trial % 50 == 49:
            infos[int(rng.randint(0, world)), ppdist.I_PAD] = 1.0
        halo = int(rng.choice([0, 500, 4000, 1 << 16]))
        host = ppdist.plan_boundaries(infos, RULES)
        dev_infos = torch.from_numpy(infos.reshape(-1)).cuda()
        for r in range(world):
            ctx.shard_plan(dev_infos.data_ptr(), r, world, 110.0, RULES, halo, plan_dev.data_ptr())
            ctx.sync()
            p = plan_dev.cpu().numpy()
            h = host[r]
            assert bool(p[0]) == h["skip_first"] and bool(p[1]) == h["skip_last"], (trial, r)
            need = sum(c for _, c in h["recv"])
            covered = h["event"] is not None and len(h["recv"]) == 1 and need <= halo
            assert bool(p[5] & 1) == bool(infos[:, ppdist.I_PAD].any())
            if h["event"] is None:
                assert p[2] == 0 and not (p[5] & 16)
            elif covered:
                assert p[2] == 1 and (int(p[3]), int(p[4])) == h["event"] and p[6] == need and not (p[5] & 16)
                n_event += 1
            else:
                assert p[2] == 0 and (p[5] & 16) and p[6] == need
                n_redo += 1
    assert n_event > 20 and n_redo > 20


@pytest.mark.parametrize("world,halo_spec", [(2, 1 << 16), (3, 1 << 16), (3, 3000)])
def test_device_planned_step_on_one_gpu(world, halo_spec):
    import numpy as np
    import torch
    import oracle
    from pypore_b200 import _lib, dist as ppdist
    from pypore_b200.parsers import statsplit_min_gain
    epr, seed0 = 12, 40
    chunks = [ppdist.synthetic_chunk(r, world, epr, seed0=seed0) for r in range(world)]
    lens = [len(c) for c in chunks]
    offsets = np.concatenate(([0], np.cumsum(lens)))
    ctxs = [_lib.Context(0) for _ in range(world)]
    try:
        mw, MW, W, gain = statsplit_min_gain(min_width=100, max_width=1000000, window_width=10000)
        recs = torch.zeros(world, ppdist.INFO_LEN, dtype=torch.float64, device="cuda")
        plans = torch.zeros(world, 8, dtype=torch.int64, device="cuda")
        res = torch.zeros(world, 8, dtype=torch.int64, device="cuda")
        for r, c in enumerate(ctxs):
            c.upload_trace(chunks[r], extra_capacity=1 << 17)
            c.shard_scan(110.0, lens[r], recs[r].data_ptr())
            c.sync()
        halos = []
        for r, c in enumerate(ctxs):
            h = min(halo_spec, lens[r + 1]) if r < world - 1 else 0
            if h:   # what the NCCL recv does: the neighbour's head lands right after this rank's chunk
                room = ppdist.device_view(c.trace_ptr + 4 * lens[r], h, torch.float32, torch.device("cuda", 0))
                room.copy_(ppdist.device_view(ctxs[r + 1].trace_ptr, h, torch.float32, torch.device("cuda", 0)))
            halos.append(h)
        torch.cuda.synchronize()
        for r, c in enumerate(ctxs):
            c.extend_trace(halos[r])
            c.shard_plan(recs.data_ptr(), r, world, 110.0, RULES, halos[r], plans[r].data_ptr())
            c.shard_finish_planned(110.0, RULES, mw, MW, W, gain, plans[r].data_ptr(), res[r].data_ptr())
            c.sync()
        allr = res.cpu().numpy()
        host = ppdist.plan_boundaries(recs.cpu().numpy(), RULES)
        if halo_spec < 5000:
            # the straddling events need up to ~7500 continuation samples: every owner must ask for the redo
            assert all(allr[r, 4] & 16 for r in range(world - 1)) and not (allr[world - 1, 4] & 16)
            return
        assert not (allr[:, 4] & 17).any()
        glob = ppdist.synthetic_global(world, epr, seed0=seed0).astype(np.float64)
        pyrules = [lambda e: e.duration > 1000, lambda e: e.min > -0.5, lambda e: e.max < 110]
        ws, wl = oracle.events(glob, 110, pyrules)
        oe, ost, oen, _ = oracle.statsplit_events(glob, ws, wl, gain=gain, threads=4)
        ev_s, ev_l, sg_e, sg_s, sg_n, base = [], [], [], [], [], 0
        for r, c in enumerate(ctxs):
            assert host[r]["event"] is None or plans[r, 2].item() == 1
            c.shard_commit(allr[r])
            es, el = c.events(int(allr[r, 1]))
            sg = c.segments(int(allr[r, 3]))
            ev_s.append(es + offsets[r]); ev_l.append(el)
            sg_e.append(sg["event"].astype(np.int64) + base); sg_s.append(sg["start"].copy()); sg_n.append(sg["end"].copy())
            base += len(es)
            # statistics of the straddling event's segments were computed over chunk + halo
            if len(es):
                e = len(es) - 1
                sel = sg["event"] == e
                a = int(offsets[r] + es[e])
                m, s, mn, mx = oracle.segment_stats(glob[a:a + el[e]], sg["start"][sel], sg["end"][sel])
                assert np.allclose(sg["mean"][sel], m, rtol=1e-9, atol=0) and np.allclose(sg["std"][sel], s, rtol=1e-9, atol=0)
                assert np.array_equal(sg["min"][sel], mn) and np.array_equal(sg["max"][sel], mx)
        assert np.array_equal(np.concatenate(ev_s), ws) and np.array_equal(np.concatenate(ev_l), wl)
        assert np.array_equal(np.concatenate(sg_e), oe) and np.array_equal(np.concatenate(sg_s), ost)
        assert np.array_equal(np.concatenate(sg_n), oen)
    finally:
        for c in ctxs:
            c.close()


def test_peer_control_exchange_on_one_gpu():
    """pp_ctl_exchange (peer stores + flags instead of an all-gather of the control records): three contexts of one
    process map each other's record buffers by pointer and exchange records of both kinds over many rounds, every
    context on its own stream; what each receives must be the stacked records of that round."""
    import numpy as np
    import torch
    from pypore_b200 import _lib
    world = 3
    ctxs = [_lib.Context(0) for _ in range(world)]
    try:
        ptrs = [c.ctl_create(r, world)[1] for r, c in enumerate(ctxs)]
        for c in ctxs:
            c.ctl_open(local_ptrs=ptrs)
        rng = np.random.RandomState(2)
        for rnd in range(40):
            n_words = 12 if rnd % 2 == 0 else 8
            recs = rng.randint(-2**62, 2**62, size=(world, n_words)).astype(np.int64)
            src = [torch.from_numpy(recs[r]).cuda() for r in range(world)]
            dst = [torch.zeros(world * n_words, dtype=torch.int64, device="cuda") for _ in range(world)]
            torch.cuda.synchronize()
            for r, c in enumerate(ctxs):                      # all three enqueued before anybody waits
                c.ctl_exchange(src[r].data_ptr(), n_words, dst[r].data_ptr())
            for c in ctxs:
                c.sync()
            for r in range(world):
                assert np.array_equal(dst[r].cpu().numpy().reshape(world, n_words), recs), (rnd, r)
    finally:
        for c in ctxs:
            c.close()
