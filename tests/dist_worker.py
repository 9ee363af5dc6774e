"""Run under torchrun (one rank per GPU) by tests/test_gpu_dist.py: sharded pipeline vs the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from pypore_b200 import _lib, dist as ppdist  # noqa: E402
from pypore_b200.parsers import statsplit_min_gain  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    epr = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.Context(local)
    chunk = ppdist.synthetic_chunk(rank, world, epr, seed0=70)
    shard = ppdist.ShardedPipeline(ctx, rank, world)
    pinned = torch.from_numpy(chunk).pin_memory()
    shard.load(pinned.numpy())
    rules = dict(rule_mask=7, duration_gt=1000, duration_lt=0, min_gt=-0.5, max_lt=110.0)
    for kw in (dict(min_width=100, max_width=1000000, window_width=10000),
               dict(min_width=100, max_width=1000000, window_width=10000, prior_segments_per_second=10)):
        mw, MW, W, gain = statsplit_min_gain(**kw)
        for _ in range(2):  # twice: buffers, counters and the asynchronous table gather are reused across steps
            shard.step(110.0, rules, mw, MW, W, gain)
        tabs = ppdist.rows(shard.download())   # (stacked copies: the column views live in a pinned arena that is reused)
        assert shard.fallbacks == 0
        # the one-kernel unpack into page-locked host tables against the tensor path on the gathered buffer
        eager = {k: v.cpu().numpy() for k, v in shard.tables.items()}
        assert all(np.array_equal(tabs[k], eager[k], equal_nan=True) for k in tabs), "pp_unpack_tables differs"
        # the host-planned step (the fallback of the device-planned one) must give the same tables
        shard.step(110.0, rules, mw, MW, W, gain, host_planned=True)
        alt = ppdist.rows(shard.download())
        assert all(np.array_equal(tabs[k], alt[k], equal_nan=True) for k in tabs), "host-planned step differs"
        shard.step(110.0, rules, mw, MW, W, gain)   # and back: the speculative halo is still in place
        alt = ppdist.rows(shard.download())
        assert all(np.array_equal(tabs[k], alt[k], equal_nan=True) for k in tabs), "device-planned step after it differs"
        # every rank's own rows (what each process copies out end to end) are its slice of the whole table
        own = ppdist.rows(shard.download_async(own_rows=True).wait())
        e0 = sum(c[0] for c in shard.counts[:rank]); s0 = sum(c[1] for c in shard.counts[:rank])
        ne, ns = shard.counts[rank]
        for k in tabs:
            lo, n = (e0, ne) if k == "events" else (s0, ns)
            assert np.array_equal(own[k], tabs[k][lo:lo + n], equal_nan=True), "own rows differ: " + k
        # steps without their closing host synchronisation: the second is enqueued before the first is finished
        t1 = shard.step_async(110.0, rules, mw, MW, W, gain)
        t2 = shard.step_async(110.0, rules, mw, MW, W, gain)
        r1, r2 = shard.finish(t1), shard.finish(t2)
        assert r1 == r2 and shard.fallbacks == 0
        alt = ppdist.rows(shard.download())
        assert all(np.array_equal(tabs[k], alt[k], equal_nan=True) for k in tabs), "step_async differs"
        glob = ppdist.synthetic_global(world, epr, seed0=70).astype(np.float64)
        pyrules = [lambda e: e.duration > 1000, lambda e: e.min > -0.5, lambda e: e.max < 110]
        ws, wl = oracle.events(glob, 110, pyrules)
        ev = tabs["events"]
        assert np.array_equal(ev[:, 0], ws) and np.array_equal(ev[:, 1], wl), "events differ on rank %d" % rank
        oe, ost, oen, _ = oracle.statsplit_events(glob, ws, wl, gain=gain, threads=4)
        si, sf = tabs["seg_int"], tabs["seg_flt"]
        assert np.array_equal(si[:, 0], oe) and np.array_equal(si[:, 1], ost) and np.array_equal(si[:, 2], oen)
        for e in range(0, len(ws), max(1, len(ws) // 12)):
            sel = oe == e
            m, s, mn, mx = oracle.segment_stats(glob[ws[e]:ws[e] + wl[e]], ost[sel], oen[sel])
            assert np.allclose(sf[sel, 0], m, rtol=1e-9, atol=0) and np.allclose(sf[sel, 1], s, rtol=1e-9, atol=0)
            assert np.array_equal(sf[sel, 2], mn) and np.array_equal(sf[sel, 3], mx)
    dist.barrier()
    if rank == 0:
        print("DIST OK world=%d events=%d segments=%d" % (world, len(ws), len(oe)), flush=True)
    dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
