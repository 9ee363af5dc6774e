"""Dev (torchrun): device-side timeline of the two-pipeline end-to-end loop of bench.py (rank 0 prints)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypore_b200 import _lib, dist as ppdist  # noqa: E402
from pypore_b200.parsers import statsplit_min_gain  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    x = ppdist.synthetic_chunk(rank, world, 5000, seed0=1)
    xp = torch.from_numpy(x).pin_memory().numpy()
    pipes = [ppdist.ShardedPipeline(_lib.Context(local), rank, world) for _ in range(2)]
    rules = dict(rule_mask=7, duration_gt=1000, duration_lt=0, min_gt=-0.5, max_lt=110.0)
    mw, MW, W, gain = statsplit_min_gain(min_width=100, max_width=1000000, window_width=10000)
    marks = []

    def mark(name, stream):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        marks.append((name, e, time.perf_counter()))

    pend = None
    for p in pipes:          # warm both
        p.load(xp)
        p.step(110.0, rules, mw, MW, W, gain)
        if rank == 0:
            p.download_async().wait()
    torch.cuda.synchronize()
    dist.barrier()
    origin = torch.cuda.Event(enable_timing=True)
    origin.record(pipes[0].stream)
    t_origin = time.perf_counter()
    pipes[0].load(xp)
    for k in range(6):
        cur, nxt = pipes[k & 1], pipes[(k + 1) & 1]
        mark("it%d nxt.load begin (host)" % k, nxt.stream)
        if os.environ.get("NOCTL"):
            nxt.ctx.upload_trace_async(xp, extra_capacity=nxt.HALO_CAPACITY)
            nxt._halo_pending = not os.environ.get("NOHALO")
        else:
            nxt.load(xp, defer_halo=True)
        mark("it%d nxt.load done on its stream" % k, nxt.stream)
        mark("it%d cur.step begin" % k, cur.stream)
        cur.step(110.0, rules, mw, MW, W, gain)
        mark("it%d cur.step kernels done" % k, cur.stream)
        if rank == 0 and not os.environ.get("NODL"):
            if pend is not None:
                pend.wait()
            pend = cur.download_async()
            mark("it%d unpack done (side)" % k, cur._side)
    torch.cuda.synchronize()
    if rank == 0:
        for name, e, th in marks:
            print("%-40s device %8.3f ms   host %8.3f ms" % (name, origin.elapsed_time(e), (th - t_origin) * 1e3), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
