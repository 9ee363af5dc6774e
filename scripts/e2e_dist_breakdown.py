"""Dev (torchrun, one rank per GPU): where the sharded end-to-end step spends its time -- load / step / download,
host wall clock with a device synchronisation after each part, median of 8 steps, per rank."""
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
    ctx = _lib.Context(local)
    x = ppdist.synthetic_chunk(rank, world, 5000, seed0=1)
    xp = torch.from_numpy(x).pin_memory().numpy()
    shard = ppdist.ShardedPipeline(ctx, rank, world)
    rules = dict(rule_mask=7, duration_gt=1000, duration_lt=0, min_gt=-0.5, max_lt=110.0)
    mw, MW, W, gain = statsplit_min_gain(min_width=100, max_width=1000000, window_width=10000)
    t = {"load": [], "step": [], "download": [], "total": []}
    for it in range(11):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        shard.load(xp)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        shard.step(110.0, rules, mw, MW, W, gain)
        shard.wait()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if rank == 0:
            shard.download()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        if it >= 3:
            for k, v in zip(("load", "step", "download", "total"), (t1 - t0, t2 - t1, t3 - t2, t3 - t0)):
                t[k].append(v * 1e3)
    # pipelined: the copy-out of step i overlaps the upload of step i + 1 (download_async, what bench.py times)
    pend = None
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(8):
        shard.load(xp)
        shard.step(110.0, rules, mw, MW, W, gain)
        if rank == 0:
            if pend is not None:
                pend.wait()
            pend = shard.download_async()
    if pend is not None:
        pend.wait()
    shard.wait()
    torch.cuda.synchronize()
    t["pipelined"] = [(time.perf_counter() - t0) * 1e3 / 8]
    line = "rank %d  " % rank + "  ".join("%s %.2f ms" % (k, float(np.median(v))) for k, v in t.items())
    for r in range(world):
        dist.barrier()
        if r == rank:
            print(line, flush=True)
    dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
