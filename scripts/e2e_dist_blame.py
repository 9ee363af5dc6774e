"""Dev (torchrun): which call of the sharded step waits for the OTHER context's 240 MB upload?  Every library /
collective call of the step is followed by a synchronisation of the step's own stream and its host time is logged."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypore_b200 import _lib, dist as ppdist
from pypore_b200.parsers import statsplit_min_gain

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = ppdist.synthetic_chunk(rank, world, 5000, seed0=1)
xp = torch.from_numpy(x).pin_memory().numpy()
pipes = [ppdist.ShardedPipeline(_lib.Context(local), rank, world) for _ in range(2)]
rules = dict(rule_mask=7, duration_gt=1000, duration_lt=0, min_gt=-0.5, max_lt=110.0)
mw, MW, W, gain = statsplit_min_gain(min_width=100, max_width=1000000, window_width=10000)
for p in pipes:
    p.load(xp)
    p.step(110.0, rules, mw, MW, W, gain)
    p.step(110.0, rules, mw, MW, W, gain)
torch.cuda.synchronize()
log = []
cur, nxt = pipes
T0 = [0.0]

def wrap(obj, name, label, stream):
    f = getattr(obj, name)
    def g(*a, **k):
        r = f(*a, **k)
        stream.synchronize()
        log.append((label, (time.perf_counter() - T0[0]) * 1e3))
        return r
    setattr(obj, name, g)

for n in ("truncate_trace", "shard_scan", "extend_trace", "shard_plan", "shard_finish_planned", "pack_tables", "shard_commit", "ctl_exchange"):
    wrap(cur.ctx, n, "ctx." + n, cur.stream)
real_ag = dist.all_gather_into_tensor
def ag(*a, **k):
    r = real_ag(*a, **k)
    if not k.get("async_op"):
        cur.stream.synchronize()
    log.append(("all_gather_into_tensor%s" % (" (async)" if k.get("async_op") else ""), (time.perf_counter() - T0[0]) * 1e3))
    return r
cur.dist = type("D", (), {})()
for n in dir(dist):
    if not n.startswith("__"):
        try:
            setattr(cur.dist, n, getattr(dist, n))
        except Exception:
            pass
cur.dist.all_gather_into_tensor = ag
dist.barrier(); torch.cuda.synchronize()
T0[0] = time.perf_counter()
if not os.environ.get("NOUPLOAD"):
    nxt.ctx.upload_trace_async(xp, extra_capacity=nxt.HALO_CAPACITY)
log.append(("other context: upload enqueued", (time.perf_counter() - T0[0]) * 1e3))
cur.step(110.0, rules, mw, MW, W, gain)
log.append(("step returned", (time.perf_counter() - T0[0]) * 1e3))
nxt.ctx.sync()
log.append(("other context: upload done", (time.perf_counter() - T0[0]) * 1e3))
if rank == 0:
    for l, t in log:
        print("%8.3f ms  %s" % (t, l), flush=True)
dist.barrier()
dist.destroy_process_group()
