"""Dev (torchrun, 2 GPUs): does a tiny NCCL collective on one stream wait for a 240 MB pinned H2D copy on another?"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypore_b200 import _lib

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
xp = torch.empty(60_000_000, dtype=torch.float32).pin_memory().numpy()
A, B = _lib.Context(local), _lib.Context(local)
sa = torch.cuda.ExternalStream(A.stream_handle)
B.upload_trace(xp)
t = torch.ones(16, device="cuda")
out = torch.empty(16 * world, device="cuda")
g2 = dist.new_group()
for name, fn in (("all_reduce default pg", lambda: dist.all_reduce(t)),
                 ("all_gather new group", lambda: dist.all_gather_into_tensor(out, t, group=g2)),
                 ("plain kernel (t.add_)", lambda: t.add_(1.0))):
    with torch.cuda.stream(sa):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); sa.synchronize(); alone = (time.perf_counter() - t0) * 1e3
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        B.upload_trace_async(xp)
        fn()
        sa.synchronize()
        under = (time.perf_counter() - t0) * 1e3
        B.sync()
        both = (time.perf_counter() - t0) * 1e3
    if rank == 0:
        print("%-26s alone %.3f ms, under a 240 MB H2D on another stream %.3f ms (copy done at %.3f ms)" % (name, alone, under, both), flush=True)
dist.destroy_process_group()
