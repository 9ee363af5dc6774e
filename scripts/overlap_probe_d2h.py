"""Dev (1 GPU): is a tiny D2H copy / memset / H2D on one stream held up by a 240 MB pinned H2D copy on another?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypore_b200 import _lib
xp = torch.empty(60_000_000, dtype=torch.float32).pin_memory().numpy()
A, B = _lib.Context(0), _lib.Context(0)
sa = torch.cuda.ExternalStream(A.stream_handle)
B.upload_trace(xp)
t = torch.ones(64, device="cuda")
host = torch.empty(64).pin_memory()
small = torch.ones(64).pin_memory()
tests = {
    "D2H pageable (.cpu())": lambda: t.cpu(),
    "D2H pinned (copy_ non_blocking)": lambda: host.copy_(t, non_blocking=True),
    "H2D pinned small": lambda: t.copy_(small, non_blocking=True),
    "H2D pageable small (torch.tensor)": lambda: torch.tensor([1, 2, 3], device="cuda"),
    "memset (zero_)": lambda: t.zero_(),
    "kernel (add_)": lambda: t.add_(1.0),
}
with torch.cuda.stream(sa):
    for name, fn in tests.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); sa.synchronize(); alone = (time.perf_counter() - t0) * 1e3
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        B.upload_trace_async(xp)
        fn()
        sa.synchronize()
        under = (time.perf_counter() - t0) * 1e3
        B.sync()
        print("%-36s alone %.3f ms, under the 240 MB H2D %.3f ms" % (name, alone, under), flush=True)
