"""(development, GPU box) What the box gives for concurrent host-to-device copies: every rank copies a 240 MB
page-locked buffer to its GPU at the same time, with different kinds of page-locked memory.  Per-rank and
aggregate GB/s.   torchrun --nproc-per-node N scripts/h2d_probe.py"""
import ctypes
import glob
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist


def cudart():
    base = os.path.dirname(torch.__file__)
    cands = glob.glob(os.path.join(base, "lib", "libcudart*.so*")) + \
        glob.glob(os.path.join(os.path.dirname(base), "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
    return ctypes.CDLL(cands[0])


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from pypore_b200.dist import bind_near_gpu
    bind_near_gpu(local)
    rt = cudart()
    rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    nbytes = 240 << 20
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def alloc(flags):
        p = ctypes.c_void_p()
        assert rt.cudaHostAlloc(ctypes.byref(p), nbytes, flags) == 0
        ctypes.memset(p.value, 1, nbytes)
        return p.value

    kinds = {"default": alloc(0), "portable": alloc(1), "write-combined": alloc(4)}
    reps = 5
    for name, host in kinds.items():
        for streams in (1, 2):
            ms = []
            for _ in range(2):
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(reps):
                    if streams == 1:
                        rt.cudaMemcpyAsync(dev.data_ptr(), host, nbytes, 1, ctypes.c_void_p(s1.cuda_stream))
                    else:
                        h = nbytes // 2
                        rt.cudaMemcpyAsync(dev.data_ptr(), host, h, 1, ctypes.c_void_p(s1.cuda_stream))
                        rt.cudaMemcpyAsync(dev.data_ptr() + h, host + h, h, 1, ctypes.c_void_p(s2.cuda_stream))
                torch.cuda.synchronize()
                mine = (time.perf_counter() - t0) * 1e3 / reps
                dist.barrier()
                ms.append(mine)
            t = torch.tensor([ms[-1]], device="cuda", dtype=torch.float64)
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            if rank == 0:
                per = [float(a.item()) for a in allt]
                print("%-15s streams %d: per-rank ms %s  -> aggregate %.1f GB/s (slowest rank), sum of rates %.1f GB/s"
                      % (name, streams, " ".join("%.2f" % p for p in per), world * nbytes / max(per) / 1e6,
                         sum(nbytes / p / 1e6 for p in per)), flush=True)
    # one rank at a time, for the per-link rate
    for r in range(world):
        dist.barrier()
        if r == rank:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                rt.cudaMemcpyAsync(dev.data_ptr(), kinds["default"], nbytes, 1, ctypes.c_void_p(s1.cuda_stream))
            torch.cuda.synchronize()
            print("rank %d alone: %.1f GB/s" % (r, nbytes * reps / (time.perf_counter() - t0) / 1e9), flush=True)
        dist.barrier()
    # pairs (0,r): which GPUs share an uplink
    for r in range(1, world):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if rank in (0, r):
            for _ in range(reps):
                rt.cudaMemcpyAsync(dev.data_ptr(), kinds["default"], nbytes, 1, ctypes.c_void_p(s1.cuda_stream))
            torch.cuda.synchronize()
            if rank == 0:
                print("ranks 0+%d together: rank 0 gets %.1f GB/s" % (r, nbytes * reps / (time.perf_counter() - t0) / 1e9),
                      flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
