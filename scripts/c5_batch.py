"""BASELINE configs[4] on this box: a batch of 250 kHz files, Event.filter(1, 2000) + SpeedyStatSplit per event,
through pypore_b200.batch.FileBatch with 1, 2, 4, 7 worker contexts per GPU (split-search wave shared between the contexts or not) (host wall clock around parse_local;
every file pass ends with its own table read-back, so the clock covers H2D, kernels and D2H of every file).

    python scripts/c5_batch.py [file_passes=128] [events_per_file=208]      # 1 GPU
    torchrun --nproc-per-node N scripts/c5_batch.py ...                      # files dealt over N ranks

Distinct synthetic files are generated once (8 of them, pinned) and reused cyclically: the timed work per pass
is identical to a batch of distinct files, only the generation time is saved."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pypore_b200 import _lib, synth  # noqa: E402
from pypore_b200.batch import FileBatch  # noqa: E402
from pypore_b200.parsers import RuleSet, SpeedyStatSplit, lambda_event_parser  # noqa: E402

FS = 2.5e5


def main():
    passes = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    epf = int(sys.argv[2]) if len(sys.argv) > 2 else 208
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    host = _lib.Context(local)
    distinct = []
    for i in range(8):
        x = synth.make_trace(epf, seed=900 + i, tier="A")
        p = host.pinned_empty(x.shape[0], np.float32)
        p[:] = x
        distinct.append(p)
    files = [distinct[i % len(distinct)] for i in range(passes * world)]
    n_samples = sum(f.shape[0] for f in files)
    det = lambda_event_parser(threshold=110, rules=RuleSet(duration_gt=1000, min_gt=-0.5, max_lt=110))
    settings = [("filter(1,2000)+psps=10", SpeedyStatSplit(min_width=100, window_width=10000, sampling_freq=FS,
                                                            cutoff_freq=2000., prior_segments_per_second=10), (1, 2000.)),
                ("filter(1,2000)+default gain", SpeedyStatSplit(min_width=100, window_width=10000), (1, 2000.))]
    lines = []
    G = 1 << 26
    # (worker contexts, split-search wave shared, samples per resident pass: 0 = one file per pass)
    combos = ((1, True, 0), (4, True, 0), (1, False, G), (2, False, G), (4, False, G))
    if len(sys.argv) > 3 and sys.argv[3] == "quick":
        combos = ((4, True, 0), (2, False, G))
    for name, seg, filt in settings:
        for workers, share, gs in combos:
            b = FileBatch(device=local, workers=workers, rank=rank, world=world, share_split=share, group_samples=gs)
            b.parse_local(files, 1000. / FS, det, seg, filt)     # every context's buffers grown, kernels loaded
            t_local, t_all = [], []
            for _ in range(3):
                if dist:
                    dist.barrier()
                t0 = time.perf_counter()
                b.parse_local(files, 1000. / FS, det, seg, filt)
                t_local.append(time.perf_counter() - t0)
            for _ in range(3):
                if dist:
                    dist.barrier()
                t0 = time.perf_counter()
                tables = b.parse(files, 1000. / FS, det, seg, filt)                       # + gather (world > 1)
                t_all.append(time.perf_counter() - t0)
            t_local, t_all = sorted(t_local)[1], sorted(t_all)[1]                         # medians of three
            if dist:
                import torch
                tt = torch.tensor([t_local, t_all], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t_local, t_all = tt.tolist()
            b.close()
            line = dict(config="C5", setting=name, n_gpus=world, workers_per_gpu=workers, split_ctas_shared=share,
                        samples_per_pass=gs, passes=len(b.groups),
                        files=len(files), samples=n_samples, events=tables.n_events, segments=tables.n_segments,
                        ms_per_file=1e3 * t_local / len(files), files_per_s=len(files) / t_local,
                        msamples_per_s=n_samples / t_local / 1e6,
                        msamples_per_s_with_gather=n_samples / t_all / 1e6,
                        clock="host wall clock, median of 3 batches after one warm-up batch, max over ranks")
            if rank == 0:
                print(json.dumps(line), flush=True)
            lines.append(line)
    if rank == 0:
        # where one file's time goes on ONE context (host wall clock per call; every call ends with a stream sync)
        from pypore_b200.DataTypes import bessel_coefficients
        name, seg, filt = settings[0]
        mw, MW, W, gain = seg._params()
        ba = bessel_coefficients(filt[0], filt[1], FS)
        rs = det._device_rules().device_args()
        acc = dict(pipeline=0.0, events=0.0, event_stats=0.0, segments=0.0, gpu_stage_sum=0.0)
        reps = 24
        for k in range(reps + 4):
            x = files[k % len(files)]
            t0 = time.perf_counter()
            c = host.pipeline(det.threshold, min_width=mw, max_width=MW, window_width=W, min_gain=gain, filter_ba=ba,
                              with_stats=True, host_trace=x, **rs)
            t1 = time.perf_counter()
            host.events(c["events"])
            t2 = time.perf_counter()
            host.event_stats(c["events"])
            t3 = time.perf_counter()
            host.segments(c["segments"])
            t4 = time.perf_counter()
            if k >= 4:
                acc["pipeline"] += t1 - t0; acc["events"] += t2 - t1; acc["event_stats"] += t3 - t2
                acc["segments"] += t4 - t3; acc["gpu_stage_sum"] += sum(host.stage_ms().values()) / 1e3
        line = dict(config="C5 one-context breakdown, ms per file", setting=name,
                    **{k: round(1e3 * v / reps, 4) for k, v in acc.items()})
        print(json.dumps(line), flush=True)
        lines.append(line)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "c5_batch_%dgpu.jsonl" % world), "w") as f:
            for line in lines:
                f.write(json.dumps(line) + "\n")
    host.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
