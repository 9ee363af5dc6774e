#!/usr/bin/env python
"""bench.py -- headline benchmark: Msamples/s segmented (threshold + SpeedyStatSplit).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass of the hot path (lambda_event_parser threshold scan + rule
selection -> SpeedyStatSplit prefix sums / split search -> segment table ->
segment statistics) over one synthetic trace.  At N=1 the workload is
BASELINE.json configs[1]: a 10 min 100 kHz trace (60 M float32 samples, 5000
events).  At N>1 every rank holds one such piece of a single N*60 M-sample
trace (weak scaling) cut mid-event, so each step also runs the boundary halo
exchange and the all-gather of the compact tables.

`--impl reference` times the reference's own CPU code for the same path: the
unmodified cparsers.pyx compiled into oracle/_ref (falling back to the C port)
plus the NumPy statements of parsers.py:148-155 / core.py:209-223, on all host
cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

THRESHOLD = 110.0
RULES = dict(duration_gt=1000, min_gt=-0.5, max_lt=110.0)
SPLIT = dict(min_width=100, max_width=1000000, window_width=10000)
EVENTS_PER_GPU = 5000
METRIC = "Msamples/s segmented (threshold+SpeedyStatSplit)"


def workload_config(n_gpus, events_per_gpu):
    return {
        "workload": "BASELINE configs[1]: synthetic 100 kHz float32 trace, %d events (~%d M samples) per GPU, "
                    "ADC-quantised (tier A); lambda_event_parser(threshold=110, rules=[duration>1000, min>-0.5, "
                    "max<110]) + SpeedyStatSplit(min_width=100, window_width=10000) (min_gain=-0.0) + segment stats"
                    % (events_per_gpu, round(events_per_gpu * 12e3 / 1e6)),
        "events_per_gpu": events_per_gpu,
        "partition": "one contiguous chunk per GPU, cut mid-event, halo + table all-gather" if n_gpus > 1
                     else "single GPU",
        "l2": "inputs (240 MB trace + 640 MB prefix sums per GPU) exceed the 126 MB L2; no explicit flush",
    }


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler(object):
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _loop(self):
        nv = self._nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def start(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------
# CPU reference path (timed as the baseline; never on the product path)
# --------------------------------------------------------------------------------------
def _cpu_threshold(x64, threshold):
    """lambda_event_parser.parse + _lambda_select, statement by statement (parsers.py:136-155)."""
    mask = np.where(x64 < threshold, 1, 0)
    mask = np.abs(np.diff(mask))
    tics = np.concatenate(([0], np.where(mask == 1)[0] + 1, [x64.shape[0]]))
    del mask
    pieces = [(tics[i], np.array(c)) for i, c in enumerate(np.split(x64, tics[1:-1]))]
    keep = []
    for start, cur in pieces:
        ok = np.all([cur.shape[0] > RULES["duration_gt"], np.min(cur) > RULES["min_gt"],
                     np.max(cur) < RULES["max_lt"]])
        if ok:
            keep.append((int(start), cur))
    return keep


_W = {}


def _cpu_split_events(events):
    """FastStatSplit.parse per event + one read of mean/std/min/max per segment."""
    import oracle
    n_seg = 0
    acc = 0.0
    if _W.get("ref") is None:
        try:
            _W["ref"] = oracle.load_ref_cparsers() if oracle.ref_available() else False
        except Exception:
            _W["ref"] = False
    ref = _W["ref"]
    for cur in events:
        if ref:
            segs = ref.FastStatSplit(SPLIT["min_width"], SPLIT["max_width"], SPLIT["window_width"]).parse(cur)
            for s in segs:
                c = s.current
                acc += np.mean(c) + np.std(c) + np.min(c) + np.max(c)
            n_seg += len(segs)
        else:
            bp = oracle.statsplit(cur, **SPLIT)
            edges = np.concatenate(([0], bp, [len(cur)]))
            for a, b in zip(edges[:-1], edges[1:]):
                c = cur[a:b]
                acc += np.mean(c) + np.std(c) + np.min(c) + np.max(c)
            n_seg += len(edges) - 1
    return n_seg, acc


def cpu_reference_run(x64, cores):
    """One pass of the reference path over x64 on `cores` processes.  Returns (seconds, events, segments, kind)."""
    import oracle
    kind = "reference" if oracle.ref_available() else "port"
    t0 = time.perf_counter()
    events = [cur for _, cur in _cpu_threshold(x64, THRESHOLD)]
    if cores <= 1:
        n_seg, _ = _cpu_split_events(events)
    else:
        import multiprocessing as mp
        chunks = [events[i::cores * 4] for i in range(cores * 4)]
        with mp.get_context("fork").Pool(cores) as pool:
            n_seg = sum(r[0] for r in pool.map(_cpu_split_events, chunks))
    return time.perf_counter() - t0, len(events), n_seg, kind


def run_reference_arm(args):
    from pypore_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    cores = os.cpu_count() or 1
    n_events = min(EVENTS_PER_GPU, 150 * cores)  # bounded sample: about a second per step
    x64 = synth.make_trace(n_events, seed=1, tier="A").astype(np.float64)
    for _ in range(args.warmup):
        cpu_reference_run(x64[:len(x64) // 8], cores)
    t = 0.0
    for _ in range(args.steps):
        dt, ne, ns, kind = cpu_reference_run(x64, cores)
        t += dt
    value = len(x64) * args.steps / t / 1e6
    sample = "%d-event prefix of the workload trace (%d samples) per step, %d processes over events" % (
        n_events, len(x64), cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, EVENTS_PER_GPU),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "events": ne, "segments": ns,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def parity_against_fixture(res_e2e):
    """The tables the end-to-end call just delivered to host memory against tests/golden/c2_full.npz -- counts and
    SHA-256 of the event rows (start, length) and segment rows (event, start, end) the REAL reference produced on this
    exact workload (tests/golden/make_golden.py --c2-only).  Outside every timed region; never raises."""
    try:
        import hashlib
        g = np.load(os.path.join(ROOT, "tests", "golden", "c2_full.npz"), allow_pickle=False)

        def sha(a):
            return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
        es, el = res_e2e["event_table"]
        t = res_e2e["segment_table"]
        ev = np.stack([np.asarray(es, np.int64), np.asarray(el, np.int64)], axis=1)
        rows = np.stack([np.asarray(t["event"]).astype(np.int64), np.asarray(t["start"], np.int64),
                         np.asarray(t["end"], np.int64)], axis=1)
        return {"fixture": "tests/golden/c2_full.npz (real reference, full size)",
                "events": int(len(ev)), "events_expected": int(g["events"]),
                "segments": int(len(rows)), "segments_expected": int(g["default_segments"]),
                "events_bit_exact": bool(sha(ev) == str(g["events_sha"])),
                "segments_bit_exact": bool(sha(rows) == str(g["default_sha"]))}
    except Exception as exc:  # a reporting extra must not cost the bench line
        return {"error": "%s: %s" % (type(exc).__name__, exc)}


def parity_sharded(tables, world):
    """The gathered tables rank 0 downloaded in the last end-to-end step (events {global start, length}, seg_int
    {global event id, start, end}) against tests/golden/sharded_full.npz: the CPU oracle on the uncut world x 60 M-sample
    trace (tests/golden/make_sharded_full.py).  Outside every timed region; never raises."""
    try:
        import hashlib
        g = np.load(os.path.join(ROOT, "tests", "golden", "sharded_full.npz"), allow_pickle=False)
        k = "w%d_" % world
        if k + "events" not in g.files:
            return {"error": "no fixture for %d GPUs" % world}

        def sha(a):
            return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
        ev = np.stack([np.asarray(tables["ev_start"], np.int64), np.asarray(tables["ev_len"], np.int64)], axis=1)
        rows = np.stack([np.asarray(tables[k], np.int64) for k in ("seg_event", "seg_start", "seg_end")], axis=1)
        return {"fixture": "tests/golden/sharded_full.npz (CPU oracle on the uncut trace, full size)",
                "events": int(len(ev)), "events_expected": int(g[k + "events"]),
                "segments": int(len(rows)), "segments_expected": int(g[k + "segments"]),
                "events_bit_exact": bool(sha(ev) == str(g[k + "events_sha"])),
                "segments_bit_exact": bool(sha(rows) == str(g[k + "segments_sha"]))}
    except Exception as exc:  # a reporting extra must not cost the bench line
        return {"error": "%s: %s" % (type(exc).__name__, exc)}


def ncu_traffic(kernel="k3_split"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full`
    capture of this same command (profiles/*_ncu_summary.csv, newest round); None without one."""
    import csv
    import glob
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_summary.csv")), reverse=True):
        try:
            rows = list(csv.reader(open(path)))
            col = next(i for i, name in enumerate(rows[0]) if name.startswith(kernel))
            tot = 0.0
            for r in rows:
                if r and r[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(r[col]) * unit[r[1]]
            if tot > 0:
                return tot, os.path.basename(path)
        except Exception:
            continue
    return None, None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pypore_b200 import _lib, synth
    from pypore_b200.parsers import statsplit_min_gain

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mw, MW, W, gain = statsplit_min_gain(**SPLIT)
    ctx = _lib.Context(local)
    if args.split_kernel is not None:
        ctx.set_split_kernel(args.split_kernel == "flow")
    # every kernel, copy-stream join and collective of a step is ordered on the context's stream:
    # the timing events are recorded there
    stream = torch.cuda.ExternalStream(ctx.stream_handle, device=torch.device("cuda", local))

    # ---- synthetic input ---------------------------------------------------------------
    epg = args.events_per_gpu
    if world == 1:
        x = synth.make_trace(epg, seed=1, tier="A")
        shard = None
    else:
        from pypore_b200 import dist as ppdist
        x = ppdist.synthetic_chunk(rank, world, epg, seed0=1)
        shard = ppdist.ShardedPipeline(ctx, rank, world)
    n_local = len(x)
    pinned = torch.from_numpy(x).pin_memory()
    xp = pinned.numpy()
    rules = dict(rule_mask=7, duration_gt=RULES["duration_gt"], duration_lt=0, min_gt=RULES["min_gt"],
                 max_lt=RULES["max_lt"])

    def step_resident():
        if shard is None:
            return ctx.pipeline(THRESHOLD, min_width=mw, max_width=MW, window_width=W, min_gain=gain,
                                with_stats=True, **rules)
        return shard.step(THRESHOLD, rules, mw, MW, W, gain)

    last_download = {}
    # Multi-GPU end to end: the upload of step i+1 (the context's copy stream, into a second device buffer) runs
    # under step i's kernels and collectives, and the copy-out of step i's tables (a side stream, the other PCIe
    # direction) under both.  Every step's H2D and D2H are inside the timed region.
    e2e = {"remaining": 0, "loaded": False}

    def step_e2e():
        if shard is None:
            # public host-memory entry point (pp_pipeline_host_tables): chunked H2D copy overlapped with the stages;
            # the event / segment rows a chunk has finalised are written into page-locked host tables while the
            # next chunk is still on its way, so the call returns with the whole result in host memory
            r = ctx.pipeline(THRESHOLD, min_width=mw, max_width=MW, window_width=W, min_gain=gain,
                             with_stats=True, host_trace=xp, export=True, **rules)
            assert len(r["segment_table"]["mean"]) == r["segments"]
            return r
        if not e2e["loaded"]:
            shard.load(xp)                     # first step of a run: nothing was prefetched for it
        e2e["remaining"] -= 1
        e2e["loaded"] = e2e["remaining"] > 0
        if e2e["loaded"]:
            shard.prefetch(xp)                 # the next step's trace starts its way up before this step runs
        r = shard.step(THRESHOLD, rules, mw, MW, W, gain)
        if rank == 0:
            # every GPU holds the whole result; the caller reads it once.  The copy-out is enqueued on a side stream
            # and collected before the next one is started; finish_e2e() collects the last one.
            if last_download.get("pending") is not None:
                last_download["tables"] = last_download["pending"].wait()
            last_download["pending"] = shard.download_async()
        if e2e["loaded"]:
            shard.swap()
        return r

    def finish_e2e():
        if shard is not None and last_download.get("pending") is not None:
            last_download["tables"] = last_download["pending"].wait()
            last_download["pending"] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r = None
        e2e["remaining"], e2e["loaded"] = steps, False
        for _ in range(steps):
            r = fn()
        if shard is not None:
            shard.wait()   # the last step's asynchronous table all-gather belongs to the timed region, and so does
            shard.join()   # the copy-out of its tables on the side stream: the closing event waits for both
        e1.record(stream)
        barrier()
        if fn is step_e2e:
            finish_e2e()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, r

    if shard is None:
        ctx.upload_trace(xp)
    else:
        shard.load(xp)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    launches0 = ctx.launch_count
    sampler.start()
    ms, res = timed(step_resident, args.steps)
    launches = ctx.launch_count - launches0
    stage = ctx.stage_ms() if shard is None else dict(shard.stage_ms)
    counters = ctx.split_counters()
    # split-kernel duration averaged over a few more steps (CUDA events inside the library, on the launch stream);
    # the clock sampler keeps polling through them -- same load -- so that a short timed region still yields samples
    split_ms = []
    k = 0
    while k < min(args.steps, 5) or (world == 1 and len(sampler.samples) < 8 and k < 200):  # ranks stay in step
        step_resident()
        split_ms.append(ctx.stage_ms()["split"] if shard is None else shard.stage_ms["split"])
        k += 1
    split_ms = float(np.mean(split_ms))
    clocks = sampler.stop()

    e2e["remaining"] = 2
    for _ in range(2):
        step_e2e()
    finish_e2e()
    ms_e2e, res_e2e = timed(step_e2e, args.steps)

    totals = torch.tensor([n_local, res["events"], res["segments"], res["event_samples"]], device="cuda",
                          dtype=torch.float64)
    if world > 1:
        if shard is not None:
            totals = torch.tensor([shard.n_owned, res["events"], res["segments"], res["event_samples"]],
                                  device="cuda", dtype=torch.float64)
        dist.all_reduce(totals)
    n_total, ev_total, seg_total, evs_total = [int(v) for v in totals.tolist()]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        sec = ms / 1e3 / args.steps
        value = n_total / sec / 1e6
        sec_e2e = ms_e2e / 1e3 / args.steps
        # dominant kernel: k3_split.  Algorithmic bytes: one {c, c2} pair (16 B) per candidate.
        split_bytes = 16.0 * counters["candidates"]
        split_gbs = split_bytes / (split_ms / 1e3) / 1e9
        traffic, traffic_src = ncu_traffic("k3_split") if world == 1 and epg == EVENTS_PER_GPU else (None, None)
        b_floor = 4.0 * n_total + 56.0 * seg_total + 24.0 * ev_total
        # SURVEY 8d's second figure, the reference's data flow: threshold reads N, the prefix kernel reads N_ev and
        # writes {c, c2}, the search touches one pair per candidate, statistics re-read N_ev, 56 B per segment row
        # (candidates are counted on rank 0; every rank holds the same workload)
        b_alg = 4.0 * n_total + 20.0 * evs_total + 16.0 * counters["candidates"] * world + 4.0 * evs_total + 56.0 * seg_total
        seg_row = 4 + 8 + 8 + 32
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, epg),
            "e2e": {"value": n_total / sec_e2e / 1e6, "unit": "Msamples/s",
                    "h2d_bytes_per_step": 4 * n_total,
                    "d2h_bytes_per_step": 56 * seg_total + 16 * ev_total + 160 * world},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k3_split", "achieved": split_gbs, "peak": peak,
                         "unit": "GB/s", "frac": split_gbs / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": split_bytes,
                         "note": "algorithmic bytes = one 16 B {c,c2} pair per candidate evaluation (SURVEY 8d); the "
                                 "prefix sums are re-read from L2 once per recursion level, so DRAM traffic is below "
                                 "the algorithmic bytes and the kernel is bound by instruction issue / latency "
                                 "(DESIGN.md 4)"},
            "pipeline_roofline": {"b_floor_bytes": b_floor, "achieved": b_floor / sec / 1e9 / world,
                                  "peak": peak, "unit": "GB/s per GPU",
                                  "frac": b_floor / sec / 1e9 / world / peak,
                                  "b_alg_bytes": b_alg, "achieved_b_alg": b_alg / sec / 1e9 / world,
                                  "frac_b_alg": b_alg / sec / 1e9 / world / peak},
            "stage_ms": stage, "split_ms": split_ms,
            "counts": {"samples": n_total, "events": ev_total, "event_samples": evs_total,
                       "segments": seg_total, "candidates_rank0": counters["candidates"]},
        }
        if shard is not None:
            line["host_planned_fallback_steps"] = int(shard.fallbacks)
            if epg == EVENTS_PER_GPU:
                line["parity"] = parity_sharded(last_download.get("tables"), world)
        elif epg == EVENTS_PER_GPU:
            line["parity"] = parity_against_fixture(res_e2e)
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = min(epg, 1500)
            xc = synth.make_trace(n_cpu, seed=1, tier="A").astype(np.float64)
            dt, ne, ns, kind = cpu_reference_run(xc, 1)
            line["cpu_baseline"] = {
                "value": len(xc) / dt / 1e6, "unit": "Msamples/s", "cores": 1, "kind": kind,
                "sample": "first %d events (%d samples) of the workload trace, one pass, single process "
                          "(the reference is single-threaded)" % (n_cpu, len(xc))}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--events-per-gpu", type=int, default=EVENTS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--split-kernel", default=None, choices=["flow", "level"],
                    help="development: force k3_flow / k3_split (default: the library's default)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
