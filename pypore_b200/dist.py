"""Multi-GPU sharding of the segmentation path.

One process per GPU.  The trace is cut into contiguous chunks, one per rank.
Every rank threshold-scans its own chunk; a run of samples that crosses a
chunk boundary is owned by the rank where it STARTS.  Ranks all-gather one tiny
record about their first and last run, every rank derives the same global plan
from those records, owners of straddling events receive the continuation
samples (the halo) from their right neighbour(s) point-to-point, and after the
(independent) split / statistics stages the compact event and segment tables are
all-gathered so that every rank ends with the whole result.  Sample data never
goes through a collective: only halos (a few thousand samples) and tables move.

The planning functions are pure (NumPy in, plain Python out) and the exchange
helpers work on CPU tensors with the gloo backend as well as on CUDA tensors
with NCCL, which is how tests/test_dist_cpu.py covers them without a GPU.
"""
import numpy as np

INFO_LEN = 12
SEG_WORDS = 4         # 8-byte words per packed segment row (pp_pack_tables): id|start, mean, std, min|max as float32
_REDO_MASK = 16 | 1   # result-record flags of the device-planned step: PP_OVF_HALO | PP_OVF_RUNS (include/pypore_b200.h)
(I_N, I_NRUNS, I_FIRST_BELOW, I_FIRST_LEN, I_FIRST_MIN, I_FIRST_MAX,
 I_LAST_BELOW, I_LAST_START, I_LAST_LEN, I_LAST_MIN, I_LAST_MAX, I_PAD) = range(INFO_LEN)


def boundary_info(n_local, n_runs, first_run, last_run):
    """The record a rank publishes: (start, length, min, max, below) of its first and last run."""
    info = np.zeros(INFO_LEN, np.float64)
    info[I_N], info[I_NRUNS] = n_local, n_runs
    info[I_FIRST_BELOW], info[I_FIRST_LEN] = float(first_run[4]), first_run[1]
    info[I_FIRST_MIN], info[I_FIRST_MAX] = first_run[2], first_run[3]
    info[I_LAST_BELOW], info[I_LAST_START], info[I_LAST_LEN] = float(last_run[4]), last_run[0], last_run[1]
    info[I_LAST_MIN], info[I_LAST_MAX] = last_run[2], last_run[3]
    return info


def rules_accept(rules, duration, mn, mx):
    """The default-shaped rules of pp_select_events, evaluated on the host for one merged run."""
    m = rules.get("rule_mask", 0)
    ok = True
    if m & 1:
        ok = ok and duration > rules["duration_gt"]
    if m & 8:
        ok = ok and duration < rules["duration_lt"]
    if m & 2:
        ok = ok and bool(mn > rules["min_gt"])   # NaN compares false, like the reference's rule
    if m & 4:
        ok = ok and bool(mx < rules["max_lt"])
    return ok


def plan_boundaries(infos, rules):
    """Global plan from the gathered records.  Returns a list (one dict per rank) of
        skip_first  this rank's first run belongs to a rank on its left
        skip_last   this rank's last run is decided here on the host (it crosses the right boundary)
        event       None, or (start_local, total_length) of the straddling event this rank owns and keeps
        recv        [(src_rank, count)] halo pieces to append after the chunk, in order
        send        [(dst_rank, count)] prefixes of this chunk to ship to owners on the left
    """
    world = len(infos)
    joins = [bool(infos[r][I_LAST_BELOW] == infos[r + 1][I_FIRST_BELOW]) for r in range(world - 1)]
    plans = [dict(skip_first=(r > 0 and joins[r - 1]), skip_last=False, event=None, recv=[], send=[])
             for r in range(world)]
    for r in range(world - 1):
        if not joins[r]:
            continue
        plans[r]["skip_last"] = True
        owns = not (infos[r][I_NRUNS] == 1 and plans[r]["skip_first"])
        if not owns:
            continue  # the run started further left; that owner accounts for this chunk
        length = int(infos[r][I_LAST_LEN])
        mn, mx = infos[r][I_LAST_MIN], infos[r][I_LAST_MAX]
        pieces = []
        q = r + 1
        while True:
            cnt = int(infos[q][I_FIRST_LEN])
            pieces.append((q, cnt))
            length += cnt
            mn = np.minimum(mn, infos[q][I_FIRST_MIN])  # np.minimum / maximum propagate NaN like np.min / np.max
            mx = np.maximum(mx, infos[q][I_FIRST_MAX])
            if infos[q][I_NRUNS] == 1 and q < world - 1 and joins[q]:
                q += 1
                continue
            break
        if rules_accept(rules, length, mn, mx):
            plans[r]["event"] = (int(infos[r][I_LAST_START]), length)
            plans[r]["recv"] = pieces
            for q, cnt in pieces:
                plans[q]["send"].append((r, cnt))
    return plans


def exchange_halo(plan, chunk, halo_out, dist, group=None):
    """Point-to-point halo transfer.  `chunk` is this rank's samples (1-D tensor, CPU or CUDA),
    `halo_out` a tensor with room for sum(recv counts).  Returns the number of samples received."""
    ops = []
    off = 0
    for src, cnt in plan["recv"]:
        ops.append(dist.P2POp(dist.irecv, halo_out[off:off + cnt], src, group))
        off += cnt
    for dst, cnt in plan["send"]:
        ops.append(dist.P2POp(dist.isend, chunk[:cnt], dst, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return off


def gather_infos(info, dist, device, group=None):
    import torch
    world = dist.get_world_size(group)
    t = torch.from_numpy(info).to(device)
    out = torch.empty(world * INFO_LEN, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.cpu().numpy().reshape(world, INFO_LEN)


def gather_tables(int_cols, flt_cols, counts, dist, group=None):
    """All-gather per-rank tables of different lengths: pad to the longest, gather, strip.

    int_cols: int64 tensor [n_local, ci]; flt_cols: float64 tensor [n_local, cf];
    counts: list of per-rank row counts (already known to every rank).
    Returns (int64 [sum(counts), ci], float64 [sum(counts), cf]).
    """
    import torch
    world = len(counts)
    m = max(max(counts), 1)
    out = []
    for cols in (int_cols, flt_cols):
        pad = torch.zeros((m, cols.shape[1]), dtype=cols.dtype, device=cols.device)
        pad[:cols.shape[0]] = cols
        g = torch.empty((world * m, cols.shape[1]), dtype=cols.dtype, device=cols.device)
        dist.all_gather_into_tensor(g, pad, group=group)
        g = g.view(world, m, cols.shape[1])
        out.append(torch.cat([g[r, :counts[r]] for r in range(world)], dim=0))
    return out[0], out[1]


def gather_packed_raw(words, m, dist, group=None, async_op=False):
    """ONE all-gather of the first m words of every rank's packed tables (pp_pack_tables layout: 2 words per
    event, then SEG_WORDS per segment).  Returns the int64 tensor [world, m]; row r holds rank r's words.
    With async_op the collective is only enqueued: returns (tensor, work) and the caller waits on `work`."""
    import torch
    world = dist.get_world_size(group)
    if words.shape[0] < m:
        words = torch.cat([words, torch.zeros(m - words.shape[0], dtype=words.dtype, device=words.device)])
    g = torch.empty(world * m, dtype=torch.int64, device=words.device)
    work = dist.all_gather_into_tensor(g, words[:m].contiguous(), group=group, async_op=async_op)
    return (g.view(world, m), work) if async_op else g.view(world, m)


def split_gathered(g, counts):
    """Contiguous raw rows from gather_packed_raw's result; counts = per-rank (events, segments).
    Returns (events int64 [sum E, 2], segment words int64 [sum S, SEG_WORDS])."""
    import torch
    world = len(counts)
    sizes = [2 * e + SEG_WORDS * s for e, s in counts]
    ev = torch.cat([g[r, :2 * counts[r][0]].view(-1, 2) for r in range(world)], dim=0)
    seg = torch.cat([g[r, 2 * counts[r][0]:sizes[r]].view(-1, SEG_WORDS) for r in range(world)], dim=0)
    return ev, seg


def unpack_gathered(g, counts):
    """The whole result as contiguous tables: dict(events int64 [sum E, 2] {global start, length},
    seg_int int64 [sum S, 3] {global event id, start, end}, seg_flt float64 [sum S, 4] {mean, std, min, max}).
    Undoes the packing of pp_pack_tables: `end` is the next row's start inside the same event, else the
    event's length; min / max come back from their float32 bit patterns."""
    import torch
    ev, seg = split_gathered(g, counts)
    w0 = seg[:, 0]
    event = w0 & 0xffffffff
    start = w0 >> 32
    n = seg.shape[0]
    if n:
        nxt_same = torch.zeros(n, dtype=torch.bool, device=seg.device)
        nxt_same[:-1] = event[1:] == event[:-1]
        nxt_start = torch.cat([start[1:], start[:1]])
        end = torch.where(nxt_same, nxt_start, ev[:, 1][event])
    else:
        end = start
    mnmx = seg[:, 3].contiguous().view(torch.int32).view(-1, 2).view(torch.float32).to(torch.float64)
    flt = torch.stack([seg[:, 1].contiguous().view(torch.float64), seg[:, 2].contiguous().view(torch.float64),
                       mnmx[:, 0], mnmx[:, 1]], dim=1)
    return dict(events=ev, seg_int=torch.stack([event, start, end], dim=1), seg_flt=flt)


def gather_packed(words, counts, dist, group=None):
    """gather_packed_raw + split_gathered for callers that know every rank's counts already."""
    m = max(max(2 * e + SEG_WORDS * s for e, s in counts), 1)
    return split_gathered(gather_packed_raw(words, m, dist, group), counts)


# ------------------------------------------------------------------------------------------
# synthetic sharded workload (bench / tests): one C2-style piece per rank, cut mid-event
# ------------------------------------------------------------------------------------------
def _cut_offset(head, threshold=110.0, depth=2500):
    first_below = int(np.argmax(head < threshold))
    return first_below + depth


def synthetic_chunk(rank, world, events_per_rank, seed0=1, tier="A"):
    """Rank `rank`'s chunk of the global trace concat(P_0 .. P_{world-1}), where P_g =
    synth.make_trace(events_per_rank, seed0 + g) and the ownership boundary between g-1 and g lies
    2500 samples inside P_g's first event, so an event straddles every boundary."""
    from . import synth
    piece = synth.make_trace(events_per_rank, seed=seed0 + rank, tier=tier, tail=(rank == world - 1))
    lo = _cut_offset(piece[:20000]) if rank > 0 else 0
    parts = [piece[lo:]]
    if rank < world - 1:
        head = synth.make_trace(1, seed=seed0 + rank + 1, tier=tier, tail=False)
        parts.append(head[:_cut_offset(head)])
    return np.ascontiguousarray(np.concatenate(parts))


def synthetic_global(world, events_per_rank, seed0=1, tier="A"):
    """The whole trace the chunks of synthetic_chunk partition (tests only)."""
    from . import synth
    return np.concatenate([synth.make_trace(events_per_rank, seed=seed0 + g, tier=tier, tail=(g == world - 1))
                           for g in range(world)])


# ------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------
class _DevArray(object):
    """Zero-copy torch view of device memory owned by the C library."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def device_view(ptr, n, dtype, device):
    import torch
    typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int64: "<i8", torch.int32: "<i4"}[dtype]
    if n == 0:
        return torch.empty(0, dtype=dtype, device=device)
    return torch.as_tensor(_DevArray(ptr, n, typestr), device=device)


class ShardedPipeline(object):
    """threshold -> select -> halo -> split -> stats -> table all-gather on one rank's chunk.

    Every torch / NCCL operation of a step is issued on the context's own CUDA stream (wrapped as a
    torch ExternalStream), so collectives and the library's kernels are ordered without extra synchronisation.
    """

    HALO_CAPACITY = 1 << 20
    HALO_SPECULATIVE = 1 << 16   # samples every rank ships to its left neighbour up front (256 KB over NVLink)
    STAGED_DOWNLOAD = not __import__("os").environ.get("PYPORE_B200_DIRECT_DOWNLOAD")   # see download_async
    RECORD_RING = 4              # steps that may be in flight between step_async() and finish()

    def __init__(self, ctx, rank, world, group=None):
        import torch
        import torch.distributed as dist
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.dist = dist
        self.device = torch.device("cuda", ctx.device)
        self.stream = torch.cuda.ExternalStream(ctx.stream_handle, device=self.device)
        self.halo = torch.empty(self.HALO_CAPACITY, dtype=torch.float32, device=self.device)
        self.n_local = 0
        self.n_owned = 0
        self.offsets = None
        self.gathered = self.counts = self._tables = None
        self.rec = self.res = self.pack = self.plan = self.infos_dev = self.allr_dev = None
        self.pad_words = 0
        self.lens = None
        self._halo = 0       # speculative halo samples resident after the chunk
        self._table_work = None   # outstanding asynchronous table all-gather
        self._side = None         # side stream of download_async
        self._stage = None        # device arena of download_async
        self._rec_ring, self._rec_next = None, 0   # page-locked slots for the result records of steps in flight
        self._n_agreed = -1
        self._next_lens = []      # chunk lengths of the prefetched traces, oldest first
        # the tiny control collectives (chunk lengths, boundary records, result records) get their own communicator:
        # they sit on every step's critical path and must not queue behind the previous step's table all-gather
        self.ctl_group = dist.new_group() if world > 1 and dist.is_initialized() else group
        self.fallbacks = 0   # steps that had to be repeated with the host-made plan
        self.peer_ctl = self._open_peer_ctl()

    def _open_peer_ctl(self):
        """Map every rank's control-record buffer (CUDA IPC over NVLink), once.  Returns True when EVERY rank
        succeeded -- the step then exchanges its boundary / result records with pp_ctl_exchange (plain peer stores +
        flags, one tiny launch) instead of NCCL all-gathers; otherwise all ranks keep the collectives."""
        import os
        import torch
        dist = self.dist
        if (self.world <= 1 or not dist.is_initialized() or not hasattr(self.ctx, "ctl_create")
                or os.environ.get("PYPORE_B200_NO_PEER_CTL")):
            return False
        ok = 1
        handle = b"\0" * 64
        try:
            handle, _ = self.ctx.ctl_create(self.rank, self.world)
        except Exception:
            ok = 0
        with torch.cuda.stream(self.stream):
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
            allh = torch.empty(self.world * 64, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allh, mine, group=self.ctl_group)
            handles = bytes(allh.cpu().numpy().tobytes())
            if ok:
                try:
                    self.ctx.ctl_open(handles=handles)
                except Exception:
                    ok = 0
            agreed = torch.tensor([ok], dtype=torch.int64, device=self.device)
            dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=self.ctl_group)
            return bool(int(agreed.item()))

    def _exchange_records(self, src, n_words, dst):
        """Every rank's `src` record (n_words 8-byte words) in `dst` on every rank."""
        if self.peer_ctl:
            self.ctx.ctl_exchange(src.data_ptr(), n_words, dst.data_ptr())
        else:
            self.dist.all_gather_into_tensor(dst, src, group=self.ctl_group)

    def _agree_lengths(self, n_local):
        """The chunk lengths are exchanged when they change, not once per load: every rank takes the same branch
        (`same` is agreed on by a tiny all-reduce)."""
        import torch
        with torch.cuda.stream(self.stream):
            same = torch.tensor([1 if (self.lens is not None and n_local == self._n_agreed) else 0],
                                dtype=torch.int64, device=self.device)
            self.dist.all_reduce(same, op=self.dist.ReduceOp.MIN, group=self.ctl_group)
            if int(same.item()) == 0:
                mine = torch.tensor([n_local], dtype=torch.int64, device=self.device)
                lens = torch.empty(self.world, dtype=torch.int64, device=self.device)
                self.dist.all_gather_into_tensor(lens, mine, group=self.ctl_group)
                self.lens = lens.cpu().numpy()
        self._n_agreed = n_local

    def load(self, host_chunk):
        """Upload this rank's chunk; ranks exchange their chunk lengths when they change (sample offsets of the
        global trace and the size of the speculative halo each neighbour sends)."""
        n_local = int(host_chunk.shape[0])
        self._agree_lengths(n_local)
        self.ctx.upload_trace_async(host_chunk, extra_capacity=self.HALO_CAPACITY)
        self.n_local = n_local
        self._exchange_speculative_halo()

    def prefetch(self, host_chunk):
        """Back-to-back traces: start the upload of the NEXT chunk (same length on every call in a row) on the
        context's copy stream; it runs under the step on the resident chunk.  swap() makes the oldest prefetched
        chunk the resident one.  Two chunks may be waiting: with the one after next already queued behind the running
        copy, the upload link does not idle while the ranks meet for the halo exchange between two steps."""
        n_local = int(host_chunk.shape[0])
        self._agree_lengths(n_local)
        self.ctx.prefetch_trace(host_chunk, extra_capacity=self.HALO_CAPACITY)
        self._next_lens.append(n_local)

    def swap(self):
        self.ctx.swap_trace()
        self.n_local = self._next_lens.pop(0)
        self._exchange_speculative_halo()

    def _exchange_speculative_halo(self):
        """The head of every chunk travels to the left neighbour as soon as the chunk is on the device -- it is
        input data, independent of any scan -- directly into the room reserved after that rank's chunk (256 KB
        per boundary over NVLink).  Steps on the resident trace find it there; k_shard_plan decides on the
        device whether it covers the straddling event."""
        import torch
        ctx, dist, dev = self.ctx, self.dist, self.device
        with torch.cuda.stream(self.stream):
            ops, halo = [], 0
            if self.rank < self.world - 1:
                halo = self._spec_halo(self.rank + 1)
                room = device_view(ctx.trace_ptr + 4 * self.n_local, halo, torch.float32, dev)
                ops.append(dist.P2POp(dist.irecv, room, self.rank + 1, self.group))
            if self.rank > 0:
                chunk = device_view(ctx.trace_ptr, self._spec_halo(self.rank), torch.float32, dev)
                ops.append(dist.P2POp(dist.isend, chunk, self.rank - 1, self.group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()                                  # stream-level wait, the host goes on
            self._halo = halo
        self.offsets = np.concatenate(([0], np.cumsum(self.lens)))

    def step(self, threshold, rules, mw, MW, W, gain, host_planned=False):
        """One pass over the sharded trace.  Default: the device-planned step (one host synchronisation, at the
        end); `host_planned=True`, or a straddling event the speculative halo does not cover, takes the step
        with the host-made plan (two more round trips)."""
        import torch
        with torch.cuda.stream(self.stream):
            if not host_planned:
                r = self._step_device_planned(threshold, rules, mw, MW, W, gain)
                if r is not None:
                    return r
                self.fallbacks += 1
            return self._step_host_planned(threshold, rules, mw, MW, W, gain)

    def _spec_halo(self, q):
        """Samples rank q sends to rank q-1 before anything is known about the runs."""
        return int(min(self.HALO_SPECULATIVE, self.lens[q]))

    def _step_device_planned(self, threshold, rules, mw, MW, W, gain):
        """scan -> all-gather of the boundary records -> k_shard_plan on the device -> select/prefix/split/
        stats -> all-gather of the result records -> packed table all-gather: no host synchronisation until
        the result records are read.  The head of the right neighbour's chunk is already resident behind this
        rank's chunk (load() put it there).  Returns None when the records ask for the host-planned step."""
        import torch
        ctx, dist, dev = self.ctx, self.dist, self.device
        if self.rec is None:
            self.rec = torch.zeros(INFO_LEN, dtype=torch.float64, device=dev)
            self.res = torch.zeros(8, dtype=torch.int64, device=dev)
        if self.plan is None:
            self.plan = torch.zeros(8, dtype=torch.int64, device=dev)
            self.infos_dev = torch.empty(self.world * INFO_LEN, dtype=torch.float64, device=dev)
        ctx.truncate_trace(self.n_local)
        ctx.shard_scan(threshold, self.n_local, self.rec.data_ptr())
        self._exchange_records(self.rec, INFO_LEN, self.infos_dev)
        halo = self._halo
        ctx.extend_trace(halo)
        ctx.shard_plan(self.infos_dev.data_ptr(), self.rank, self.world, threshold, rules, halo, self.plan.data_ptr())
        ctx.shard_finish_planned(threshold, rules, mw, MW, W, gain, self.plan.data_ptr(), self.res.data_ptr())
        return self._gather_results(redo_mask=_REDO_MASK)

    def _gather_results(self, redo_mask=0):
        """Exchange of the result records and all-gather of the packed tables, then the only host synchronisation."""
        return self._finish_results(self._enqueue_results(), redo_mask)

    def _enqueue_results(self):
        """Device side of the step's end: result records to every rank, tables packed and their all-gather started,
        the records on their way to page-locked host memory.  No host synchronisation; returns a ticket."""
        import torch
        ctx, dist, dev = self.ctx, self.dist, self.device
        allr_dev = torch.empty(self.world * 8, dtype=torch.int64, device=dev)
        self._exchange_records(self.res, 8, allr_dev)
        # Speculative sizing: pack and all-gather with the padding of the previous step (+12.5 %) before the host
        # knows this step's counts -- the pack kernel takes them from allr_dev on the device -- so the GPU is not
        # idle during the host round trip.  Every rank reads the same records and takes the same decision.
        # The table all-gather is asynchronous: the context's stream does not wait for it here, so the next step's
        # scan (or whatever the caller enqueues next) overlaps it; `tables` / `download()` and the next step's pack
        # kernel wait for it.
        self._wait_tables()                      # the previous step's gather still reads self.pack
        g = work = None
        pad = self.pad_words
        if pad:
            if self.pack is None or self.pack.shape[0] < pad:
                self.pack = torch.empty(pad, dtype=torch.int64, device=dev)
            ctx.pack_tables(allr_dev.data_ptr(), self.rank, int(self.offsets[self.rank]), self.pack.data_ptr(), pad)
            g, work = gather_packed_raw(self.pack, pad, dist, self.group, async_op=True)
            self._table_work = work
        # page-locked room for the records: a small ring owned by the context (at most RECORD_RING tickets may be
        # outstanding), not torch's pinned allocator -- that one records events when a block is released, which
        # at interpreter exit is after the CUDA context is gone
        if self._rec_ring is None:
            self._rec_ring = [torch.from_numpy(ctx.pinned_empty(self.world * 8, np.int64))
                              for _ in range(self.RECORD_RING)]
        allr_host = self._rec_ring[self._rec_next % self.RECORD_RING]
        self._rec_next += 1
        allr_host.copy_(allr_dev, non_blocking=True)
        done = torch.cuda.Event()
        done.record(self.stream)
        return dict(allr_dev=allr_dev, allr_host=allr_host, done=done, g=g, work=work, pad=pad, n_local=self.n_local)

    def _finish_results(self, t, redo_mask=0):
        """Host side: read the records (waits for the step that produced them, not for anything enqueued after it),
        commit the counts, redo the table all-gather if the speculative padding was too small."""
        import torch
        ctx, dist, dev = self.ctx, self.dist, self.device
        t["done"].synchronize()
        allr_dev, g, work = t["allr_dev"], t["g"], t["work"]
        allr = t["allr_host"].numpy().reshape(self.world, 8).copy()
        self.allr_dev = allr_dev
        if redo_mask and (allr[:, 4] & redo_mask).any():
            if work is not None:
                work.wait()
            return None
        ctx.shard_commit(allr[self.rank])
        self.n_owned = t["n_local"]
        counts = [(int(r[1]), int(r[3])) for r in allr]
        need_words = max(max(2 * e + SEG_WORDS * s for e, s in counts), 1)
        if g is None or need_words > t["pad"]:
            with torch.cuda.stream(self.stream):
                if work is not None:
                    work.wait()
                self._wait_tables()
                self.pad_words = (int(need_words * 1.125) + 65) & ~1   # even: rows of the gathered buffer stay 16-byte aligned
                self.pack = torch.empty(self.pad_words, dtype=torch.int64, device=dev)
                ctx.pack_tables(allr_dev.data_ptr(), self.rank, int(self.offsets[self.rank]), self.pack.data_ptr(),
                                self.pad_words)
                g, work = gather_packed_raw(self.pack, self.pad_words, dist, self.group, async_op=True)
                self._table_work = work
        else:
            self.pad_words = max((int(need_words * 1.125) + 65) & ~1, 2)
        self.gathered, self.counts, self._tables = g, counts, None
        ne, n_seg = counts[self.rank]
        return dict(runs=int(allr[self.rank, 0]), events=ne, event_samples=int(allr[self.rank, 2]), segments=n_seg)

    def step_async(self, threshold, rules, mw, MW, W, gain):
        """The device-planned step WITHOUT its closing host synchronisation: everything is enqueued and a ticket comes
        back; `finish(ticket)` reads the result records.  A caller that runs steps on the same resident trace back to
        back enqueues step i+1 before it finishes step i, so the device never waits for the host (the records of
        step i are read while step i+1 runs).  A step whose records ask for the host-planned redo is repeated inside
        finish() -- on the trace that is resident THEN, so do not swap traces between step_async and finish."""
        import torch
        with torch.cuda.stream(self.stream):
            ctx = self.ctx
            dev = self.device
            if self.rec is None:
                self.rec = torch.zeros(INFO_LEN, dtype=torch.float64, device=dev)
                self.res = torch.zeros(8, dtype=torch.int64, device=dev)
            if self.plan is None:
                self.plan = torch.zeros(8, dtype=torch.int64, device=dev)
                self.infos_dev = torch.empty(self.world * INFO_LEN, dtype=torch.float64, device=dev)
            ctx.truncate_trace(self.n_local)
            ctx.shard_scan(threshold, self.n_local, self.rec.data_ptr())
            self._exchange_records(self.rec, INFO_LEN, self.infos_dev)
            ctx.extend_trace(self._halo)
            ctx.shard_plan(self.infos_dev.data_ptr(), self.rank, self.world, threshold, rules, self._halo,
                           self.plan.data_ptr())
            ctx.shard_finish_planned(threshold, rules, mw, MW, W, gain, self.plan.data_ptr(), self.res.data_ptr())
            t = self._enqueue_results()
        t["args"] = (threshold, rules, mw, MW, W, gain)
        return t

    def finish(self, ticket):
        import torch
        with torch.cuda.stream(self.stream):
            r = self._finish_results(ticket, redo_mask=_REDO_MASK)
            if r is None:
                self.fallbacks += 1
                r = self._step_host_planned(*ticket["args"])
        return r

    def _step_host_planned(self, threshold, rules, mw, MW, W, gain):
        """Two more host synchronisations: after the all-gather of the boundary records the plan is made on
        the host (plan_boundaries) and halos of any length travel point-to-point from any number of ranks."""
        import torch
        ctx, dist, dev = self.ctx, self.dist, self.device
        if self.rec is None:
            self.rec = torch.zeros(INFO_LEN, dtype=torch.float64, device=dev)
            self.res = torch.zeros(8, dtype=torch.int64, device=dev)
        for attempt in range(2):
            ctx.truncate_trace(self.n_local)
            ctx.shard_scan(threshold, self.n_local, self.rec.data_ptr())
            out = torch.empty(self.world * INFO_LEN, dtype=torch.float64, device=dev)
            self._exchange_records(self.rec, INFO_LEN, out)
            infos = out.cpu().numpy().reshape(self.world, INFO_LEN)      # host sync 1
            if not infos[:, I_PAD].any():
                break
            # some rank's run table overflowed (very noisy chunk): grow it everywhere and scan again
            ctx.threshold_scan(threshold, scan_len=self.n_local)
        else:
            raise RuntimeError("run table overflow")
        plan = plan_boundaries(infos, rules)[self.rank]
        need = sum(c for _, c in plan["recv"])
        if need > self.halo.shape[0]:
            raise RuntimeError("halo of %d samples exceeds HALO_CAPACITY" % need)
        chunk = device_view(ctx.trace_ptr, self.n_local, torch.float32, dev)
        got = exchange_halo(plan, chunk, self.halo, dist, self.group)
        if got:
            ctx.append_trace(self.halo.data_ptr(), got, True)
        ctx.shard_finish(threshold, rules, mw, MW, W, gain, plan["skip_first"], plan["skip_last"], plan["event"],
                         self.res.data_ptr())
        return self._gather_results()

    @property
    def stage_ms(self):
        """Stage times of the last step (read on demand: seven event queries are host time between steps)."""
        return self.ctx.stage_ms()

    def wait(self):
        """The context's stream waits for everything the last step left in flight (the table all-gather)."""
        self._wait_tables()

    def join(self):
        """The context's stream waits for the side stream of download_async (stream-level, the host goes on)."""
        if self._side is not None:
            self.stream.wait_stream(self._side)

    def _wait_tables(self):
        """Make the context's stream wait for the outstanding table all-gather (stream-level, the host goes on)."""
        if self._table_work is not None:
            import torch
            with torch.cuda.stream(self.stream):
                self._table_work.wait()
            self._table_work = None

    @property
    def tables(self):
        """The whole result on this GPU as contiguous tensors (built from the gathered buffer on first use)."""
        if self._tables is None and self.gathered is not None:
            import torch
            self._wait_tables()
            with torch.cuda.stream(self.stream):
                self._tables = unpack_gathered(self.gathered, self.counts)
        return self._tables

    def download(self):
        """The gathered tables in host memory (what a caller of the public API receives), one array per column:
        ev_start, ev_len (global samples), seg_event (global event id), seg_start, seg_end, mean, std, min, max.
        ONE kernel unpacks the all-gathered rows straight into page-locked host tables (pp_unpack_tables); the
        arrays are views of the context's pinned arena, valid until its next pinned use.  Without a library context
        (the CPU tests) the tensor path answers."""
        import torch
        if self.gathered is None:
            return None
        if not hasattr(self.ctx, "unpack_tables") or self.gathered.device.type != "cuda":
            with torch.cuda.stream(self.stream):
                return columns({k: v.cpu().numpy() for k, v in self.tables.items()})
        self._wait_tables()
        n_ev = sum(c[0] for c in self.counts)
        n_seg = sum(c[1] for c in self.counts)
        return self.ctx.unpack_tables(self.gathered.data_ptr(), self.world, self.gathered.shape[1],
                                      self.allr_dev.data_ptr(), n_ev, n_seg)

    def download_async(self, own_rows=False):
        """download() that only ENQUEUES the copy-out, on a side stream: returns a handle whose wait() gives the
        column arrays.  The next step's load() -- the other PCIe direction -- overlaps it.  Two pinned table arenas
        alternate, so the arrays of one handle stay valid until the handle after the next is created.
        `own_rows`: only the rows of this rank's chunk (global event ids and starts).  With one process per GPU the
        result then reaches host memory the way the trace went up -- every rank's share over its own link -- instead
        of one rank pulling the whole table through one link (at 8 GPUs: 17 MB per rank instead of 138 MB on one)."""
        import torch
        if self.gathered is None:
            return None
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
            self._slot = 0
        side = self._side
        with torch.cuda.stream(side):
            if self._table_work is not None:
                self._table_work.wait()        # the side stream waits for the table all-gather ...
        side.wait_stream(self.stream)          # ... and for everything the step enqueued on the context's stream
        n_ev = sum(c[0] for c in self.counts)
        n_seg = sum(c[1] for c in self.counts)
        ranks = None
        if own_rows:
            n_ev, n_seg = self.counts[self.rank]
            ranks = (self.rank, self.rank + 1)
        self._slot ^= 1
        if self.STAGED_DOWNLOAD:
            # unpack in device memory, then ONE copy-engine transfer: large posted writes on the upstream PCIe lanes
            # disturb the concurrent upload of the next trace (whose read requests share those lanes) less than the
            # 32-byte stores of a kernel writing host memory directly
            nbytes = self.ctx.unpacked_bytes(n_ev, n_seg)
            if self._stage is None or self._stage.shape[0] < nbytes:
                self._stage = torch.empty(int(nbytes * 1.25), dtype=torch.uint8, device=self.device)
            cols, whole = self.ctx.unpack_tables(self.gathered.data_ptr(), self.world, self.gathered.shape[1],
                                                 self.allr_dev.data_ptr(), n_ev, n_seg, stream=side.cuda_stream,
                                                 slot=self._slot, staging_ptr=self._stage.data_ptr(), ranks=ranks)
            with torch.cuda.stream(side):
                torch.from_numpy(whole).copy_(self._stage[:nbytes], non_blocking=True)
        else:
            cols = self.ctx.unpack_tables(self.gathered.data_ptr(), self.world, self.gathered.shape[1],
                                          self.allr_dev.data_ptr(), n_ev, n_seg, stream=side.cuda_stream,
                                          slot=self._slot, ranks=ranks)
        done = torch.cuda.Event()
        done.record(side)
        keep = (self.gathered, self.allr_dev)   # the kernel reads them: they must outlive it

        class Handle(object):
            def wait(h):
                done.synchronize()
                h.keep = None
                return cols
        h = Handle()
        h.keep = keep
        return h


def bind_near_gpu(device_index):
    """Pin this process to the CPU cores of the NUMA node the GPU hangs off (one process per GPU: the page-locked
    trace buffers are then first-touched on the memory next to the GPU's PCIe root, so eight uploads do not cross
    the socket link).  Call before allocating host buffers.  Returns the number of cores bound to, 0 when the
    topology cannot be read (then nothing changes)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = device_index
        if visible:
            entry = visible.split(",")[device_index].strip()
            if entry.isdigit():
                index = int(entry)
            else:
                index = None
                handle = pynvml.nvmlDeviceGetHandleByUUID(entry.encode() if hasattr(entry, "encode") else entry)
        if index is not None:
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cores = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cores &= os.sched_getaffinity(0)
        if not cores:
            return 0
        os.sched_setaffinity(0, cores)
        return len(cores)
    except Exception:
        return 0


def device_for_rank(local_rank, n_devices):
    """Which GPU of the node a local rank uses when the node has more GPUs than ranks.  On an HGX board the
    first half of the GPUs hangs off one CPU socket and the second half off the other; NVLink is all-to-all,
    so ranks are dealt to the two halves in turn (0, n/2, 1, n/2+1, ...) and a 2- or 4-rank job gets both
    sockets' PCIe and memory bandwidth for its uploads.  (Measured on the 8 x B200 box: four concurrent
    uploads through GPUs 0-3 share 115 GB/s; GPUs 0,4,1,5 get their full 55 GB/s each.)"""
    if n_devices < 2 or n_devices % 2:
        return local_rank % max(n_devices, 1)
    half = n_devices // 2
    r = local_rank % n_devices
    return (r // 2) + half * (r % 2)


def measure_upload_rates(ctx, pinned, dist, device, world, window_ms=100.0, piece=16 << 20):
    """Samples per second every rank's host-to-device path delivers while ALL ranks upload at once (they share
    root ports and host memory channels unevenly).  Every rank copies pieces of `pinned` for the same wall-clock
    window.  The resident trace is overwritten: load again afterwards.  Returns one rate per rank."""
    import time
    import torch
    piece = int(min(piece, pinned.shape[0]))
    part = pinned[:piece]
    dist.barrier()
    ctx.sync()
    t0 = time.perf_counter()
    copied = 0
    while (time.perf_counter() - t0) * 1e3 < window_ms:
        ctx.upload_trace_async(part)
        ctx.sync()
        copied += piece
    mine = torch.tensor([copied / (time.perf_counter() - t0)], dtype=torch.float64, device=device)
    rates = torch.empty(world, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(rates, mine)
    return rates.cpu().numpy()


def proportional_cuts(total, rates):
    """Boundaries (world + 1 sample offsets) that give every rank a share of `total` samples proportional to
    its rate: all uploads then end together."""
    share = np.cumsum(np.asarray(rates, np.float64)) / float(np.sum(rates))
    cuts = np.concatenate(([0], np.round(total * share).astype(np.int64)))
    cuts[-1] = total
    return cuts


def recut_chunk(rank, cuts, lens, own_chunk, other_chunk):
    """This rank's share [cuts[rank], cuts[rank + 1]) of the global trace whose even cut has the chunk lengths `lens`
    (this rank holds `own_chunk`; `other_chunk(q)` produces rank q's chunk of the even cut, for the pieces of the
    new share that lie in a neighbour's old chunk)."""
    offs = np.concatenate(([0], np.cumsum(np.asarray(lens, np.int64))))
    a, b = int(cuts[rank]), int(cuts[rank + 1])
    parts = []
    for q in range(len(lens)):
        lo, hi = max(a, int(offs[q])), min(b, int(offs[q + 1]))
        if lo < hi:
            src = own_chunk if q == rank else other_chunk(q)
            parts.append(src[lo - int(offs[q]):hi - int(offs[q])])
    return np.concatenate(parts) if parts else np.zeros(0, own_chunk.dtype)


def columns(t):
    """dict(events [E,2], seg_int [S,3], seg_flt [S,4]) -> one array per column (the layout of download())."""
    out = dict(ev_start=t["events"][:, 0], ev_len=t["events"][:, 1], seg_event=t["seg_int"][:, 0],
               seg_start=t["seg_int"][:, 1], seg_end=t["seg_int"][:, 2])
    for j, k in enumerate(("mean", "std", "min", "max")):
        out[k] = t["seg_flt"][:, j]
    return out


def rows(c):
    """The inverse of `columns`: stacked row tables from the column arrays of download()."""
    return dict(events=np.stack([c["ev_start"], c["ev_len"]], axis=1),
                seg_int=np.stack([c["seg_event"], c["seg_start"], c["seg_end"]], axis=1),
                seg_flt=np.stack([c["mean"], c["std"], c["min"], c["max"]], axis=1))
