// split_flow.cuh -- K3 (flow form): the recursive changepoint search of FastStatSplit
// (PyPore/cparsers.pyx:157-203) as a persistent kernel WITHOUT level barriers.
//
// k3_split (split.cuh) resolves a task level by level: SCREEN | barrier | DECIDE | barrier.  On B200 its warps spend
// 35 % of their time waiting at those barriers (profiles/r01o_ncu_summary.csv: stall_barrier 4.5 of 13 cycles per
// issue) and the per-event serial steps bound the kernel: 512-thread CTAs are 2.6x slower than 128-thread ones
// (profiles/r02a_k3_geometry_variants.txt).  Here every WARP is an independent worker and a CTA only shares a
// work stack in shared memory:
//
//   * unit of work = a PIECE: a contiguous range of 32-candidate chunks of one window.  Windows of up to
//     K3F_MIN_PIECE chunks are one piece; larger ones are cut into at most K3F_MAXP pieces so that the warps of
//     the CTA cooperate on them.
//   * a warp pops a piece (LIFO: depth first, so a CTA keeps few windows -- and few KB of prefix sums -- live),
//     screens its candidates exactly like k3_split does (k3_screen_lane: two smallest integer keys per lane),
//     and stores the piece summary in the window record.  The warp that finishes the LAST piece of a window
//     decides it on its own -- minimum over the pieces, single contender clear of min_gain: no exact arithmetic;
//     otherwise the contenders are evaluated with the reference's arithmetic by the warp's lanes -- then runs the
//     bookkeeping of _recursive_split for the children (f_place) and pushes their pieces.
//   * all structural state (stack, free list of window records, event slots) is touched by lane 0 of a warp under
//     one CTA-level spin lock in shared memory; the critical sections are a few dozen instructions.
//   * when the stack runs dry a warp takes the next task from the global queue while its siblings finish, so a
//     CTA may hold up to K3F_EVS events at once.
//
// Results are those of k3_split bit for bit: same screening key, same error bound, same exact arithmetic, same
// transition function; only the order in which windows are visited differs, and the result of a window depends
// on nothing but its interval and the event's prefix sums.
#pragma once
#include "split.cuh"

#ifndef K3F_CFG_CTAS
#define K3F_CFG_CTAS 7
#endif
#ifndef K3F_CFG_MIN_PIECE
#define K3F_CFG_MIN_PIECE 16
#endif
#ifndef K3F_CFG_MAXP
#define K3F_CFG_MAXP 8
#endif
#ifndef K3F_CFG_THREADS
#define K3F_CFG_THREADS 128
#endif
constexpr int K3F_THREADS = K3F_CFG_THREADS;
constexpr int K3F_WARPS = K3F_THREADS / 32;
constexpr int K3F_CTAS_PER_SM = K3F_CFG_CTAS;
constexpr int K3F_MAXP = K3F_CFG_MAXP;            // pieces per window (<= 16: a stack entry packs the piece in 4 bits)
constexpr int K3F_MIN_PIECE = K3F_CFG_MIN_PIECE;  // chunks; a window of at most this many chunks is one piece
constexpr int K3F_WINS = 48;                      // window records per CTA
constexpr int K3F_STACK = K3F_WINS * K3F_MAXP;
constexpr int K3F_EVS = 4;                        // events in flight per CTA
#ifndef PP_SPLIT_KERNEL_DEFAULT
#define PP_SPLIT_KERNEL_DEFAULT 0
#endif
constexpr long long K3F_SPIN_LIMIT = 1LL << 26;   // safety net for the waits (a broken queue must not hang the GPU)

struct K3FWin {
    int s, e, ps;            // interval and the start of the window being scanned
    short evs;               // event slot
    unsigned char npieces;
    unsigned char exact;     // pieces are scanned with the reference's arithmetic (screening impossible / switched off)
    int piece_chunks;
    int done;                // finished pieces
    int pad;
    unsigned long long k1[K3F_MAXP];   // screened: smallest key of the piece; exact: bits of its best gain
    unsigned long long k2[K3F_MAXP];   // screened: second smallest key
    int i1[K3F_MAXP];                  // candidate of k1 (| K3_BAD_FLAG); exact: best split position or -1
};

struct K3FEvent {
    long long off;
    int ev;      // -1: slot is free
    int ebase;
    int live;    // windows registered and not yet resolved (+1 while an interval is being placed)
    int pad;
};

struct K3FShared {
    K3FWin win[K3F_WINS];
    int free_list[K3F_WINS];
    int stack[K3F_STACK];
    K3FEvent evs[K3F_EVS];
    int n_free, top, lock, fetching, drained, few, failed, pad;
    unsigned long long cand, scans, exact, tasks;
    K3Global G;
    K3Params P;
};

constexpr size_t K3F_SMEM_BYTES = sizeof(K3FShared);

__device__ __forceinline__ void f_lock(K3FShared &S)
{
    while (atomicCAS(&S.lock, 0, 1) != 0) { }
    __threadfence_block();
}

__device__ __forceinline__ void f_unlock(K3FShared &S)
{
    __threadfence_block();
    atomicExch(&S.lock, 0);
}

__device__ __forceinline__ void f_push_global(const K3Global &G, int ev, int s, int e, int ps)
{
    atomicAdd((unsigned long long *)&G.ctr->q_pending, 1ull);
    const unsigned long long slot = atomicAdd(&G.ctr->q_tail, 1ull);
    if ((int64_t)slot >= G.q_cap) {
        atomicOr(&G.ctr->overflow, (unsigned)PP_OVF_QUEUE);
        atomicAdd((unsigned long long *)&G.ctr->q_pending, (unsigned long long)(-1LL));
        return;
    }
    PPTask t;
    t.ev = ev; t.s = s; t.e = e; t.flags = ps - s;  // the windows before ps were scanned without a split
    *reinterpret_cast<int4 *>(&G.tasks[slot]) = *reinterpret_cast<int4 *>(&t);
    __threadfence();
    atomicExch(&G.ready[slot], 1);
}

// An event slot lost its last live window: the task is finished.
__device__ __forceinline__ void f_finish_event(K3FShared &S, int es)
{
    S.evs[es].ev = -1;
    S.tasks += 1;
    __threadfence();
    atomicAdd((unsigned long long *)&S.G.ctr->q_pending, (unsigned long long)(-1LL));
}

// Register window [it.ps, min(it.e, it.ps + W)) of interval `it` and push its pieces.  Lane 0, lock held.
__device__ __forceinline__ void f_register(K3FShared &S, int es, const K3Item &it, int screen)
{
    const K3Params &P = S.P;
    const int pe = k3_window_end(P, it);
    const int nch = (pe - it.ps - 2 * P.mw + 1 + 31) >> 5;
    if (S.n_free == 0) {
        f_push_global(S.G, S.evs[es].ev, it.s, it.e, it.ps);   // another CTA (or this one, later) continues it
        return;
    }
    const int w = S.free_list[--S.n_free];
    K3FWin &Wn = S.win[w];
    int np = nch <= K3F_MIN_PIECE ? 1 : (nch + K3F_MIN_PIECE - 1) / K3F_MIN_PIECE;
    if (np > K3F_MAXP) np = K3F_MAXP;
    const int pc = (nch + np - 1) / np;
    np = (nch + pc - 1) / pc;
    Wn.s = it.s; Wn.e = it.e; Wn.ps = it.ps;
    Wn.evs = (short)es;
    Wn.npieces = (unsigned char)np;
    Wn.exact = (screen && S.evs[es].ebase != K3_NO_EBASE) ? 0 : 1;
    Wn.piece_chunks = pc;
    Wn.done = 0;
    S.evs[es].live += 1;
    for (int p = np - 1; p >= 0; --p) S.stack[S.top++] = (w << 4) | p;
}

// The window loop of _recursive_split (cparsers.pyx:186-203) for interval `it` up to its next scannable window,
// which is registered; forced max_width splits on the way are emitted and their children handled the same way.
// Lane 0, lock held.
__device__ __noinline__ void f_place(K3FShared *Sp, int es, int screen, K3Item it)
{
    K3FShared &S = *Sp;
    const K3Global &G = S.G;
    const K3Params &P = S.P;
    const int mw = P.mw, MW = P.MW;
    const int64_t off = (int64_t)S.evs[es].off;
    K3Item st[4];
    int sp = 0;
    for (;;) {
        for (;;) {  // one interval
            const long long lim = (long long)it.e - 2LL * mw;
            if (it.ps >= lim) {
                if (it.e - it.s > MW) {
                    const int x = k3_forced(P, it.s, it.e);
                    k3_emit(G, off, x);
                    K3Item l, r;
                    l.s = it.s; l.e = x; l.ps = it.s; l.pad = 0;
                    r.s = x; r.e = it.e; r.ps = x; r.pad = 0;
                    const bool wl = k3_worth(P, l.s, l.e), wr = k3_worth(P, r.s, r.e);
                    if (wl && wr) {
                        if (sp < 4) st[sp++] = r; else f_push_global(G, S.evs[es].ev, r.s, r.e, r.s);
                        it = l;
                        continue;
                    }
                    if (wl) { it = l; continue; }
                    if (wr) { it = r; continue; }
                }
                break;
            }
            if (it.ps > (long long)it.s + MW) {
                const int x = k3_forced(P, it.s, it.e);
                k3_emit(G, off, x);  // the left part is not revisited (cparsers.pyx:189-191)
                if (!k3_worth(P, x, it.e)) break;
                it.s = x; it.ps = x;
                continue;
            }
            const int pe = k3_window_end(P, it);
            if (pe - it.ps <= 2 * mw) { it.ps = k3_next_ps(P, it.ps, it.e); continue; }
            f_register(S, es, it, screen);
            break;
        }
        if (sp == 0) break;
        it = st[--sp];
    }
}

// Apply the decision of a window scan (x = split position or -1) to its interval (cparsers.pyx:194-203), free the
// window record, and finish the event when this was its last live window.  Lane 0 takes the lock.
__device__ __forceinline__ void f_resolve(K3FShared &S, int w, int x, int screen, unsigned nexact)
{
    const K3Params &P = S.P;
    const K3FWin &Wn = S.win[w];
    K3Item it;
    it.s = Wn.s; it.e = Wn.e; it.ps = Wn.ps; it.pad = 0;
    const int es = Wn.evs;
    const int pe = k3_window_end(P, it);
    bool away = false;
    if (x >= 0 && K3_IDLE_SHARE > 0 && S.few && x - it.s >= K3_IDLE_SHARE && it.e - x >= K3_IDLE_SHARE) {
        // CTAs waiting for work hold tickets past the queue's tail: hand them the right child
        const unsigned long long qh = *((volatile unsigned long long *)&S.G.ctr->q_head);
        const unsigned long long qt = *((volatile unsigned long long *)&S.G.ctr->q_tail);
        away = (long long)(qh - qt) > 0;
    }
    f_lock(S);
    S.cand += (unsigned long long)(pe - it.ps - 2 * P.mw + 1);
    S.scans += 1;
    S.exact += nexact;
    K3Item c;
    c.pad = 0;
    if (x >= 0) {
        k3_emit(S.G, (int64_t)S.evs[es].off, x);
        // the right child is placed first: its pieces end up below the left child's on the stack (depth first, left to right)
        if (k3_worth(P, x, it.e)) {
            if (away) f_push_global(S.G, S.evs[es].ev, x, it.e, x);
            else { c.s = x; c.e = it.e; c.ps = x; f_place(&S, es, screen, c); }
        }
        if (k3_worth(P, it.s, x)) { c.s = it.s; c.e = x; c.ps = it.s; f_place(&S, es, screen, c); }
    } else {
        c.s = it.s; c.e = it.e; c.ps = k3_next_ps(P, it.ps, it.e);
        f_place(&S, es, screen, c);
    }
    S.free_list[S.n_free++] = w;
    if (--S.evs[es].live == 0) f_finish_event(S, es);
    f_unlock(S);
}

// Take the next task of the global queue and place its first window.  Whole warp; lane 0 acts.
__device__ __forceinline__ void f_fetch(K3FShared &S, int screen)
{
    const K3Global &G = S.G;
    if ((threadIdx.x & 31) != 0) return;
    const unsigned long long h = atomicAdd(&G.ctr->q_head, 1ull);
    bool ok = false;
    for (long long spin = 0;; ++spin) {
        if ((int64_t)h < G.q_cap && *((volatile int *)&G.ready[h]) != 0) { ok = true; break; }
        if (*((volatile long long *)&G.ctr->q_pending) <= 0) break;
        if (spin > K3F_SPIN_LIMIT) { atomicOr(&G.ctr->overflow, (unsigned)PP_OVF_QUEUE); S.failed = 1; break; }
        __nanosleep(256);
    }
    if (!ok) {
        f_lock(S);
        S.drained = 1;
        S.fetching = 0;
        f_unlock(S);
        return;
    }
    __threadfence();
    const int4 t = __ldcg(reinterpret_cast<const int4 *>(&G.tasks[h]));
    const long long off = (long long)G.ev_off[t.x];
    // screening exponent base of the task: the variance of its whole interval (any value within 2^+-127 of the
    // candidates' variances will do; the validity test checks each candidate)
    int ebase = K3_NO_EBASE;
    if (t.z > t.y) {
        K3GlobalCC a;
        a.g = G.cc + off;
        int eb = 0;
        if (screen && k3_window_ebase(k3_var(a.at(t.z - 1), a.at(t.y - 1), t.z - t.y), eb)) ebase = eb;
    }
    f_lock(S);
    int es = 0;
    while (es < K3F_EVS && S.evs[es].ev >= 0) ++es;   // a slot is free: the fetch was granted on that condition
    S.evs[es].ev = t.x;
    S.evs[es].off = off;
    S.evs[es].ebase = ebase;
    S.evs[es].live = 1;
    S.fetching = 0;
    if (k3_worth(S.P, t.y, t.z)) {
        K3Item it;
        it.s = t.y; it.e = t.z; it.ps = t.y + t.w; it.pad = 0;  // k3_spine hands over a remainder with its window position
        f_place(&S, es, screen, it);
    }
    if (--S.evs[es].live == 0) f_finish_event(S, es);
    f_unlock(S);
}

// Decide a screened window from its pieces' summaries (whole warp, uniform result): split position or -1.
// K1 / K2 / I1 hold, in lane p < np, the summary of piece p (other lanes: K3_NOKEY / -1).
__device__ __forceinline__ int f_decide(const K3Global &G, const K3Params &P, const K3FWin &Wn, const double2 *ccg,
                                        const double2 lo, const double2 hi, int ps, int pe, int ebase, int np,
                                        unsigned long long K1, unsigned long long K2, int I1, unsigned &nexact)
{
    const int lane = threadIdx.x & 31;
    const int mw = P.mw;
    K3GlobalCC acc;
    acc.g = ccg;
    const bool anybad = __any_sync(PP_FULL, lane < np && (I1 & K3_BAD_FLAG));
    const unsigned long long gmin = k3_warp_min_u64(K1);
    if (anybad || gmin == K3_NOKEY) {
        // a candidate failed the validity test: the whole window in the reference's arithmetic
        K3Best b = k3_scan_range(acc, ps, pe, mw, P.min_gain, lane, 32);
        b = k3_warp_reduce(b);
        nexact += (unsigned)(pe - ps - 2 * mw + 1);
        return b.x;
    }
    const unsigned long long eps2 = k3_eps2_key(pe - ps);
    const unsigned long long thr = gmin + eps2;
    const unsigned rescan = __ballot_sync(PP_FULL, lane < np && K2 <= thr);
    const unsigned single = __ballot_sync(PP_FULL, lane < np && K2 > thr && K1 <= thr);
    const double tot = k3_exact_tot(lo, hi, ps, pe);   // cheap next to a scan; only used on the exact paths
    if (!rescan && __popc(single) == 1) {
        const int i_one = ps + (__shfl_sync(PP_FULL, I1, __ffs(single) - 1) & 0x1fffffff);
        unsigned long long kt = 0ull;
        const unsigned nw = (unsigned)(pe - ps);
        const bool tok = k3_side(__dsub_rn(hi.x, lo.x), __dsub_rn(hi.y, lo.y), __ldg(G.RN + nw), nw, ebase, kt);
        if (tok) {
            // i_one is the argmax (every other candidate is more than 2 eps worse); its screened gain
            // (kt - gmin) ln2 / 2^23 is within 2 eps (+ conversions) of the reference's
            const double d = (double)(long long)(kt - gmin);
            const double want = P.min_gain * K3_KEY_PER_NAT;
            const double margin = (double)eps2 + 64.0;
            if (d > want + margin) return i_one;
            if (d < want - margin) return -1;
        }
        const double g = k3_exact_gain(lo, acc.at(i_one - 1), hi, ps, pe, i_one, tot);
        nexact += 1;
        return g > P.min_gain ? i_one : -1;
    }
    // several contenders: their exact gains, largest wins, lowest index on ties
    K3Best b;
    b.g = P.min_gain;
    b.x = -1;
    unsigned cnt = 0;
    if (single & (1u << lane)) {
        const int i = ps + (I1 & 0x1fffffff);
        const double g = k3_exact_gain(lo, acc.at(i - 1), hi, ps, pe, i, tot);
        ++cnt;
        if (g > b.g) { b.g = g; b.x = i; }
    }
    const int w_last = pe - mw;
    for (unsigned m = rescan; m; m &= m - 1) {
        const int p = __ffs(m) - 1;
        const int c_lo = p * Wn.piece_chunks, c_hi = c_lo + Wn.piece_chunks;
        int i_end = ps + mw + c_hi * 32 - 1;
        i_end = i_end < w_last ? i_end : w_last;
        K3Best r;
        r.g = P.min_gain;
        r.x = -1;
        for (int i = ps + mw + c_lo * 32 + lane; i <= i_end; i += 32) {
            unsigned long long key;
            k3_screen_key_at(ccg, lo, hi, ps, pe, i, G.RN, ebase, key);
            if (key <= thr) {
                const double g = k3_exact_gain(lo, acc.at(i - 1), hi, ps, pe, i, tot);
                ++cnt;
                if (g > r.g) { r.g = g; r.x = i; }
            }
        }
        b = k3_better(b, r);
    }
    b = k3_warp_reduce(b);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(PP_FULL, cnt, d);
    nexact += cnt;
    return b.x;
}

__global__ void __launch_bounds__(K3F_THREADS, K3F_CTAS_PER_SM) k3_flow(K3Global G, K3Params P)
{
    extern __shared__ __align__(16) unsigned char k3f_smem[];
    K3FShared &S = *reinterpret_cast<K3FShared *>(k3f_smem);
    const int tid = threadIdx.x, lane = tid & 31;
    const int mw = P.mw;
    const int screen = (G.screen && mw >= 1 && P.W <= K3_MAX_SCREEN_W) ? 1 : 0;

    if (tid == 0) {
        S.G = G;
        S.P = P;
        S.few = ((long long)G.ctr->n_events - (long long)G.ctr->ev_begin) < 2LL * gridDim.x ? 1 : 0;
        S.n_free = K3F_WINS;
        S.top = 0; S.lock = 0; S.fetching = 0; S.drained = 0; S.failed = 0;
        S.cand = S.scans = S.exact = S.tasks = 0ull;
        for (int e = 0; e < K3F_EVS; ++e) { S.evs[e].ev = -1; S.evs[e].live = 0; }
    }
    for (int k = tid; k < K3F_WINS; k += K3F_THREADS) S.free_list[k] = K3F_WINS - 1 - k;
    if (tid == 0 && blockIdx.x == 0) G.ctr->n_long = 0ull;  // k3_spine is done with it; ready for the next search
    __syncthreads();

    long long idle = 0;
    for (;;) {
        // ---- find something to do ---------------------------------------------------------------
        int action = 0, entry = 0;   // 0: nothing right now, 1: a piece, 2: fetch a task, 3: leave
        if (lane == 0) {
            f_lock(S);
            if (S.top > 0) {
                entry = S.stack[--S.top];
                action = 1;
            } else if (S.failed) {
                action = 3;
            } else {
                int free_ev = 0, busy_ev = 0;
                for (int e = 0; e < K3F_EVS; ++e) {
                    free_ev += S.evs[e].ev < 0;
                    busy_ev += S.evs[e].ev >= 0;
                }
                if (S.drained) action = (busy_ev == 0 && !S.fetching) ? 3 : 0;
                else if (!S.fetching && free_ev > 0 && S.n_free >= 4) { S.fetching = 1; action = 2; }
            }
            f_unlock(S);
        }
        action = __shfl_sync(PP_FULL, action, 0);
        entry = __shfl_sync(PP_FULL, entry, 0);
        if (action == 3) break;
        if (action == 0) {
            if (++idle > K3F_SPIN_LIMIT) {   // never expected: leave instead of hanging the device
                if (lane == 0) { atomicOr(&G.ctr->overflow, (unsigned)PP_OVF_QUEUE); S.failed = 1; }
                break;
            }
            __nanosleep(64);
            continue;
        }
        idle = 0;
        if (action == 2) {
            f_fetch(S, screen);
            __syncwarp();
            continue;
        }

        // ---- one piece -----------------------------------------------------------------------------
        const int w = entry >> 4, p = entry & 15;
        K3FWin &Wn = S.win[w];
        K3Item it;
        it.s = Wn.s; it.e = Wn.e; it.ps = Wn.ps; it.pad = 0;
        const int es = Wn.evs, np = Wn.npieces, pc = Wn.piece_chunks;
        const bool exact = Wn.exact != 0;
        const int ebase = S.evs[es].ebase;
        const double2 *ccg = G.cc + S.evs[es].off;
        K3GlobalCC acc;
        acc.g = ccg;
        const int pe = k3_window_end(P, it);
        const int w_last = pe - mw;
        int i_end = it.ps + mw + (p + 1) * pc * 32 - 1;
        i_end = i_end < w_last ? i_end : w_last;
        const int i0 = it.ps + mw + p * pc * 32 + lane;
        const double2 lo = acc.at(it.ps - 1), hi = acc.at(pe - 1);

        unsigned long long K1 = K3_NOKEY, K2 = K3_NOKEY;
        int I1 = -1;
        if (!exact) {
            K3Scr a;
            k3_scr_init(a);
            k3_screen_lane(ccg, lo, hi, it.ps, pe, ebase, G.RN, i0, i_end, 32, a);
            bool bad;
            k3_warp_summary(a, K1, K2, I1, bad);
            I1 = ((I1 - it.ps) & 0x1fffffff) | (bad ? K3_BAD_FLAG : 0);
        } else {
            // the reference's arithmetic for every candidate of the piece (_best_split_stepwise, cparsers.pyx:171-177)
            const double tot = k3_exact_tot(lo, hi, it.ps, pe);
            K3Best b;
            b.g = P.min_gain;
            b.x = -1;
            for (int i = i0; i <= i_end; i += 32) {
                const double g = k3_exact_gain(lo, acc.at(i - 1), hi, it.ps, pe, i, tot);
                if (g > b.g) { b.g = g; b.x = i; }
            }
            b = k3_warp_reduce(b);
            K1 = (unsigned long long)__double_as_longlong(b.g);
            I1 = b.x;
        }
        bool last = true;
        if (np > 1) {
            int old = 0;
            if (lane == 0) {
                Wn.k1[p] = K1; Wn.k2[p] = K2; Wn.i1[p] = I1;
                __threadfence_block();
                old = atomicAdd(&Wn.done, 1);
                __threadfence_block();
            }
            old = __shfl_sync(PP_FULL, old, 0);
            last = old == np - 1;
            if (last) {   // this warp decides: lane q takes piece q's summary
                const volatile K3FWin &V = Wn;
                K1 = lane < np ? V.k1[lane] : K3_NOKEY;
                K2 = lane < np ? V.k2[lane] : K3_NOKEY;
                I1 = lane < np ? V.i1[lane] : -1;
            }
        } else if (lane != 0) {   // one piece: its summary sits in lane 0, like piece 0 of a shared window
            K1 = K2 = K3_NOKEY;
            I1 = exact ? I1 : -1;
        }
        if (!last) continue;

        // ---- the window is complete: decide, then the recursion's bookkeeping --------------------------
        int x;
        unsigned nexact = 0;
        if (!exact) {
            x = f_decide(G, P, Wn, ccg, lo, hi, it.ps, pe, ebase, np, K1, K2, I1, nexact);
        } else {
            K3Best b;
            b.g = __longlong_as_double((long long)K1);
            b.x = I1;
            if (np > 1) {
                if (lane >= np) { b.g = P.min_gain; b.x = -1; }
                b = k3_warp_reduce(b);
            }
            x = __shfl_sync(PP_FULL, b.x, 0);
            if (screen) nexact = (unsigned)(pe - it.ps - 2 * mw + 1);
        }
        if (lane == 0) f_resolve(S, w, x, screen, nexact);
        __syncwarp();
    }
    __syncthreads();
    if (tid == 0) {
        atomicAdd(&G.ctr->n_cand, S.cand);
        atomicAdd(&G.ctr->n_scan, S.scans);
        atomicAdd(&G.ctr->n_exact, S.exact);
        atomicAdd(&G.ctr->n_tasks, S.tasks);
    }
}
