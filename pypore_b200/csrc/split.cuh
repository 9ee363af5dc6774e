// split.cuh -- K3: the recursive maximum-likelihood changepoint search of
// FastStatSplit (PyPore/cparsers.pyx:26-38 var_c, :157-178 _best_split_stepwise,
// :180-203 _recursive_split) as a persistent device work queue of intervals.
//
// The recursion's result depends only on the interval [s,e) and the event's
// prefix sums, so intervals can be processed in any order.
// A task is an interval of one event.  A CTA that pops a task
//   * walks the "right spine" of a long interval window by window (the
//     sequential dependency of App. A.3), scanning each window CTA-wide, pushing
//     the left child to the global queue for another CTA and continuing with
//     the right child itself; and
//   * once the interval fits in shared memory (<= K3_CAP samples), stages its
//     slab of {c, c2} once and resolves the whole subtree locally, level by
//     level, with big windows scanned by the whole CTA and small windows one
//     per warp.
// Breakpoints are recorded as bits in flat event space; a compaction pass turns
// the bitmap into the sorted segment table, so no ordering is needed here.
//
// Arithmetic contract (bit-exact split decisions): every operation of var_c and
// of the gain is a separate IEEE fp64 operation in the reference's association
// (__dsub_rn/__ddiv_rn/__dmul_rn/__dadd_rn are never contracted to FMA), the
// comparison is a strict '>' against a running maximum seeded with min_gain,
// and ties go to the lowest index.
#pragma once
#include "common.cuh"

constexpr int K3_THREADS = 512;
constexpr int K3_WARPS = K3_THREADS / 32;
constexpr int K3_CAP = 12288;   // samples of one interval staged in shared memory
constexpr int K3_LIST = 1024;   // items per level list
constexpr int K3_BIG = 1024;    // candidates from which a window is scanned CTA-wide

struct PPTask { int ev, s, e, flags; };
struct K3Item { int s, e, ps; };
struct K3Params { int mw, MW, W; double min_gain; };
struct K3Best { double g; int x; };

struct K3Global {
    const double2 *cc;
    const int64_t *ev_off;
    const int64_t *ev_len;
    unsigned *bits;
    PPTask *tasks;
    int *ready;
    int64_t q_cap;
    PPCounters *ctr;
};

struct K3Shared {
    K3Item list[2][K3_LIST];
    int scanlist[K3_LIST];
    double red_g[K3_WARPS];
    int red_x[K3_WARPS];
    int nA, nB, nbig, nsmall;
    PPTask task;
    int have_task;
    unsigned long long cand, scans;
};

constexpr size_t K3_SMEM_CC = sizeof(double2) * (K3_CAP + 1);
constexpr size_t K3_SMEM_BYTES = K3_SMEM_CC + sizeof(K3Shared);

// prefix-sum accessors: at(p) = {c[p], c2[p]} with c[-1] = c2[-1] = 0
struct K3SmemCC {
    const double2 *sm;
    int S0;  // sm[k] holds position S0 - 1 + k
    __device__ __forceinline__ double2 at(int p) const { return sm[p - S0 + 1]; }
};
struct K3GlobalCC {
    const double2 *g;  // event base
    __device__ __forceinline__ double2 at(int p) const
    {
        if (p < 0) return make_double2(0.0, 0.0);
        return __ldcg(g + p);
    }
};

// var_c (cparsers.pyx:31-38); start == 0 subtracts an exact 0.0, bit-identical
// to the reference's special case.
__device__ __forceinline__ double k3_var(const double2 hi, const double2 lo, int cnt)
{
    if (cnt == 0) return 0.0;
    const double n = (double)cnt;
    const double m = __ddiv_rn(__dsub_rn(hi.x, lo.x), n);
    return __dsub_rn(__ddiv_rn(__dsub_rn(hi.y, lo.y), n), __dmul_rn(m, m));
}

__device__ __forceinline__ K3Best k3_better(const K3Best a, const K3Best b)
{
    if (b.x >= 0 && (a.x < 0 || b.g > a.g || (b.g == a.g && b.x < a.x))) return b;
    return a;
}

// Candidates ps+mw+first, +stride, ... <= pe-mw of window [ps,pe)
// (_best_split_stepwise loop, cparsers.pyx:171-177).
template <class CC>
__device__ __forceinline__ K3Best k3_scan_range(const CC &cc, int ps, int pe, int mw,
                                                double min_gain, int first, int stride)
{
    const double2 lo = cc.at(ps - 1), hi = cc.at(pe - 1);
    const double tot = __dmul_rn((double)(pe - ps), log(k3_var(hi, lo, pe - ps)));
    K3Best b;
    b.g = min_gain;
    b.x = -1;
    const int last = pe - mw;
    for (int i = ps + mw + first; i <= last; i += stride) {
        const double2 mid = cc.at(i - 1);
        const double low = __dmul_rn((double)(i - ps), log(k3_var(mid, lo, i - ps)));
        const double high = __dmul_rn((double)(pe - i), log(k3_var(hi, mid, pe - i)));
        const double g = __dsub_rn(tot, __dadd_rn(low, high));
        if (g > b.g) { b.g = g; b.x = i; }
    }
    return b;
}

__device__ __forceinline__ K3Best k3_warp_reduce(K3Best b)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        K3Best o;
        o.g = __shfl_xor_sync(PP_FULL, b.g, d);
        o.x = __shfl_xor_sync(PP_FULL, b.x, d);
        b = k3_better(b, o);
    }
    return b;
}

template <class CC>
__device__ __forceinline__ K3Best k3_cta_scan(const CC &cc, int ps, int pe, const K3Params &P,
                                              K3Shared &S)
{
    const int tid = threadIdx.x;
    K3Best b = k3_scan_range(cc, ps, pe, P.mw, P.min_gain, tid, K3_THREADS);
    b = k3_warp_reduce(b);
    __syncthreads();
    if ((tid & 31) == 0) { S.red_g[tid >> 5] = b.g; S.red_x[tid >> 5] = b.x; }
    __syncthreads();
    K3Best r;
    r.g = S.red_g[0];
    r.x = S.red_x[0];
#pragma unroll
    for (int w = 1; w < K3_WARPS; ++w) {
        K3Best o;
        o.g = S.red_g[w];
        o.x = S.red_x[w];
        r = k3_better(r, o);
    }
    return r;
}

__device__ __forceinline__ void k3_emit(const K3Global &G, int64_t off, int x)
{
    const int64_t f = off + x;
    atomicOr(&G.bits[f >> 5], 1u << (unsigned)(f & 31));
}

__device__ __forceinline__ bool k3_worth(const K3Params &P, int s, int e)
{
    // an interval that can neither be scanned nor force-split is a leaf
    return ((long long)e - s > 2LL * P.mw) || (e - s > P.MW);
}

__device__ void k3_push_global(const K3Global &G, int ev, int s, int e)
{
    atomicAdd((unsigned long long *)&G.ctr->q_pending, 1ull);
    const unsigned long long slot = atomicAdd(&G.ctr->q_tail, 1ull);
    if ((int64_t)slot >= G.q_cap) {
        atomicOr(&G.ctr->overflow, (unsigned)PP_OVF_QUEUE);
        atomicAdd((unsigned long long *)&G.ctr->q_pending, (unsigned long long)(-1LL));
        return;
    }
    PPTask t;
    t.ev = ev; t.s = s; t.e = e; t.flags = 0;
    *reinterpret_cast<int4 *>(&G.tasks[slot]) = *reinterpret_cast<int4 *>(&t);
    __threadfence();
    atomicExch(&G.ready[slot], 1);
}

__device__ __forceinline__ void k3_push_local(const K3Global &G, K3Shared &S, K3Item *next, int ev,
                                              int s, int e, int ps)
{
    const int idx = atomicAdd(&S.nB, 1);
    if (idx < K3_LIST) {
        K3Item it;
        it.s = s; it.e = e; it.ps = ps;
        next[idx] = it;
    } else {
        k3_push_global(G, ev, s, e);  // restart at ps = s elsewhere: redundant scans, same result
    }
}

__device__ __forceinline__ int k3_forced(const K3Params &P, int s, int e)
{
    const long long a = (long long)s + P.MW, b = (long long)e - P.mw;
    return (int)(a < b ? a : b);
}

__device__ __forceinline__ int k3_next_ps(const K3Params &P, int ps, int e)
{
    const long long n = (long long)ps + P.W / 2;
    return (int)(n < e ? n : e);
}

// Apply a scan result to an item in local mode (cparsers.pyx:194-203).
__device__ __forceinline__ void k3_resolve_local(const K3Global &G, K3Shared &S, K3Item *next,
                                                 const K3Params &P, int ev, int64_t off,
                                                 const K3Item it, int x)
{
    if (x >= 0) {
        k3_emit(G, off, x);
        if (k3_worth(P, it.s, x)) k3_push_local(G, S, next, ev, it.s, x, it.s);
        if (k3_worth(P, x, it.e)) k3_push_local(G, S, next, ev, x, it.e, x);
    } else {
        k3_push_local(G, S, next, ev, it.s, it.e, k3_next_ps(P, it.ps, it.e));
    }
}

__global__ void __launch_bounds__(K3_THREADS, 1) k3_split(K3Global G, K3Params P)
{
    extern __shared__ __align__(16) unsigned char k3_smem[];
    double2 *sm_cc = reinterpret_cast<double2 *>(k3_smem);
    K3Shared &S = *reinterpret_cast<K3Shared *>(k3_smem + K3_SMEM_CC);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mw = P.mw, MW = P.MW, W = P.W;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const unsigned long long h = atomicAdd(&G.ctr->q_head, 1ull);
            int ok = 0;
            for (;;) {
                if ((int64_t)h < G.q_cap && *((volatile int *)&G.ready[h]) != 0) {
                    __threadfence();
                    const int4 t = __ldcg(reinterpret_cast<const int4 *>(&G.tasks[h]));
                    S.task.ev = t.x; S.task.s = t.y; S.task.e = t.z; S.task.flags = t.w;
                    ok = 1;
                    break;
                }
                if (*((volatile long long *)&G.ctr->q_pending) <= 0) break;
                __nanosleep(256);
            }
            S.have_task = ok;
            S.cand = 0;
            S.scans = 0;
        }
        __syncthreads();
        if (!S.have_task) break;
        const int ev = S.task.ev;
        int s = S.task.s;
        const int e = S.task.e;
        const int64_t off = G.ev_off[ev];
        const double2 *ccg = G.cc + off;

        // ---- spine mode: interval too long for shared memory -------------------
        int ps = s;
        bool done = false;
        while ((long long)e - s > K3_CAP) {
            const long long lim = (long long)e - 2LL * mw;
            if (ps >= lim) {
                if (e - s <= MW) { done = true; break; }
                const int x = k3_forced(P, s, e);
                if (tid == 0) {
                    k3_emit(G, off, x);
                    if (k3_worth(P, s, x)) k3_push_global(G, ev, s, x);
                }
                s = x; ps = s;
                continue;
            }
            if (ps > (long long)s + MW) {
                const int x = k3_forced(P, s, e);
                if (tid == 0) k3_emit(G, off, x);
                s = x; ps = s;  // the left part is not revisited (cparsers.pyx:189-191)
                continue;
            }
            const long long pe_l = (long long)ps + W;
            const int pe = (int)(pe_l < e ? pe_l : e);
            if (pe - ps <= 2 * mw) { ps = k3_next_ps(P, ps, e); continue; }
            K3Best b;
            if (pe - ps <= K3_CAP) {
                __syncthreads();
                for (int k = tid; k <= pe - ps; k += K3_THREADS) {
                    const int p = ps - 1 + k;
                    sm_cc[k] = p < 0 ? make_double2(0.0, 0.0) : __ldcg(ccg + p);
                }
                __syncthreads();
                K3SmemCC acc;
                acc.sm = sm_cc; acc.S0 = ps;
                b = k3_cta_scan(acc, ps, pe, P, S);
            } else {
                K3GlobalCC acc;
                acc.g = ccg;
                b = k3_cta_scan(acc, ps, pe, P, S);
            }
            if (tid == 0) { S.cand += (unsigned long long)(pe - ps - 2 * mw + 1); S.scans += 1; }
            if (b.x >= 0) {
                if (tid == 0) {
                    k3_emit(G, off, b.x);
                    if (k3_worth(P, s, b.x)) k3_push_global(G, ev, s, b.x);
                }
                s = b.x; ps = s;
            } else {
                ps = k3_next_ps(P, ps, e);
            }
        }

        // ---- local mode: whole subtree from one staged slab --------------------
        if (!done) {
            __syncthreads();
            for (int k = tid; k <= e - s; k += K3_THREADS) {
                const int p = s - 1 + k;
                sm_cc[k] = p < 0 ? make_double2(0.0, 0.0) : __ldcg(ccg + p);
            }
            if (tid == 0) {
                K3Item it;
                it.s = s; it.e = e; it.ps = ps;
                S.list[0][0] = it;
                S.nA = 1;
            }
            K3SmemCC acc;
            acc.sm = sm_cc; acc.S0 = s;
            int cur = 0;
            __syncthreads();
            for (;;) {
                const int nA = S.nA;
                if (nA == 0) break;
                K3Item *A = S.list[cur], *Bn = S.list[cur ^ 1];
                __syncthreads();  // everyone has read nA
                if (tid == 0) { S.nB = 0; S.nbig = 0; S.nsmall = 0; }
                __syncthreads();
                // step 1: the window-loop bookkeeping of _recursive_split per item
                for (int t = tid; t < nA; t += K3_THREADS) {
                    const K3Item it = A[t];
                    const long long lim = (long long)it.e - 2LL * mw;
                    if (it.ps >= lim) {
                        if (it.e - it.s > MW) {
                            const int x = k3_forced(P, it.s, it.e);
                            k3_emit(G, off, x);
                            if (k3_worth(P, it.s, x)) k3_push_local(G, S, Bn, ev, it.s, x, it.s);
                            if (k3_worth(P, x, it.e)) k3_push_local(G, S, Bn, ev, x, it.e, x);
                        }
                    } else if (it.ps > (long long)it.s + MW) {
                        const int x = k3_forced(P, it.s, it.e);
                        k3_emit(G, off, x);
                        if (k3_worth(P, x, it.e)) k3_push_local(G, S, Bn, ev, x, it.e, x);
                    } else {
                        const long long pe_l = (long long)it.ps + W;
                        const int pe = (int)(pe_l < it.e ? pe_l : it.e);
                        const int ncand = pe - it.ps - 2 * mw + 1;
                        if (pe - it.ps <= 2 * mw) {
                            k3_push_local(G, S, Bn, ev, it.s, it.e, k3_next_ps(P, it.ps, it.e));
                        } else if (ncand >= K3_BIG) {
                            S.scanlist[atomicAdd(&S.nbig, 1)] = t;
                        } else {
                            S.scanlist[K3_LIST - 1 - atomicAdd(&S.nsmall, 1)] = t;
                        }
                    }
                }
                __syncthreads();
                const int nbig = S.nbig, nsmall = S.nsmall;
                // step 2: big windows, one at a time, whole CTA
                for (int bi = 0; bi < nbig; ++bi) {
                    const K3Item it = A[S.scanlist[bi]];
                    const long long pe_l = (long long)it.ps + W;
                    const int pe = (int)(pe_l < it.e ? pe_l : it.e);
                    const K3Best b = k3_cta_scan(acc, it.ps, pe, P, S);
                    if (tid == 0) {
                        atomicAdd(&S.cand, (unsigned long long)(pe - it.ps - 2 * mw + 1));
                        atomicAdd(&S.scans, 1ull);
                        k3_resolve_local(G, S, Bn, P, ev, off, it, b.x);
                    }
                }
                // step 3: small windows, one per warp
                for (int k = warp; k < nsmall; k += K3_WARPS) {
                    const K3Item it = A[S.scanlist[K3_LIST - 1 - k]];
                    const long long pe_l = (long long)it.ps + W;
                    const int pe = (int)(pe_l < it.e ? pe_l : it.e);
                    K3Best b = k3_scan_range(acc, it.ps, pe, mw, P.min_gain, lane, 32);
                    b = k3_warp_reduce(b);
                    if (lane == 0) {
                        atomicAdd(&S.cand, (unsigned long long)(pe - it.ps - 2 * mw + 1));
                        atomicAdd(&S.scans, 1ull);
                        k3_resolve_local(G, S, Bn, P, ev, off, it, b.x);
                    }
                }
                __syncthreads();
                if (tid == 0) S.nA = S.nB < K3_LIST ? S.nB : K3_LIST;
                cur ^= 1;
                __syncthreads();
            }
        }
        __syncthreads();
        if (tid == 0) {
            atomicAdd(&G.ctr->n_cand, S.cand);
            atomicAdd(&G.ctr->n_scan, S.scans);
            atomicAdd(&G.ctr->n_tasks, 1ull);
            __threadfence();
            atomicAdd((unsigned long long *)&G.ctr->q_pending, (unsigned long long)(-1LL));
        }
    }
}

// One initial task per event; event starts are segment starts.
__global__ void __launch_bounds__(256)
k3_init_queue(K3Global G)
{
    const int64_t n_events = (int64_t)G.ctr->n_events;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_events;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t off = G.ev_off[e];
        const int64_t len = G.ev_len[e];
        atomicOr(&G.bits[off >> 5], 1u << (unsigned)(off & 31));
        if (e < G.q_cap && len < 0x7fffffffLL) {
            PPTask t;
            t.ev = (int)e; t.s = 0; t.e = (int)len; t.flags = 0;
            G.tasks[e] = t;
            G.ready[e] = 1;
        } else {
            atomicOr(&G.ctr->overflow, (unsigned)PP_OVF_QUEUE);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t q = n_events < G.q_cap ? n_events : G.q_cap;
        G.ctr->q_head = 0;
        G.ctr->q_tail = (unsigned long long)q;
        G.ctr->q_pending = (long long)q;
        G.ctr->n_cand = 0;
        G.ctr->n_scan = 0;
        G.ctr->n_tasks = 0;
    }
}

// ---------------------------------------------------------------------------
// Bitmap -> sorted segment table (flat start, event id, event-relative start).
// ---------------------------------------------------------------------------
constexpr int CP_THREADS = 256;
constexpr int CP_WPT = 4;                           // words per thread
constexpr int CP_BLOCK_WORDS = CP_THREADS * CP_WPT;  // 1024 words = 32768 samples

__global__ void __launch_bounds__(CP_THREADS)
k3c_count(const unsigned *__restrict__ bits, int64_t n_words, unsigned *__restrict__ block_count)
{
    __shared__ unsigned wsum[CP_THREADS / 32];
    const int tid = threadIdx.x;
    const int64_t w0 = (int64_t)blockIdx.x * CP_BLOCK_WORDS + tid * CP_WPT;
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < CP_WPT; ++k)
        if (w0 + k < n_words) c += __popc(bits[w0 + k]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(PP_FULL, c, d);
    if ((tid & 31) == 0) wsum[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
        unsigned t = 0;
        for (int w = 0; w < CP_THREADS / 32; ++w) t += wsum[w];
        block_count[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
k3c_scan(const unsigned *__restrict__ block_count, int64_t n_blocks,
         unsigned long long *__restrict__ block_off, PPCounters *ctr, int64_t cap_segs)
{
    __shared__ unsigned long long wtot[32];
    __shared__ unsigned long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < n_blocks; c0 += 1024) {
        const int64_t b = c0 + tid;
        const unsigned long long v = b < n_blocks ? block_count[b] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(PP_FULL, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        unsigned long long add = 0;
        for (int w = 0; w < warp; ++w) add += wtot[w];
        const unsigned long long excl = s_carry + add + inc - v;
        if (b < n_blocks) block_off[b] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        ctr->n_segments = s_carry;
        if ((int64_t)s_carry > cap_segs) atomicOr(&ctr->overflow, (unsigned)PP_OVF_SEGS);
    }
}

__global__ void __launch_bounds__(CP_THREADS)
k3c_write(const unsigned *__restrict__ bits, int64_t n_words,
          const unsigned long long *__restrict__ block_off, const int64_t *__restrict__ ev_off,
          const PPCounters *ctr, int64_t *__restrict__ seg_flat, int *__restrict__ seg_event,
          int64_t *__restrict__ seg_start, int64_t cap_segs)
{
    __shared__ unsigned wsum[CP_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t w0 = (int64_t)blockIdx.x * CP_BLOCK_WORDS + tid * CP_WPT;
    unsigned wv[CP_WPT];
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < CP_WPT; ++k) {
        wv[k] = (w0 + k < n_words) ? bits[w0 + k] : 0u;
        c += __popc(wv[k]);
    }
    unsigned inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned t = __shfl_up_sync(PP_FULL, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned add = 0;
    for (int w = 0; w < warp; ++w) add += wsum[w];
    if (c == 0) return;
    unsigned long long k = block_off[blockIdx.x] + add + inc - c;
    const int64_t n_events = (int64_t)ctr->n_events;
    int64_t ev = -1, ev_end = 0, ev_begin = 0;
#pragma unroll
    for (int q = 0; q < CP_WPT; ++q) {
        unsigned w = wv[q];
        while (w) {
            const int bit = __ffs(w) - 1;
            w &= w - 1;
            const int64_t f = (w0 + q) * 32 + bit;
            if (ev < 0) {
                ev = pp_upper_index(ev_off, n_events, f);
                ev_begin = ev_off[ev];
                ev_end = ev_off[ev + 1];
            }
            while (f >= ev_end) { ++ev; ev_begin = ev_end; ev_end = ev_off[ev + 1]; }
            if ((int64_t)k < cap_segs) {
                seg_flat[k] = f;
                seg_event[k] = (int)ev;
                seg_start[k] = f - ev_begin;
            }
            ++k;
        }
    }
}

__global__ void __launch_bounds__(256)
k3c_ends(const PPCounters *ctr, const int64_t *__restrict__ seg_flat,
         const int *__restrict__ seg_event, const int64_t *__restrict__ ev_off,
         int64_t *__restrict__ seg_end, int64_t cap_segs)
{
    int64_t S = (int64_t)ctr->n_segments;
    if (S > cap_segs) S = cap_segs;
    const int64_t total = (int64_t)ctr->n_event_samples;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < S;
         k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t fe = (k + 1 < S) ? seg_flat[k + 1] : total;
        seg_end[k] = fe - ev_off[seg_event[k]];
    }
}
