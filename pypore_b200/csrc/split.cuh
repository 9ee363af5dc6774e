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
// Two-stage evaluation of a window (bit-exact result, ~6x fewer instructions):
//   SCREEN  every candidate i gets H~(i) ~= low(i) + high(i) from a division-free
//           formulation  n1*log(D1) + n2*log(D2) - T[n1] - T[n2],
//           D = Q*n - S*S = n^2 * V,  T[n] = 2 n ln n  (table), with a table-driven
//           fp64 log (|err| < 1e-11).  For candidates that pass the validity test
//           (variance not smaller than 2^-24 of the mean square) the distance to
//           the reference's own rounded value is rigorously bounded by
//           eps = n_window * 3e-8 + 1e-6  (derivation in DESIGN.md).
//   EXACT   only candidates with H~ <= min H~ + 2 eps, and every candidate that
//           failed the validity test, are evaluated with the reference's exact
//           arithmetic below; the decision (strict '>' against min_gain, lowest
//           index on ties, NaN never wins) is taken on exact values only.
//
// Exact arithmetic contract: every operation of var_c and of the gain is a
// separate IEEE fp64 operation in the reference's association
// (__dsub_rn/__ddiv_rn/__dmul_rn/__dadd_rn are never contracted to FMA).
#pragma once
#include "common.cuh"

constexpr int K3_THREADS = 128;
constexpr int K3_CTAS_PER_SM = 6;
constexpr int K3_WARPS = K3_THREADS / 32;
constexpr int K3_CAP = 10240;   // longest interval resolved level by level inside one CTA
constexpr int K3_LIST = 256;    // items per level list
constexpr int K3_REQ = 256;     // exact-evaluation requests per level
constexpr int K3_CHUNKS = 640;  // 32-candidate chunks per level (<= K3_CAP/32 + K3_LIST)
constexpr int K3_FULL_FLAG = 0x40000000;  // window entry: screening inconclusive / list overflow, scan exactly
constexpr int K3_AMB_FLAG = (int)0x80000000;

constexpr double K3_RATIO_MAX = 16777216.0;  // 2^24: validity bound on mean-square / variance
constexpr double K3_EPS_PER_SAMPLE = 3e-8;   // > 10.03 * 2^-53 * 2^24 = 1.87e-8
constexpr double K3_EPS_CONST = 1e-6;
constexpr double K3_TINY = 1e-280, K3_HUGE = 1e280;

struct PPTask { int ev, s, e, flags; };
struct K3Item { int s, e, ps; };
struct K3Params { int mw, MW, W; double min_gain; };
struct K3Best { double g; int x; };
struct K3Approx { double b1, b2; int i1, bad; };

struct K3Global {
    const double2 *cc;
    const int64_t *ev_off;
    const int64_t *ev_len;
    unsigned *bits;
    PPTask *tasks;
    int *ready;
    int64_t q_cap;
    PPCounters *ctr;
    const double *T;  // T[n] = 2 n ln n for n in [0, W]
    int screen;       // 0: exact evaluation of every candidate (validation mode)
};

struct K3Shared {
    K3Item list[2][K3_LIST];
    int win_item[K3_LIST];                 // scan window slot -> index into the level list (| K3_FULL_FLAG)
    int win_chunk0[K3_LIST + 1];           // first chunk of the window (exclusive prefix of chunk counts)
    unsigned long long win_gmin[K3_LIST];  // ordered key of the window's minimum screened value
    unsigned long long best_key[K3_LIST];  // ordered key of the best exact gain (0 = none beats min_gain)
    int best_idx[K3_LIST];
    double chunk_b1[K3_CHUNKS];            // best screened value of the chunk
    int chunk_i1[K3_CHUNKS];               // its candidate index (| K3_AMB_FLAG: chunk needs the exact scan)
    unsigned short chunk_slot[K3_CHUNKS];
    int req_k[K3_REQ];
    int req_i[K3_REQ];
    double req_g[K3_REQ];
    double2 logtab[256];
    double red_g[K3_WARPS];
    double red_b[K3_WARPS];
    int red_x[K3_WARPS];
    int nA, nB, nwin, nchunk, nreq;
    PPTask task;
    int have_task;
    unsigned long long cand, scans, exact;
};

constexpr size_t K3_SMEM_BYTES = sizeof(K3Shared);

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
        return __ldg(g + p);
    }
};

// ---------------------------------------------------------------------------
// exact path
// ---------------------------------------------------------------------------
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

// gain(i) of window [ps,pe) exactly as cparsers.pyx:172-174
__device__ __forceinline__ double k3_exact_gain(const double2 lo, const double2 mid, const double2 hi,
                                                int ps, int pe, int i, double tot)
{
    const double low = __dmul_rn((double)(i - ps), log(k3_var(mid, lo, i - ps)));
    const double high = __dmul_rn((double)(pe - i), log(k3_var(hi, mid, pe - i)));
    return __dsub_rn(tot, __dadd_rn(low, high));
}

__device__ __forceinline__ double k3_exact_tot(const double2 lo, const double2 hi, int ps, int pe)
{
    return __dmul_rn((double)(pe - ps), log(k3_var(hi, lo, pe - ps)));
}

// Candidates ps+mw+first, +stride, ... <= pe-mw of window [ps,pe), all exact
// (_best_split_stepwise loop, cparsers.pyx:171-177).
template <class CC>
__device__ __forceinline__ K3Best k3_scan_range(const CC &cc, int ps, int pe, int mw,
                                                double min_gain, int first, int stride)
{
    const double2 lo = cc.at(ps - 1), hi = cc.at(pe - 1);
    const double tot = k3_exact_tot(lo, hi, ps, pe);
    K3Best b;
    b.g = min_gain;
    b.x = -1;
    const int last = pe - mw;
    for (int i = ps + mw + first; i <= last; i += stride) {
        const double g = k3_exact_gain(lo, cc.at(i - 1), hi, ps, pe, i, tot);
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

// CTA-wide lexicographic reduction of per-thread exact results (all threads get it)
__device__ __forceinline__ K3Best k3_cta_reduce(K3Best b, K3Shared &S)
{
    const int tid = threadIdx.x;
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

// ---------------------------------------------------------------------------
// screening path
// ---------------------------------------------------------------------------
// ln(x) for positive, normal, finite x: exponent + 8-bit table + degree-3 series.
// |r| <= 2^-9, truncation r^4/4 <= 3.6e-12, rounding a few ulp of the result.
__device__ __forceinline__ double k3_fastlog(double x, const double2 *tab)
{
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int k = (hi >> 12) & 255;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    // (double)(exponent) without an I2F: 2^52 + 2^31 + (e ^ 2^31) as bits, minus the magic
    const double ed = __hiloint2double(0x43300000, ((hi >> 20) - 1023) ^ 0x80000000) - 4503601774854144.0;
    const double2 t = tab[k];
    const double r = fma(m, t.x, -1.0);
    double p = fma(r, 1.0 / 3.0, -0.5);
    p = fma(r, p, 1.0);
    p = p * r;
    return fma(ed, 0.6931471805599453094, t.y + p);
}

// positive, finite and comfortably normal: 2^-900 <= d < 2^900
__device__ __forceinline__ bool k3_sane(double d)
{
    return (unsigned)(__double2hiint(d) - 0x07b00000) < 0x70800000u;
}

// H~(i) and its validity (see file header).  S and Q are the same fp64
// differences the exact path forms, so both paths start from identical values.
// n1 = i - ps and n2 = pe - i arrive as doubles (exact small integers).
template <class CC>
__device__ __forceinline__ bool k3_screen_eval(const CC &cc, const double2 lo, const double2 hi, int ps,
                                               int pe, int i, double n1, double n2,
                                               const double *__restrict__ T, const double2 *tab, double &H)
{
    const double2 mid = cc.at(i - 1);
    const double t12 = __ldg(T + (i - ps)) + __ldg(T + (pe - i));
    const double S1 = __dsub_rn(mid.x, lo.x), Q1 = __dsub_rn(mid.y, lo.y);
    const double S2 = __dsub_rn(hi.x, mid.x), Q2 = __dsub_rn(hi.y, mid.y);
    const double P1 = __dmul_rn(Q1, n1), SQ1 = __dmul_rn(S1, S1), D1 = __dsub_rn(P1, SQ1);
    const double P2 = __dmul_rn(Q2, n2), SQ2 = __dmul_rn(S2, S2), D2 = __dsub_rn(P2, SQ2);
    const double R1 = K3_RATIO_MAX * D1, R2 = K3_RATIO_MAX * D2;
    // max(|P|, SQ) <= R * D, written as two comparisons (false on NaN)
    bool ok = k3_sane(D1) & k3_sane(D2);
    ok = ok & (fabs(P1) <= R1) & (SQ1 <= R1) & (fabs(P2) <= R2) & (SQ2 <= R2);
    const double L1 = k3_fastlog(D1, tab), L2 = k3_fastlog(D2, tab);
    H = fma(n1, L1, n2 * L2) - t12;
    return ok;
}

// Screen candidates ps+mw+first, +stride, ...  (requires mw >= 1 so that n1, n2 >= 1)
template <class CC>
__device__ __forceinline__ K3Approx k3_screen_range(const CC &cc, const double2 lo, const double2 hi,
                                                    int ps, int pe, int mw, int first, int stride,
                                                    const double *__restrict__ T, const double2 *tab)
{
    K3Approx a;
    a.b1 = a.b2 = __longlong_as_double(0x7ff0000000000000LL);
    a.i1 = -1;
    a.bad = 0;
    const int last = pe - mw;
    const double dstride = (double)stride;
    double n1 = (double)(mw + first), n2 = (double)(pe - ps - mw - first);
    for (int i = ps + mw + first; i <= last; i += stride) {
        double H;
        const bool ok = k3_screen_eval(cc, lo, hi, ps, pe, i, n1, n2, T, tab, H);
        n1 += dstride;
        n2 -= dstride;
        if (!ok) a.bad = 1;
        else if (H < a.b2) {  // uncommon after the first few candidates
            if (H < a.b1) { a.b2 = a.b1; a.b1 = H; a.i1 = i; }
            else a.b2 = H;
        }
    }
    return a;
}

__device__ __forceinline__ double k3_eps(int n) { return (double)n * K3_EPS_PER_SAMPLE + K3_EPS_CONST; }

// Monotone double -> u64 key (non-NaN); -0.0 is canonicalised to +0.0 first so
// that equal doubles have equal keys.
__device__ __forceinline__ unsigned long long k3_okey(double g)
{
    const long long b = __double_as_longlong(__dadd_rn(g, 0.0));
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000LL));
}

__device__ __forceinline__ double k3_okey_inv(unsigned long long k)
{
    const long long b = (long long)k;
    return __longlong_as_double(b < 0 ? (b ^ (long long)0x8000000000000000LL) : ~b);
}

// warp-wide minimum of a 64-bit key with two 32-bit REDUX operations
__device__ __forceinline__ unsigned long long k3_warp_min_u64(unsigned long long k)
{
    const unsigned hi = (unsigned)(k >> 32);
    const unsigned mh = __reduce_min_sync(PP_FULL, hi);
    const unsigned lo = hi == mh ? (unsigned)k : 0xffffffffu;
    const unsigned ml = __reduce_min_sync(PP_FULL, lo);
    return ((unsigned long long)mh << 32) | ml;
}

// After the window's min H~ is known: which of this thread's candidates need the
// exact evaluation?  f(i) is called for each.
template <class CC, class F>
__device__ __forceinline__ void k3_screen_select(const CC &cc, const double2 lo, const double2 hi, int ps,
                                                 int pe, int mw, int first, int stride,
                                                 const double *__restrict__ T, const double2 *tab,
                                                 const K3Approx &a, double thr, F f)
{
    if (a.bad || a.b2 <= thr) {
        // rare: several of this thread's candidates qualify -> rescan its subset
        const int last = pe - mw;
        for (int i = ps + mw + first; i <= last; i += stride) {
            double H;
            const bool ok = k3_screen_eval(cc, lo, hi, ps, pe, i, (double)(i - ps), (double)(pe - i), T, tab, H);
            if (!ok || H <= thr) f(i);
        }
    } else if (a.i1 >= 0 && a.b1 <= thr) {
        f(a.i1);
    }
}

// Whole CTA scans one window and returns the exact decision (spine mode).
template <class CC>
__device__ __forceinline__ K3Best k3_cta_scan(const CC &cc, int ps, int pe, const K3Params &P,
                                              const K3Global &G, K3Shared &S)
{
    const int tid = threadIdx.x;
    K3Best b;
    b.g = P.min_gain;
    b.x = -1;
    const double2 lo = cc.at(ps - 1), hi = cc.at(pe - 1);
    const double tot = k3_exact_tot(lo, hi, ps, pe);
    if (!G.screen || P.mw < 1 || !(fabs(tot) <= K3_HUGE)) {
        b = k3_scan_range(cc, ps, pe, P.mw, P.min_gain, tid, K3_THREADS);
        return k3_cta_reduce(b, S);
    }
    const K3Approx a = k3_screen_range(cc, lo, hi, ps, pe, P.mw, tid, K3_THREADS, G.T, S.logtab);
    double m = a.b1;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmin(m, __shfl_xor_sync(PP_FULL, m, d));
    __syncthreads();
    if ((tid & 31) == 0) S.red_b[tid >> 5] = m;
    __syncthreads();
    m = S.red_b[0];
#pragma unroll
    for (int w = 1; w < K3_WARPS; ++w) m = fmin(m, S.red_b[w]);
    const double thr = m + 2.0 * k3_eps(pe - ps);
    unsigned nexact = 0;
    k3_screen_select(cc, lo, hi, ps, pe, P.mw, tid, K3_THREADS, G.T, S.logtab, a, thr, [&](int i) {
        const double g = k3_exact_gain(lo, cc.at(i - 1), hi, ps, pe, i, tot);
        ++nexact;
        if (g > b.g) { b.g = g; b.x = i; }
    });
    if (nexact) atomicAdd(&S.exact, (unsigned long long)nexact);
    return k3_cta_reduce(b, S);
}

// ---------------------------------------------------------------------------
// bookkeeping shared by both modes
// ---------------------------------------------------------------------------
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

__device__ __forceinline__ int k3_window_end(const K3Params &P, const K3Item &it)
{
    const long long pe_l = (long long)it.ps + P.W;
    return (int)(pe_l < it.e ? pe_l : it.e);
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

__device__ __forceinline__ void k3_request(K3Shared &S, int k, int i)
{
    const int r = atomicAdd(&S.nreq, 1);
    if (r < K3_REQ) { S.req_k[r] = k; S.req_i[r] = i; }
    else atomicOr(&S.win_item[k], K3_FULL_FLAG);
}

__global__ void __launch_bounds__(K3_THREADS, K3_CTAS_PER_SM) k3_split(K3Global G, K3Params P)
{
    extern __shared__ __align__(16) unsigned char k3_smem[];
    K3Shared &S = *reinterpret_cast<K3Shared *>(k3_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mw = P.mw, MW = P.MW, W = P.W;

    for (int k = tid; k < 256; k += K3_THREADS) {
        // fast-log table: centre of mantissa bucket k, its reciprocal and its logarithm
        const double c = 1.0 + ((double)k + 0.5) / 256.0;
        S.logtab[k] = make_double2(1.0 / c, log(c));
    }

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
            S.exact = 0;
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
            K3GlobalCC gacc;
            gacc.g = ccg;
            const K3Best b = k3_cta_scan(gacc, ps, pe, P, G, S);
            if (tid == 0) {
                atomicAdd(&S.cand, (unsigned long long)(pe - ps - 2 * mw + 1));
                atomicAdd(&S.scans, 1ull);
            }
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
            if (tid == 0) {
                K3Item it;
                it.s = s; it.e = e; it.ps = ps;
                S.list[0][0] = it;
                S.nA = 1;
            }
            K3GlobalCC acc;
            acc.g = ccg;
            int cur = 0;
            __syncthreads();
            for (;;) {
                const int nA = S.nA;
                if (nA == 0) break;
                K3Item *A = S.list[cur], *Bn = S.list[cur ^ 1];
                __syncthreads();  // everyone has read nA
                if (tid == 0) { S.nB = 0; S.nwin = 0; S.nreq = 0; S.nchunk = 0; }
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
                        const int pe = k3_window_end(P, it);
                        if (pe - it.ps <= 2 * mw) {
                            k3_push_local(G, S, Bn, ev, it.s, it.e, k3_next_ps(P, it.ps, it.e));
                        } else {
                            const int slot = atomicAdd(&S.nwin, 1);
                            S.win_item[slot] = t;
                            S.win_chunk0[slot] = (pe - it.ps - 2 * mw + 1 + 31) >> 5;  // chunk count for now
                            S.win_gmin[slot] = ~0ull;
                            S.best_key[slot] = 0ull;
                            S.best_idx[slot] = 0x7fffffff;
                        }
                    }
                }
                __syncthreads();
                const int nwin = S.nwin;
                const bool screen = G.screen && mw >= 1;
                if (screen) {
                    // exclusive prefix of the chunk counts (warp 0), in place
                    if (warp == 0) {
                        int carry = 0;
                        for (int b0 = 0; b0 < nwin; b0 += 32) {
                            const int v = b0 + lane < nwin ? S.win_chunk0[b0 + lane] : 0;
                            int inc = v;
#pragma unroll
                            for (int d = 1; d < 32; d <<= 1) {
                                const int t = __shfl_up_sync(PP_FULL, inc, d);
                                if (lane >= d) inc += t;
                            }
                            if (b0 + lane < nwin) S.win_chunk0[b0 + lane] = carry + inc - v;
                            carry += __shfl_sync(PP_FULL, inc, 31);
                        }
                        if (lane == 0) { S.win_chunk0[nwin] = carry; S.nchunk = carry; }
                    }
                    __syncthreads();
                    const int nchunk = S.nchunk;
                    // step 2: SCREEN -- all windows cut into 32-candidate chunks, spread evenly over the warps
                    {
                        const int per = (nchunk + K3_WARPS - 1) / K3_WARPS;
                        const int c_lo = warp * per;
                        const int c_hi = c_lo + per < nchunk ? c_lo + per : nchunk;
                        if (c_lo < c_hi) {
                            int slot = 0;
                            {   // largest slot with win_chunk0[slot] <= c_lo
                                int lo_s = 0, hi_s = nwin;
                                while (hi_s - lo_s > 1) {
                                    const int mid = (lo_s + hi_s) >> 1;
                                    if (S.win_chunk0[mid] <= c_lo) lo_s = mid; else hi_s = mid;
                                }
                                slot = lo_s;
                            }
                            int next_c0 = S.win_chunk0[slot + 1], w_c0 = 0;
                            int w_ps = 0, w_pe = 0, w_last = 0;
                            bool fresh = true;
                            double2 w_lo = make_double2(0.0, 0.0), w_hi = w_lo;
                            double w_eps2 = 0.0;
                            for (int c = c_lo; c < c_hi; ++c) {
                                if (c >= next_c0) {
                                    do { ++slot; next_c0 = S.win_chunk0[slot + 1]; } while (c >= next_c0);
                                    fresh = true;
                                }
                                if (fresh) {
                                    fresh = false;
                                    const K3Item it = A[S.win_item[slot]];
                                    w_c0 = S.win_chunk0[slot];
                                    w_ps = it.ps;
                                    w_pe = k3_window_end(P, it);
                                    w_last = w_pe - mw;
                                    w_lo = acc.at(w_ps - 1);
                                    w_hi = acc.at(w_pe - 1);
                                    w_eps2 = 2.0 * k3_eps(w_pe - w_ps);
                                }
                                const int i0 = w_ps + mw + (c - w_c0) * 32;
                                const int i = i0 + lane;
                                const bool valid = i <= w_last;
                                const int ic = valid ? i : w_last;  // clamp: evaluate a real candidate, ignore it
                                double H;
                                const bool ok = k3_screen_eval(acc, w_lo, w_hi, w_ps, w_pe, ic, (double)(ic - w_ps),
                                                               (double)(w_pe - ic), G.T, S.logtab, H);
                                const bool use = valid && ok;
                                const unsigned bad = __ballot_sync(PP_FULL, valid && !ok);
                                const unsigned long long k1 = k3_warp_min_u64(use ? k3_okey(H) : ~0ull);
                                const int l1 = __ffs(__ballot_sync(PP_FULL, use && k3_okey(H) == k1)) - 1;
                                const double b1 = __shfl_sync(PP_FULL, H, l1 < 0 ? 0 : l1);
                                // another candidate of this chunk within 2 eps of its best -> exact scan decides
                                const unsigned close = __ballot_sync(PP_FULL, use && H <= b1 + w_eps2);
                                if (lane == 0) {
                                    const bool amb = l1 < 0 || __popc(close) > 1;
                                    // a candidate that failed the validity test forces the exact scan of its window
                                    S.chunk_b1[c] = (bad || l1 < 0) ? __longlong_as_double(0xfff0000000000000LL) : b1;
                                    S.chunk_i1[c] = (i0 + (l1 < 0 ? 0 : l1)) | ((amb || bad) ? K3_AMB_FLAG : 0);
                                    S.chunk_slot[c] = (unsigned short)slot;
                                }
                            }
                        }
                    }
                    __syncthreads();
                    // window minimum of the screened values: one warp per window over its chunk records
                    for (int k = warp; k < nwin; k += K3_WARPS) {
                        const int c0 = S.win_chunk0[k], c1 = S.win_chunk0[k + 1];
                        double m = __longlong_as_double(0x7ff0000000000000LL);
                        bool any_ok = false;
                        for (int c = c0 + lane; c < c1; c += 32) {
                            const double v = S.chunk_b1[c];
                            if (v > __longlong_as_double(0xfff0000000000000LL)) { m = fmin(m, v); any_ok = true; }
                        }
                        const unsigned long long km = k3_warp_min_u64(any_ok ? k3_okey(m) : ~0ull);
                        if (lane == 0) S.win_gmin[k] = km;
                    }
                    __syncthreads();
                    // step 3a: chunks whose best is within 2 eps of the window minimum ask for the exact value
                    for (int c = tid; c < nchunk; c += K3_THREADS) {
                        const int slot = S.chunk_slot[c];
                        const int i1 = S.chunk_i1[c];
                        const unsigned long long gk = S.win_gmin[slot];
                        if (gk == ~0ull) { atomicOr(&S.win_item[slot], K3_FULL_FLAG); continue; }
                        const K3Item it = A[S.win_item[slot] & ~K3_FULL_FLAG];
                        const int pe = k3_window_end(P, it);
                        const double thr = k3_okey_inv(gk) + 2.0 * k3_eps(pe - it.ps);
                        if (i1 & K3_AMB_FLAG) {
                            // only matters if this chunk can reach the window minimum at all
                            if (!(S.chunk_b1[c] > thr)) atomicOr(&S.win_item[slot], K3_FULL_FLAG);
                        } else if (S.chunk_b1[c] <= thr) {
                            k3_request(S, slot, i1);
                        }
                    }
                    __syncthreads();
                    // step 3b: EXACT evaluation of the requests; 3 lanes per request (tot / low / high)
                    const int nreq = S.nreq < K3_REQ ? S.nreq : K3_REQ;
                    for (int r0 = 0; r0 < nreq; r0 += 10 * K3_WARPS) {
                        const int part = lane % 3;
                        const int r = r0 + warp * 10 + lane / 3;
                        const bool act = lane < 30 && r < nreq;
                        double v = 0.0;
                        int k = 0, i = 0, w_ps = 0, w_pe = 0;
                        if (act) {
                            k = S.req_k[r];
                            i = S.req_i[r];
                            const K3Item it = A[S.win_item[k] & ~K3_FULL_FLAG];
                            w_ps = it.ps;
                            w_pe = k3_window_end(P, it);
                            const double2 lo = acc.at(w_ps - 1), hi = acc.at(w_pe - 1), mid = acc.at(i - 1);
                            if (part == 0) v = k3_exact_tot(lo, hi, w_ps, w_pe);
                            else if (part == 1) v = __dmul_rn((double)(i - w_ps), log(k3_var(mid, lo, i - w_ps)));
                            else v = __dmul_rn((double)(w_pe - i), log(k3_var(hi, mid, w_pe - i)));
                        }
                        const double low = __shfl_down_sync(PP_FULL, v, 1), high = __shfl_down_sync(PP_FULL, v, 2);
                        if (act && part == 0) {
                            const double g = __dsub_rn(v, __dadd_rn(low, high));   // cparsers.pyx:174
                            S.req_g[r] = g;
                            if (g > P.min_gain) atomicMax(&S.best_key[k], k3_okey(g));
                        }
                    }
                    if (tid == 0) atomicAdd(&S.exact, (unsigned long long)nreq);
                    __syncthreads();
                    for (int r = tid; r < nreq; r += K3_THREADS) {
                        const double g = S.req_g[r];
                        const int k = S.req_k[r];
                        if (g > P.min_gain && k3_okey(g) == S.best_key[k]) atomicMin(&S.best_idx[k], S.req_i[r]);
                    }
                    __syncthreads();
                    // step 4: resolve every window whose screening was conclusive
                    for (int k = tid; k < nwin; k += K3_THREADS) {
                        const int entry = S.win_item[k];
                        if (entry & K3_FULL_FLAG) continue;
                        const K3Item it = A[entry];
                        const int pe = k3_window_end(P, it);
                        atomicAdd(&S.cand, (unsigned long long)(pe - it.ps - 2 * mw + 1));
                        atomicAdd(&S.scans, 1ull);
                        k3_resolve_local(G, S, Bn, P, ev, off, it, S.best_key[k] ? S.best_idx[k] : -1);
                    }
                }
                // step 5: windows left for the exact scan (validation mode: all of them), whole CTA each
                for (int k = 0; k < nwin; ++k) {
                    const int entry = S.win_item[k];
                    if (screen && !(entry & K3_FULL_FLAG)) continue;
                    const K3Item it = A[entry & ~K3_FULL_FLAG];
                    const int pe = k3_window_end(P, it);
                    K3Best b = k3_scan_range(acc, it.ps, pe, mw, P.min_gain, tid, K3_THREADS);
                    b = k3_cta_reduce(b, S);
                    if (tid == 0) {
                        atomicAdd(&S.cand, (unsigned long long)(pe - it.ps - 2 * mw + 1));
                        atomicAdd(&S.scans, 1ull);
                        if (screen) atomicAdd(&S.exact, (unsigned long long)(pe - it.ps - 2 * mw + 1));
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
            atomicAdd(&G.ctr->n_exact, S.exact);
            atomicAdd(&G.ctr->n_tasks, 1ull);
            __threadfence();
            atomicAdd((unsigned long long *)&G.ctr->q_pending, (unsigned long long)(-1LL));
        }
    }
}

// Debug / validation: for window [ps,pe) of event `ev`, write per candidate the screened
// value H~, the reference-arithmetic value fl(low + high) and the validity flag.
__global__ void __launch_bounds__(256)
k3_debug_screen(K3Global G, int ev, int ps, int pe, int mw, double *h_screen, double *h_exact,
                unsigned char *ok_out)
{
    __shared__ double2 tab[256];
    {
        const double c = 1.0 + ((double)threadIdx.x + 0.5) / 256.0;
        tab[threadIdx.x] = make_double2(1.0 / c, log(c));
    }
    __syncthreads();
    K3GlobalCC cc;
    cc.g = G.cc + G.ev_off[ev];
    const double2 lo = cc.at(ps - 1), hi = cc.at(pe - 1);
    const int n = pe - mw - (ps + mw) + 1;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int i = ps + mw + j;
        double H;
        const bool ok = k3_screen_eval(cc, lo, hi, ps, pe, i, (double)(i - ps), (double)(pe - i), G.T, tab, H);
        const double2 mid = cc.at(i - 1);
        const double low = __dmul_rn((double)(i - ps), log(k3_var(mid, lo, i - ps)));
        const double high = __dmul_rn((double)(pe - i), log(k3_var(hi, mid, pe - i)));
        h_screen[j] = H;
        h_exact[j] = __dadd_rn(low, high);
        ok_out[j] = ok ? 1 : 0;
    }
}

// T[n] = 2 n ln n (n = 0 .. len-1), the part of the screening formula that only
// depends on the sub-window length.
__global__ void __launch_bounds__(256) k3_fill_T(double *T, int len)
{
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < len; n += gridDim.x * blockDim.x)
        T[n] = n > 0 ? 2.0 * (double)n * log((double)n) : 0.0;
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
        G.ctr->n_exact = 0;
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
