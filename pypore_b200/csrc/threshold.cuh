// threshold.cuh -- K1: lambda_event_parser.parse as one streaming pass.
//
// Replaces PyPore/parsers.py:148-155 (mask / diff / where / split) and the
// np.min / np.max second pass of _lambda_select (parsers.py:136-140 through
// core.py:215-220).  One read of the float32 trace (4 B/sample): every tile of
// 4096 samples builds a below-threshold bitmask with 128-bit loads, finds the
// crossings with one XOR per 32 samples, gets its global run offset by a
// decoupled look-back over the per-tile crossing counts, scatters the run
// starts, and folds per-run min/max through shared-memory keys into the global
// run table (one atomic pair per run per tile).
#pragma once
#include "common.cuh"

constexpr int K1_THREADS = 256;
#ifndef K1_CFG_ROWS  // development knob
#define K1_CFG_ROWS 4
#endif
constexpr int K1_ROWS = K1_CFG_ROWS;  // <= 8 (one thread per mask word)
constexpr int K1_TILE = K1_THREADS * 4 * K1_ROWS;  // 4096 samples
constexpr int K1_WORDS = K1_TILE / 32;             // 128 mask words
constexpr int K1_LOCAL_RUNS = 64;

#define PP_TS_AGG (1ull << 62)
#define PP_TS_INC (2ull << 62)
#define PP_TS_MASK (3ull << 62)

__device__ __forceinline__ void k1_local_update(unsigned rid, float m, float M, bool nan,
                                                unsigned *lmin, unsigned *lmax)
{
    unsigned kmin = nan ? 0u : pp_fkey(m);
    unsigned kmax = nan ? 0xffffffffu : pp_fkey(M);
    // rid >= K1_LOCAL_RUNS is handled by the caller
    atomicMin(&lmin[rid], kmin);
    atomicMax(&lmax[rid], kmax);
}

#ifdef K1_CFG_MINBLOCKS  // development knob: 8 = all 2048 threads of an SM resident (32 registers per thread)
__global__ void __launch_bounds__(K1_THREADS, K1_CFG_MINBLOCKS)
#else
__global__ void __launch_bounds__(K1_THREADS)
#endif
k1_threshold_scan(const float *__restrict__ x, int64_t n, float thr_f,
                  unsigned long long *__restrict__ tile_state, PPCounters *ctr,
                  int64_t *__restrict__ run_start, unsigned *__restrict__ run_minkey,
                  unsigned *__restrict__ run_maxkey, int64_t cap_runs)
{
    __shared__ unsigned B[K1_WORDS];
    __shared__ unsigned E[K1_WORDS];
    __shared__ unsigned Eoff[K1_WORDS];
    __shared__ unsigned warp_tot[K1_WORDS / 32];
    __shared__ unsigned lmin[K1_LOCAL_RUNS], lmax[K1_LOCAL_RUNS];
    __shared__ float red_min[K1_THREADS / 32], red_max[K1_THREADS / 32];
    __shared__ int red_nan[K1_THREADS / 32];
    __shared__ unsigned long long s_tile, s_prefix;
    __shared__ unsigned s_total;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(&ctr->ticket, 1ull);
    if (tid < K1_LOCAL_RUNS) { lmin[tid] = PP_MINKEY_EMPTY; lmax[tid] = PP_MAXKEY_EMPTY; }
    __syncthreads();
    const unsigned long long tile = s_tile;
    const int64_t base = (int64_t)tile * K1_TILE;
    const bool full_tile = base + K1_TILE <= n;
    const float fill = full_tile ? 0.f : __ldg(x + (n - 1));

    float v[K1_ROWS][4];
#pragma unroll
    for (int r = 0; r < K1_ROWS; ++r) {
        const int64_t idx = base + r * (K1_THREADS * 4) + tid * 4;
        if (idx + 3 < n) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(x + idx));
            v[r][0] = q.x; v[r][1] = q.y; v[r][2] = q.z; v[r][3] = q.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[r][k] = (idx + k < n) ? __ldg(x + idx + k) : fill;
        }
    }
    // below-threshold bitmask of the tile
#pragma unroll
    for (int r = 0; r < K1_ROWS; ++r) {
        unsigned nib = (v[r][0] < thr_f ? 1u : 0u) | (v[r][1] < thr_f ? 2u : 0u) |
                       (v[r][2] < thr_f ? 4u : 0u) | (v[r][3] < thr_f ? 8u : 0u);
        unsigned w = nib << ((lane & 7) * 4);
        w |= __shfl_xor_sync(PP_FULL, w, 1);
        w |= __shfl_xor_sync(PP_FULL, w, 2);
        w |= __shfl_xor_sync(PP_FULL, w, 4);
        if ((lane & 7) == 0) B[r * 32 + warp * 4 + (lane >> 3)] = w;
    }
    __syncthreads();
    // crossings: bit p set <=> sample base+p is on the other side than base+p-1
    if (tid < K1_WORDS) {
        const unsigned b = B[tid];
        unsigned carry;
        if (tid > 0) carry = B[tid - 1] >> 31;
        else if (base == 0) carry = b & 1u;  // sample 0 never starts a new run
        else carry = (__ldg(x + base - 1) < thr_f) ? 1u : 0u;
        const unsigned e = b ^ ((b << 1) | carry);
        E[tid] = e;
        const unsigned cnt = __popc(e);
        unsigned inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned t = __shfl_up_sync(PP_FULL, inc, d);
            if (lane >= d) inc += t;
        }
        Eoff[tid] = inc - cnt;
        if (lane == 31) warp_tot[warp] = inc;
    }
    __syncthreads();
    if (tid < K1_WORDS) {
        unsigned add = 0;
        for (int w = 0; w < warp; ++w) add += warp_tot[w];
        Eoff[tid] += add;
        if (tid == K1_WORDS - 1) s_total = Eoff[tid] + __popc(E[tid]);
    }
    __syncthreads();
    const unsigned total = s_total;

    // decoupled look-back over per-tile crossing counts
    if (warp == 0) {
        unsigned long long prefix = 0;
        if (tile == 0) {
            if (lane == 0) {
                atomicExch(&tile_state[0], PP_TS_INC | (unsigned long long)total);
                ctr->first_below = B[0] & 1u;
                if (cap_runs > 0) run_start[0] = 0;
            }
        } else {
            if (lane == 0) atomicExch(&tile_state[tile], PP_TS_AGG | (unsigned long long)total);
            long long j = (long long)tile - 1;
            for (;;) {
                const long long idx = j - lane;
                unsigned long long s = idx >= 0 ? pp_ld_volatile_u64(tile_state + idx) : PP_TS_INC;
                while (__any_sync(PP_FULL, (s & PP_TS_MASK) == 0ull))
                    s = idx >= 0 ? pp_ld_volatile_u64(tile_state + idx) : PP_TS_INC;
                const unsigned inc_mask = __ballot_sync(PP_FULL, (s & PP_TS_MASK) == PP_TS_INC);
                unsigned long long val = s & ~PP_TS_MASK;
                if (inc_mask) {
                    const int first = __ffs(inc_mask) - 1;
                    if (lane > first) val = 0;
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(PP_FULL, val, d);
                prefix += val;
                if (inc_mask) break;
                j -= 32;
            }
            if (lane == 0) atomicExch(&tile_state[tile], PP_TS_INC | (prefix + total));
        }
        if (lane == 0) {
            s_prefix = prefix;
            if (base + K1_TILE >= n) {  // last tile
                ctr->n_edges = prefix + total;
                ctr->n_runs = prefix + total + 1;
            }
        }
    }

    // per-thread min/max (independent of the look-back)
    if (total == 0) {
        float m = v[0][0], M = v[0][0];
        bool nan = false;
#pragma unroll
        for (int r = 0; r < K1_ROWS; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                m = fminf(m, v[r][k]); M = fmaxf(M, v[r][k]);
                nan |= (v[r][k] != v[r][k]);
            }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            m = fminf(m, __shfl_xor_sync(PP_FULL, m, d));
            M = fmaxf(M, __shfl_xor_sync(PP_FULL, M, d));
        }
        const bool wnan = __any_sync(PP_FULL, nan);
        if (lane == 0) { red_min[warp] = m; red_max[warp] = M; red_nan[warp] = wnan; }
        __syncthreads();
        if (tid == 0) {
            bool anynan = false;
            for (int w = 0; w < K1_THREADS / 32; ++w) {
                m = fminf(m, red_min[w]); M = fmaxf(M, red_max[w]); anynan |= (red_nan[w] != 0);
            }
            const unsigned long long g = s_prefix;
            if ((int64_t)g < cap_runs) {
                atomicMin(&run_minkey[g], anynan ? 0u : pp_fkey(m));
                atomicMax(&run_maxkey[g], anynan ? 0xffffffffu : pp_fkey(M));
            }
        }
        return;
    }

    __syncthreads();  // s_prefix visible
    const unsigned long long prefix = s_prefix;
    // scatter run starts: crossing k (0-based, global) starts run k+1
    if (tid < K1_WORDS) {
        unsigned e = E[tid];
        unsigned long long k = prefix + Eoff[tid];
        while (e) {
            const int bit = __ffs(e) - 1;
            e &= e - 1;
            const unsigned long long rid = ++k;
            if ((int64_t)rid < cap_runs) run_start[rid] = base + tid * 32 + bit;
        }
    }
#pragma unroll
    for (int r = 0; r < K1_ROWS; ++r) {
        const int j = r * 32 + warp * 4 + (lane >> 3);
        const int b0 = (lane & 7) * 4;
        const unsigned e = E[j];
        const unsigned rid0 = Eoff[j] + __popc(e & ((2u << b0) - 1u));
        const unsigned inner = (e >> (b0 + 1)) & 7u;
        const unsigned rid_first = __shfl_sync(PP_FULL, rid0, 0);
        const bool uniform = __all_sync(PP_FULL, inner == 0u && rid0 == rid_first);
        if (uniform) {
            float m = fminf(fminf(v[r][0], v[r][1]), fminf(v[r][2], v[r][3]));
            float M = fmaxf(fmaxf(v[r][0], v[r][1]), fmaxf(v[r][2], v[r][3]));
            bool nan = (v[r][0] != v[r][0]) | (v[r][1] != v[r][1]) | (v[r][2] != v[r][2]) |
                       (v[r][3] != v[r][3]);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                m = fminf(m, __shfl_xor_sync(PP_FULL, m, d));
                M = fmaxf(M, __shfl_xor_sync(PP_FULL, M, d));
            }
            nan = __any_sync(PP_FULL, nan);
            if (lane == 0) {
                if (rid0 < K1_LOCAL_RUNS) k1_local_update(rid0, m, M, nan, lmin, lmax);
                else if ((int64_t)(prefix + rid0) < cap_runs) {
                    atomicMin(&run_minkey[prefix + rid0], nan ? 0u : pp_fkey(m));
                    atomicMax(&run_maxkey[prefix + rid0], nan ? 0xffffffffu : pp_fkey(M));
                }
            }
        } else {
            unsigned rid = rid0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k > 0) rid += (inner >> (k - 1)) & 1u;
                const float val = v[r][k];
                const bool nan = val != val;
                if (rid < K1_LOCAL_RUNS) k1_local_update(rid, val, val, nan, lmin, lmax);
                else if ((int64_t)(prefix + rid) < cap_runs) {
                    atomicMin(&run_minkey[prefix + rid], nan ? 0u : pp_fkey(val));
                    atomicMax(&run_maxkey[prefix + rid], nan ? 0xffffffffu : pp_fkey(val));
                }
            }
        }
    }
    __syncthreads();
    if (tid < K1_LOCAL_RUNS && (unsigned)tid <= total) {
        const unsigned long long g = prefix + tid;
        if ((int64_t)g < cap_runs && lmin[tid] != PP_MINKEY_EMPTY) {
            atomicMin(&run_minkey[g], lmin[tid]);
            atomicMax(&run_maxkey[g], lmax[tid]);
        }
    }
}

// Decode the run table: length, side, min/max as float64 (what the rules see).
__global__ void __launch_bounds__(256)
k1_finalize_runs(int64_t n, const PPCounters *ctr, const int64_t *__restrict__ run_start,
                 const unsigned *__restrict__ run_minkey, const unsigned *__restrict__ run_maxkey,
                 int64_t cap_runs, int64_t *__restrict__ run_len, double *__restrict__ run_min,
                 double *__restrict__ run_max, unsigned char *__restrict__ run_below)
{
    int64_t n_runs = (int64_t)ctr->n_runs;
    if (n_runs > cap_runs) n_runs = cap_runs;
    const unsigned first_below = ctr->first_below;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_runs;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = run_start[r];
        const int64_t e = (r + 1 < (int64_t)ctr->n_runs && r + 1 < cap_runs) ? run_start[r + 1] : n;
        run_len[r] = e - s;
        run_min[r] = pp_decode_min(run_minkey[r]);
        run_max[r] = pp_decode_max(run_maxkey[r]);
        run_below[r] = (unsigned char)((first_below ^ (unsigned)(r & 1)) & 1u);
    }
}

// _lambda_select for the default-shaped rules + compaction into the event
// table (ev_start, ev_len, ev_off = exclusive prefix of lengths).  One CTA;
// the run table is tiny next to the trace.
constexpr int SEL_THREADS = 1024;

__global__ void __launch_bounds__(SEL_THREADS)
k1_select_events(PPCounters *ctr, const int64_t *__restrict__ run_start,
                 const int64_t *__restrict__ run_len, const double *__restrict__ run_min,
                 const double *__restrict__ run_max, int64_t cap_runs, int rule_mask,
                 int64_t duration_gt, int64_t duration_lt, double min_gt, double max_lt,
                 int skip_first, int skip_last, int64_t *__restrict__ ev_start,
                 int64_t *__restrict__ ev_len, int64_t *__restrict__ ev_off, int64_t cap_events,
                 int incremental /* 0: whole run table; 1: append the runs completed since the last call,
                                    the still open last run excluded; 2: same, last call (open run included) */,
                 const long long *__restrict__ plan = nullptr /* multi-GPU: k_shard_plan's record overrides
                                                                 skip_first / skip_last */)
{
    if (plan) { skip_first = (int)plan[0]; skip_last = (int)plan[1]; }
    __shared__ unsigned wcnt[SEL_THREADS / 32];
    __shared__ long long wlen[SEL_THREADS / 32];
    __shared__ unsigned long long s_cnt_carry;
    __shared__ long long s_len_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t n_runs = (int64_t)ctr->n_runs;
    if (n_runs > cap_runs) n_runs = cap_runs;
    int64_t r_begin = 0;
    const unsigned long long ev_before = incremental ? ctr->n_events : 0ull;
    if (incremental) {
        r_begin = (int64_t)ctr->sel_next_run;
        if (incremental == 1 && n_runs > 0) n_runs -= 1;
        if (n_runs < r_begin) n_runs = r_begin;
    }
    if (tid == 0) {
        s_cnt_carry = ev_before;
        s_len_carry = incremental ? (long long)ctr->n_event_samples : 0;
    }
    __syncthreads();
    for (int64_t c0 = r_begin; c0 < n_runs; c0 += SEL_THREADS) {
        const int64_t r = c0 + tid;
        bool keep = false;
        long long len = 0;
        if (r < n_runs) {
            len = run_len[r];
            keep = true;
            if (rule_mask & PP_RULE_DURATION_GT) keep = keep && (len > duration_gt);
            if (rule_mask & PP_RULE_DURATION_LT) keep = keep && (len < duration_lt);
            if (rule_mask & PP_RULE_MIN_GT) keep = keep && (run_min[r] > min_gt);
            if (rule_mask & PP_RULE_MAX_LT) keep = keep && (run_max[r] < max_lt);
            if (skip_first && r == 0) keep = false;
            if (skip_last && r == n_runs - 1) keep = false;
        }
        unsigned c = keep ? 1u : 0u;
        long long l = keep ? len : 0;
        unsigned ci = c;
        long long li = l;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned tc = __shfl_up_sync(PP_FULL, ci, d);
            long long tl = __shfl_up_sync(PP_FULL, li, d);
            if (lane >= d) { ci += tc; li += tl; }
        }
        if (lane == 31) { wcnt[warp] = ci; wlen[warp] = li; }
        __syncthreads();
        unsigned cb = 0;
        long long lb = 0;
        for (int w = 0; w < warp; ++w) { cb += wcnt[w]; lb += wlen[w]; }
        const unsigned long long idx = s_cnt_carry + cb + (ci - c);
        const long long off = s_len_carry + lb + (li - l);
        if (keep && (int64_t)idx < cap_events) {
            ev_start[idx] = run_start[r];
            ev_len[idx] = len;
            ev_off[idx] = off;
        }
        __syncthreads();
        if (tid == SEL_THREADS - 1) {
            s_cnt_carry = idx + c;
            s_len_carry = off + l;
        }
        __syncthreads();
    }
    if (tid == 0) {
        unsigned long long ne = s_cnt_carry;
        if ((int64_t)ne > cap_events) { ne = cap_events; atomicOr(&ctr->overflow, PP_OVF_RUNS); }
        ctr->n_events = ne;
        ctr->n_event_samples = (unsigned long long)s_len_carry;
        ctr->sel_next_run = (unsigned long long)n_runs;
        ctr->ev_begin = ev_before < ne ? ev_before : ne;
        ev_off[ne] = s_len_carry;
    }
}

// Host-chosen events: exclusive prefix of the lengths (same single-CTA scan).
__global__ void __launch_bounds__(SEL_THREADS)
k1_event_offsets(PPCounters *ctr, const int64_t *__restrict__ ev_len, int64_t n_events,
                 int64_t *__restrict__ ev_off)
{
    __shared__ long long wlen[SEL_THREADS / 32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < n_events; c0 += SEL_THREADS) {
        const int64_t r = c0 + tid;
        long long l = r < n_events ? ev_len[r] : 0;
        long long li = l;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            long long tl = __shfl_up_sync(PP_FULL, li, d);
            if (lane >= d) li += tl;
        }
        if (lane == 31) wlen[warp] = li;
        __syncthreads();
        long long lb = 0;
        for (int w = 0; w < warp; ++w) lb += wlen[w];
        const long long off = s_carry + lb + (li - l);
        if (r < n_events) ev_off[r] = off;
        __syncthreads();
        if (tid == SEL_THREADS - 1) s_carry = off + l;
        __syncthreads();
    }
    if (tid == 0) {
        ev_off[n_events] = s_carry;
        ctr->n_events = (unsigned long long)n_events;
        ctr->n_event_samples = (unsigned long long)s_carry;
        ctr->ev_begin = 0;
    }
}

// Multi-GPU: the record a rank publishes about its chunk (pypore_b200/dist.py boundary_info):
// [n_local, n_runs, first_below, first_len, first_min, first_max,
//  last_below, last_start, last_len, last_min, last_max, run-table-overflow flag]
__global__ void k1_boundary_record(int64_t n_local, const PPCounters *ctr, const int64_t *__restrict__ run_start,
                                   const int64_t *__restrict__ run_len, const double *__restrict__ run_min,
                                   const double *__restrict__ run_max, const unsigned char *__restrict__ run_below,
                                   int64_t cap_runs, double *__restrict__ rec)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int64_t n_runs = (int64_t)ctr->n_runs;
    const bool ovf = n_runs > cap_runs || n_runs <= 0;
    const int64_t last = ovf ? 0 : n_runs - 1;
    rec[0] = (double)n_local;
    rec[1] = (double)n_runs;
    rec[2] = (double)run_below[0];
    rec[3] = (double)run_len[0];
    rec[4] = run_min[0];
    rec[5] = run_max[0];
    rec[6] = (double)run_below[last];
    rec[7] = (double)run_start[last];
    rec[8] = (double)run_len[last];
    rec[9] = run_min[last];
    rec[10] = run_max[last];
    rec[11] = ovf ? 1.0 : 0.0;
}

// Multi-GPU: this rank's row of dist.py plan_boundaries, derived on the DEVICE from the all-gathered
// boundary records (infos[q] = rank q's k1_boundary_record), so that a step needs no host round trip
// between the threshold scan and the split search.  `halo_avail` continuation samples of the right
// neighbour were placed after the chunk speculatively, before anything was known about the runs.
// plan = [skip_first, skip_last, has_event, ev_start, ev_len, redo, halo samples needed, 0]:
//   redo = PP_OVF_HALO when the straddling event this rank owns needs more than halo_avail samples
//          (or continues past the right neighbour), PP_OVF_RUNS when any rank's run table overflowed;
//          the host then repeats the step with the host-planned exchange (dist.py).
constexpr int PP_PLAN_WORDS = 8;

__device__ __forceinline__ double pp_nan_min(double a, double b) { return (a != a || b != b) ? a + b : (a < b ? a : b); }
__device__ __forceinline__ double pp_nan_max(double a, double b) { return (a != a || b != b) ? a + b : (a > b ? a : b); }

__global__ void k_shard_plan(const double *__restrict__ infos, int rank, int world, int rule_mask,
                             int64_t duration_gt, int64_t duration_lt, double min_gt, double max_lt,
                             int64_t halo_avail, long long *__restrict__ plan)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int L = 12;
    auto joins = [&](int q) { return infos[q * L + 6] == infos[(q + 1) * L + 2]; };  // last_below == next first_below
    long long redo = 0;
    for (int q = 0; q < world; ++q)
        if (infos[q * L + 11] != 0.0) redo |= PP_OVF_RUNS;
    const bool skip_first = rank > 0 && joins(rank - 1);
    bool skip_last = false;
    long long has_event = 0, ev_start = 0, ev_len = 0, need = 0;
    if (rank < world - 1 && joins(rank)) {
        skip_last = true;
        const double *me = infos + rank * L;
        const bool owns = !(me[1] == 1.0 && skip_first);  // else the run started further left
        if (owns) {
            long long length = (long long)me[8];
            double mn = me[9], mx = me[10];
            int q = rank + 1;
            for (;;) {
                const double *o = infos + q * L;
                length += (long long)o[3];
                mn = pp_nan_min(mn, o[4]);
                mx = pp_nan_max(mx, o[5]);
                if (o[1] == 1.0 && q < world - 1 && joins(q)) { ++q; continue; }
                break;
            }
            bool ok = true;
            if (rule_mask & PP_RULE_DURATION_GT) ok = ok && (length > duration_gt);
            if (rule_mask & PP_RULE_DURATION_LT) ok = ok && (length < duration_lt);
            if (rule_mask & PP_RULE_MIN_GT) ok = ok && (mn > min_gt);
            if (rule_mask & PP_RULE_MAX_LT) ok = ok && (mx < max_lt);
            if (ok) {
                need = length - (long long)me[8];
                if (q == rank + 1 && need <= halo_avail) {
                    has_event = 1;
                    ev_start = (long long)me[7];
                    ev_len = length;
                } else {
                    redo |= PP_OVF_HALO;
                }
            }
        }
    }
    plan[0] = skip_first;
    plan[1] = skip_last;
    plan[2] = has_event;
    plan[3] = ev_start;
    plan[4] = ev_len;
    plan[5] = redo;
    plan[6] = need;
    plan[7] = 0;
}

// The straddling event of k_shard_plan appended after the selected events.
__global__ void k_append_planned_event(PPCounters *ctr, int64_t *ev_start, int64_t *ev_len, int64_t *ev_off,
                                       int64_t cap_events, const long long *__restrict__ plan)
{
    if (blockIdx.x != 0 || threadIdx.x != 0 || !plan[2]) return;
    const unsigned long long e = ctr->n_events;
    if ((int64_t)e >= cap_events) { atomicOr(&ctr->overflow, PP_OVF_RUNS); return; }
    ev_start[e] = plan[3];
    ev_len[e] = plan[4];
    ev_off[e + 1] = ev_off[e] + plan[4];  // ev_off[e] already holds the running total
    ctr->n_events = e + 1;
    ctr->n_event_samples += (unsigned long long)plan[4];
}

// Multi-GPU: [n_runs, n_events, n_event_samples, n_segments, overflow flags, candidates, scans, exact]
__global__ void k_result_record(const PPCounters *ctr, long long *__restrict__ rec,
                                const long long *__restrict__ plan = nullptr)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    rec[0] = (long long)ctr->n_runs;
    rec[1] = (long long)ctr->n_events;
    rec[2] = (long long)ctr->n_event_samples;
    rec[3] = (long long)ctr->n_segments;
    rec[4] = (long long)ctr->overflow | (plan ? plan[5] : 0);
    rec[5] = (long long)ctr->n_cand;
    rec[6] = (long long)ctr->n_scan;
    rec[7] = (long long)ctr->n_exact;
}

// Multi-GPU: event and segment tables packed into 8-byte words for ONE all-gather:
// n_events rows {global start, length}, then n_segments rows of PP_SEG_WORDS words
//   { global event id (low 32 bits) | event-relative start (high 32 bits),  mean,  std,  min | max as float32 }
// -- 32 B instead of the table's 56: `end` is the next row's start (or the event's length) and is rebuilt
// after the gather, the sharded path works on float32 traces so min / max are float32 values, and an event
// is shorter than 2^31 samples.  Counts and the global event-id base come from the all-gathered result
// records in DEVICE memory (rec_all[r] = rank r's k_result_record), so the host does not have to know them yet.
constexpr int PP_SEG_WORDS = 4;

__global__ void __launch_bounds__(256)
k_pack_tables(const long long *__restrict__ rec_all, int rank, int64_t cap_words, int64_t sample_offset,
              const int64_t *__restrict__ ev_start, const int64_t *__restrict__ ev_len,
              const int *__restrict__ seg_event, const int64_t *__restrict__ seg_start,
              const int64_t *__restrict__ seg_end, const double *__restrict__ mean,
              const double *__restrict__ sd, const double *__restrict__ mn, const double *__restrict__ mx,
              long long *__restrict__ out)
{
    const int64_t n_events = rec_all[8 * rank + 1], n_segments = rec_all[8 * rank + 3];
    if (2 * n_events + PP_SEG_WORDS * n_segments > cap_words) return;  // the host notices from the same records and retries
    int64_t event_base = 0;
    for (int r = 0; r < rank; ++r) event_base += rec_all[8 * r + 1];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t e = t0; e < n_events; e += stride) {
        out[2 * e] = ev_start[e] + sample_offset;
        out[2 * e + 1] = ev_len[e];
    }
    longlong2 *seg = reinterpret_cast<longlong2 *>(out + 2 * n_events);  // 16-byte aligned: 2 words per event
    for (int64_t k = t0; k < n_segments; k += stride) {
        const unsigned long long id = (unsigned long long)((long long)seg_event[k] + event_base) & 0xffffffffull;
        const unsigned long long st = (unsigned long long)seg_start[k] << 32;
        const unsigned long long lo = (unsigned long long)__float_as_uint((float)mn[k]);
        const unsigned long long hi = (unsigned long long)__float_as_uint((float)mx[k]) << 32;
        seg[2 * k] = make_longlong2((long long)(id | st), __double_as_longlong(mean[k]));
        seg[2 * k + 1] = make_longlong2(__double_as_longlong(sd[k]), (long long)(lo | hi));
    }
}
