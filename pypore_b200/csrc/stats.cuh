// stats.cuh -- K4: Segment.mean / std / min / max (PyPore/core.py:209-223) as a
// segmented reduction over flat event space, fp64.
//
// A group of G lanes (8 for short rows, 32 for long ones) owns one row (segment or event).
// std is computed from deviations, never from prefix-sum differences:
//   * rows of at most K4_ONE_PASS samples: ONE pass with the shift K = mean of the row's
//     first min(G, len) samples:  d = x - K,  mean = K + sum(d)/n,
//     var = sum(d*d)/n - (sum(d)/n)^2.  The cancellation in the last line is bounded by
//     1 + (K - mean)^2 / var <= ~K4_ONE_PASS / G even when the first samples are all
//     outliers, which keeps the relative error of std below 1e-10;
//   * longer rows: two passes (mean first, then deviations from it), like np.std.
// min / max run on the samples' own type; a NaN anywhere makes both NaN like np.min / np.max.
#pragma once
#include "common.cuh"

constexpr int K4_ONE_PASS = 4096;
constexpr int K4_U = 4;  // independent loads per lane and trip

template <int G>
__device__ __forceinline__ double k4_group_sum(double v)
{
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v += __shfl_xor_sync(PP_FULL, v, d);
    return v;
}
template <int G>
__device__ __forceinline__ float k4_group_min(float v)
{
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v = fminf(v, __shfl_xor_sync(PP_FULL, v, d));
    return v;
}
template <int G>
__device__ __forceinline__ float k4_group_max(float v)
{
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(PP_FULL, v, d));
    return v;
}
template <int G>
__device__ __forceinline__ double k4_group_min(double v)
{
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v = fmin(v, __shfl_xor_sync(PP_FULL, v, d));
    return v;
}
template <int G>
__device__ __forceinline__ double k4_group_max(double v)
{
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(PP_FULL, v, d));
    return v;
}
__device__ __forceinline__ float k4_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ float k4_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double k4_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ double k4_max(double a, double b) { return fmax(a, b); }

// Rows k = group, group + ngroups, ...; every lane of a warp runs the same number of rows
// so that the full-mask shuffles stay convergent (inactive groups work on an empty row).
template <typename T, int G>
__device__ __forceinline__ void k4_rows(const T *__restrict__ samples, const int64_t *__restrict__ ev_base,
                                        const int64_t *__restrict__ ev_off, int64_t row0, int64_t rows, int64_t total,
                                        const int64_t *__restrict__ flat_start, const int *__restrict__ row_event,
                                        double *__restrict__ o_mean, double *__restrict__ o_std,
                                        double *__restrict__ o_min, double *__restrict__ o_max)
{
    const int gl = threadIdx.x & (G - 1);
    const int64_t group0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G;
    const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t warp_first = group0 - ((threadIdx.x & 31) / G);  // row of the warp's first group
    const T inf = (T)__longlong_as_double(0x7ff0000000000000LL);
    for (int64_t kw = row0 + warp_first; kw < rows; kw += ngroups) {  // rows [row0, rows)
        const int64_t k = kw + ((threadIdx.x & 31) / G);
        const bool live = k < rows;
        int64_t len = 0;
        const T *p = samples;
        if (live) {
            const int64_t f0 = flat_start[k];
            const int64_t f1 = (k + 1 < rows) ? flat_start[k + 1] : total;
            const int64_t ev = row_event ? (int64_t)row_event[k] : k;
            len = f1 - f0;
            p = samples + ev_base[ev] + (f0 - ev_off[ev]);
        }
        const int n0 = (int)(len < G ? len : G);
        // the lane's first K4_U samples are requested together (a row of ~130 samples is 4 trips of 8 lanes x 4
        // loads instead of 16 dependent round trips); the very first one also serves the shift K
        T xq[K4_U];
#pragma unroll
        for (int u = 0; u < K4_U; ++u) xq[u] = (gl + u * G) < len ? p[gl + u * G] : (T)0;
        double K = k4_group_sum<G>(gl < n0 ? (double)xq[0] : 0.0) / (double)(n0 > 0 ? n0 : 1);
        if (!(fabs(K) <= 1.7e308)) K = 0.0;  // inf / NaN among the first samples: plain sums below
        double s1 = 0.0, s2 = 0.0;
        T mn = inf, mx = -inf;
        int nan = 0;
        double mean, var;
        if (__all_sync(PP_FULL, len <= K4_ONE_PASS)) {
            for (int64_t j = gl; j < len; j += K4_U * G) {
                T xn[K4_U];
#pragma unroll
                for (int u = 0; u < K4_U; ++u) xn[u] = (j + (K4_U + u) * G) < len ? p[j + (K4_U + u) * G] : (T)0;
#pragma unroll
                for (int u = 0; u < K4_U; ++u) {
                    if (j + u * G < len) {
                        const T x = xq[u];
                        const double d = (double)x - K;
                        s1 += d;
                        s2 = fma(d, d, s2);
                        mn = k4_min(mn, x);
                        mx = k4_max(mx, x);
                        nan |= (x != x);
                    }
                }
#pragma unroll
                for (int u = 0; u < K4_U; ++u) xq[u] = xn[u];
            }
            s1 = k4_group_sum<G>(s1);
            s2 = k4_group_sum<G>(s2);
            const double m1 = s1 / (double)len;
            mean = K + m1;
            var = s2 / (double)len - m1 * m1;
        } else {
            for (int64_t j = gl; j < len; j += G) {
                const T x = p[j];
                s1 += (double)x;
                mn = k4_min(mn, x);
                mx = k4_max(mx, x);
                nan |= (x != x);
            }
            mean = k4_group_sum<G>(s1) / (double)len;
            for (int64_t j = gl; j < len; j += G) {
                const double d = (double)p[j] - mean;
                s2 = fma(d, d, s2);
            }
            var = k4_group_sum<G>(s2) / (double)len;
        }
        mn = k4_group_min<G>(mn);
        mx = k4_group_max<G>(mx);
#pragma unroll
        for (int d = G / 2; d > 0; d >>= 1) nan |= __shfl_xor_sync(PP_FULL, nan, d);
        if (live && gl == 0) {
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
            o_mean[k] = mean;
            o_std[k] = sqrt(var < 0.0 ? 0.0 : var);
            o_min[k] = nan ? qnan : (double)mn;
            o_max[k] = nan ? qnan : (double)mx;
        }
    }
}

// flat_start[k] .. flat_start[k+1] (or `total` for the last) bounds row k;
// row_event == nullptr means row k is event k (statistics of whole events).
template <typename T>
__global__ void __launch_bounds__(256)
k4_segment_stats(const T *__restrict__ samples, const int64_t *__restrict__ ev_base,
                 const int64_t *__restrict__ ev_off, const PPCounters *ctr, int rows_are_events,
                 const int64_t *__restrict__ flat_start, const int *__restrict__ row_event,
                 int64_t cap_rows, double *__restrict__ o_mean, double *__restrict__ o_std,
                 double *__restrict__ o_min, double *__restrict__ o_max)
{
    int64_t rows = rows_are_events ? (int64_t)ctr->n_events : (int64_t)ctr->n_segments;
    if (rows > cap_rows) rows = cap_rows;
    const int64_t total = (int64_t)ctr->n_event_samples;
    // segments: only the rows not finalised yet (the whole table unless the streamed pipeline exports per chunk)
    const int64_t row0 = rows_are_events ? 0 : (int64_t)ctr->seg_done;
    if (rows <= row0) return;
    if (total / rows < 512)
        k4_rows<T, 8>(samples, ev_base, ev_off, row0, rows, total, flat_start, row_event, o_mean, o_std, o_min, o_max);
    else
        k4_rows<T, 32>(samples, ev_base, ev_off, row0, rows, total, flat_start, row_event, o_mean, o_std, o_min, o_max);
}
