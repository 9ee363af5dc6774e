// stats.cuh -- K4: Segment.mean / std / min / max (PyPore/core.py:209-223) as a
// segmented reduction over flat event space: one warp per segment, fp64,
// two passes over the (cache-resident) samples so that std is computed from
// deviations, not from prefix-sum differences.
#pragma once
#include "common.cuh"

// flat_start[k] .. flat_start[k+1] (or `total` for the last) bounds row k;
// row_event == nullptr means row k is event k (statistics of whole events).
__global__ void __launch_bounds__(256)
k4_segment_stats(PPSource src, const PPCounters *ctr, int rows_are_events,
                 const int64_t *__restrict__ flat_start, const int *__restrict__ row_event,
                 int64_t cap_rows, double *__restrict__ o_mean, double *__restrict__ o_std,
                 double *__restrict__ o_min, double *__restrict__ o_max)
{
    int64_t rows = rows_are_events ? (int64_t)ctr->n_events : (int64_t)ctr->n_segments;
    if (rows > cap_rows) rows = cap_rows;
    const int64_t total = (int64_t)ctr->n_event_samples;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp0; k < rows; k += nwarps) {
        const int64_t f0 = flat_start[k];
        const int64_t f1 = (k + 1 < rows) ? flat_start[k + 1] : total;
        const int64_t ev = row_event ? (int64_t)row_event[k] : k;
        const int64_t j0 = f0 - src.ev_off[ev];
        const int64_t len = f1 - f0;
        double sum = 0.0;
        double mn = __longlong_as_double(0x7ff0000000000000LL);
        double mx = __longlong_as_double(0xfff0000000000000LL);
        for (int64_t j = lane; j < len; j += 32) {
            const double x = pp_sample(src, ev, j0 + j);
            sum += x;
            mn = fmin(mn, x);
            mx = fmax(mx, x);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            sum += __shfl_xor_sync(PP_FULL, sum, d);
            mn = fmin(mn, __shfl_xor_sync(PP_FULL, mn, d));
            mx = fmax(mx, __shfl_xor_sync(PP_FULL, mx, d));
        }
        const double mean = sum / (double)len;
        double ss = 0.0;
        for (int64_t j = lane; j < len; j += 32) {
            const double d = pp_sample(src, ev, j0 + j) - mean;
            ss += d * d;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(PP_FULL, ss, d);
        if (lane == 0) {
            const bool bad = !(sum == sum);  // NaN anywhere: np.min/np.max return NaN too
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
            o_mean[k] = mean;
            o_std[k] = sqrt(ss / (double)len);
            o_min[k] = bad ? qnan : mn;
            o_max[k] = bad ? qnan : mx;
        }
    }
}
