// prefix.cuh -- K2: FastStatSplit's cumsums (PyPore/cparsers.pyx:110-111).
//
//   c  = np.cumsum(x)          c2 = np.cumsum(x * x)        (fp64, restart per event)
//
// stored interleaved as double2 {c, c2} per event sample in "flat event space"
// (event e occupies [ev_off[e], ev_off[e+1])).
//
// np.cumsum is a strictly sequential fp64 accumulation and the split decisions
// downstream are compared bit-exactly, so two kernels cooperate:
//   * k2_prefix_tiled: a per-event tiled scan with decoupled look-back that ALSO
//     proves, with error-free TwoSum residuals, whether every addition it made was
//     exact.  If all of an event's additions were exact, every prefix value is the
//     exact real sum and therefore equals what the sequential order produces
//     (induction over i: c[i-1] exact and representable, c[i-1]+x[i] exact and
//     representable => fl() returns it).  ADC-quantised traces (real .abf data are
//     int16 counts, read_abf.py:208-210) always pass.
//   * k2_prefix_sequential: strict np.cumsum order, one thread per event, run for
//     the events the tiled scan flagged inexact (PP_PREFIX_AUTO) or for all
//     events (PP_PREFIX_SEQUENTIAL).
#pragma once
#include "common.cuh"

// One thread per event, strict left-to-right order.
__global__ void __launch_bounds__(128)
k2_prefix_sequential(PPSource src, const int64_t *__restrict__ ev_len, PPCounters *ctr,
                     const unsigned *__restrict__ inexact /* nullable: redo only flagged events */,
                     double2 *__restrict__ cc)
{
    const int64_t n_events = (int64_t)ctr->n_events;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_events;
         e += (int64_t)gridDim.x * blockDim.x) {
        if (inexact) {
            if (inexact[e] == 0u) continue;
            atomicAdd(&ctr->n_seq_redo, 1ull);
        }
        const int64_t len = ev_len[e];
        const int64_t off = src.ev_off[e];
        double c = 0.0, c2 = 0.0;
        int64_t j = 0;
        for (; j + 8 <= len; j += 8) {
            double xv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) xv[k] = pp_sample(src, e, j + k);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                c = __dadd_rn(c, xv[k]);
                c2 = __dadd_rn(c2, __dmul_rn(xv[k], xv[k]));
                cc[off + j + k] = make_double2(c, c2);
            }
        }
        for (; j < len; ++j) {
            const double xv = pp_sample(src, e, j);
            c = __dadd_rn(c, xv);
            c2 = __dadd_rn(c2, __dmul_rn(xv, xv));
            cc[off + j] = make_double2(c, c2);
        }
    }
}
