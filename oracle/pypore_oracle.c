/*
 * pypore_oracle.c -- CPU restatement ("port") of the PyPore segmentation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pypore_b200/ may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and there only as the checker.
 *
 * Parity pin: the reference ships no tests, golden vectors or fixtures
 * (SURVEY.md section 4).  This restatement is pinned instead against outputs of
 * the reference itself: (1) the unmodified cparsers.pyx compiled into
 * oracle/_ref by oracle/build_oracle.py, compared in tests/test_oracle.py, and
 * (2) fixtures under tests/golden/ produced by importing the Python reference
 * (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off (never -march=native / -mfma: FMA
 * contraction changes the rounding of var_c and of the gain).
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* lambda_event_parser.parse: PyPore/parsers.py:148-150                      */
/*   mask = where(current < thr, 1, 0); edges where mask[i] != mask[i+1];    */
/*   tics = [0] ++ (edge_index + 1) ++ [N]                                   */
/* Returns the number of tics written (runs = tics - 1).  NaN compares false */
/* and therefore counts as "above" (SURVEY App. A.1).                        */
/* ------------------------------------------------------------------------ */
int64_t orc_threshold_tics(const double *x, int64_t n, double thr,
                           int64_t *tics, int64_t cap)
{
    int64_t k = 0;
    if (cap < 2) return -1;
    tics[k++] = 0;
    for (int64_t i = 0; i + 1 < n; ++i) {
        int a = x[i] < thr, b = x[i + 1] < thr;
        if (a != b) {
            if (k + 1 >= cap) return -1;
            tics[k++] = i + 1;
        }
    }
    tics[k++] = n;
    return k;
}

/* Per-run min / max as the rules see them: PyPore/core.py:215-220 (np.min /  */
/* np.max over the piece; NaN propagates like numpy).                        */
void orc_run_minmax(const double *x, const int64_t *tics, int64_t ntics,
                    double *mn, double *mx)
{
    for (int64_t r = 0; r + 1 < ntics; ++r) {
        double lo = INFINITY, hi = -INFINITY;
        int nan = 0;
        for (int64_t i = tics[r]; i < tics[r + 1]; ++i) {
            double v = x[i];
            if (v != v) nan = 1;
            if (v < lo) lo = v;
            if (v > hi) hi = v;
        }
        mn[r] = nan ? NAN : lo;
        mx[r] = nan ? NAN : hi;
    }
}

/* ------------------------------------------------------------------------ */
/* FastStatSplit.parse cumsums: PyPore/cparsers.pyx:110-111                  */
/* np.cumsum is a strictly sequential fp64 accumulation.                     */
/* ------------------------------------------------------------------------ */
void orc_cumsum(const double *x, int64_t n, double *c, double *c2)
{
    double s = 0.0, s2 = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        s = s + x[i];
        s2 = s2 + x[i] * x[i];
        c[i] = s;
        c2[i] = s2;
    }
}

/* var_c: PyPore/cparsers.pyx:31-38.  "** 2" on a C double is x*x. */
static inline double var_c(int start, int end, const double *c, const double *c2)
{
    if (start == end) return 0.0;
    if (start == 0) {
        double m = c[end - 1] / end;
        return c2[end - 1] / end - m * m;
    }
    double m = (c[end - 1] - c[start - 1]) / (end - start);
    return (c2[end - 1] - c2[start - 1]) / (end - start) - m * m;
}

typedef struct {
    const double *c, *c2;
    int min_width, max_width, window_width;
    double min_gain;
    int *bp;          /* breakpoint output, in-order */
    int nbp, cap;
    int overflow;
    int64_t ncand;    /* candidate evaluations (SURVEY 8d "C") */
    int64_t nscan;    /* window scans */
    double *gain_out; /* optional: winning gain per emitted breakpoint (NaN for forced) */
    double *margin_out; /* optional: best - second best gain of the winning scan */
    double last_gain, last_margin;
    /* margin audit: the smallest distance of any decision from flipping */
    double min_split_margin;   /* over scans that split: best gain - max(second best gain, min_gain) */
    double min_nosplit_margin; /* over scans that did not: min_gain - best gain */
} split_ctx;

/* _best_split_stepwise: PyPore/cparsers.pyx:157-178 */
static int best_split_stepwise(split_ctx *S, int start, int end)
{
    const int mw = S->min_width;
    if (end - start <= 2 * mw) return -1;
    const double *c = S->c, *c2 = S->c2;
    double var_summed = (end - start) * log(var_c(start, end, c, c2));
    double min_gain = S->min_gain, second = -INFINITY, top = -INFINITY;
    int x = -1;
    S->nscan++;
    for (int i = start + mw; i < end + 1 - mw; ++i) {
        double low = (i - start) * log(var_c(start, i, c, c2));
        double high = (end - i) * log(var_c(i, end, c, c2));
        double gain = var_summed - (low + high);
        S->ncand++;
        if (gain > min_gain) {
            second = min_gain;
            min_gain = gain;
            x = i;
        } else if (gain > second) {
            second = gain;
        }
        if (gain > top) top = gain;
    }
    S->last_gain = min_gain;
    S->last_margin = min_gain - second;
    if (x >= 0) {
        if (S->last_margin < S->min_split_margin) S->min_split_margin = S->last_margin;
    } else if (S->min_gain - top < S->min_nosplit_margin) {
        S->min_nosplit_margin = S->min_gain - top;
    }
    return x;
}

static void emit(split_ctx *S, int at, double gain, double margin)
{
    if (S->nbp >= S->cap) { S->overflow = 1; return; }
    if (S->gain_out) S->gain_out[S->nbp] = gain;
    if (S->margin_out) S->margin_out[S->nbp] = margin;
    S->bp[S->nbp++] = at;
}

/* _recursive_split: PyPore/cparsers.pyx:180-203.  The list concatenation
 * rec(start,split)+[split]+rec(split,end) is an in-order traversal: recurse on
 * the left part, emit, then continue with the right part (iteratively, so the
 * long right spine of SURVEY App. A.3 does not deepen the C stack). */
static void recursive_split(split_ctx *S, int start, int end)
{
    const int mw = S->min_width, W = S->window_width, MW = S->max_width;
    for (;;) {
        int split_at = -1, forced_right_only = 0;
        double g = NAN, m = NAN;
        for (long ps = start; ps < (long)end - 2L * mw; ps += W / 2) {
            if (ps > (long)start + MW) {
                split_at = (start + MW < end - mw) ? start + MW : end - mw;
                forced_right_only = 1;
                break;
            }
            int pe = (end < ps + W) ? end : (int)(ps + W);
            split_at = best_split_stepwise(S, (int)ps, pe);
            if (split_at >= 0) { g = S->last_gain; m = S->last_margin; break; }
        }
        if (forced_right_only) {
            emit(S, split_at, NAN, NAN);
            start = split_at;
            continue;
        }
        if (split_at == -1) {
            if (end - start <= MW) return;
            split_at = (start + MW < end - mw) ? start + MW : end - mw;
        }
        recursive_split(S, start, split_at);
        emit(S, split_at, g, m);
        start = split_at;
    }
}

/* The scoring loop shared by _best_single_split (PyPore/cparsers.pyx:142-151, start = 0,
 * end = len-1, candidates 2 .. end-3) and _best_split_stepwise_score (cparsers.pyx:234-240,
 * candidates start+min_width .. end-min_width): gain of every candidate first .. last of window
 * [start, end) in the reference's operation order.  out[i - first] = gain(i). */
void orc_window_gains(const double *c, const double *c2, int start, int end, int first, int last,
                      double *out)
{
    double var_summed = (end - start) * log(var_c(start, end, c, c2));
    for (int i = first; i <= last; ++i) {
        double low_var_summed = (i - start) * log(var_c(start, i, c, c2));
        double high_var_summed = (end - i) * log(var_c(i, end, c, c2));
        out[i - first] = var_summed - (low_var_summed + high_var_summed);
    }
}

/* FastStatSplit.parse minus the Python object construction
 * (PyPore/cparsers.pyx:103-118).  x must be float64 like the reference
 * requires.  Returns the number of breakpoints (segments = n + 1), or -1 on
 * output overflow / allocation failure.  stats_out[0] = candidate
 * evaluations, stats_out[1] = window scans (either may be NULL). */
int orc_statsplit(const double *x, int n, int min_width, int max_width,
                  int window_width, double min_gain, int *bp, int cap,
                  int64_t *stats_out, double *gain_out, double *margin_out)
{
    split_ctx S;
    double *c = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *c2 = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (!c || !c2) { free(c); free(c2); return -1; }
    orc_cumsum(x, n, c, c2);
    memset(&S, 0, sizeof S);
    S.c = c; S.c2 = c2;
    S.min_width = min_width; S.max_width = max_width; S.window_width = window_width;
    S.min_gain = min_gain;
    S.bp = bp; S.cap = cap;
    S.gain_out = gain_out; S.margin_out = margin_out;
    S.min_split_margin = INFINITY; S.min_nosplit_margin = INFINITY;
    recursive_split(&S, 0, n);
    free(c); free(c2);
    if (stats_out) { stats_out[0] = S.ncand; stats_out[1] = S.nscan; }
    return S.overflow ? -1 : S.nbp;
}

/* Margin audit of one event: out[0] = smallest (best gain - runner-up) over the scans that
 * split, the runner-up being the second best candidate or min_gain itself, whichever is closer; out[1] = smallest
 * (min_gain - best gain) over the scans that did not split; out[2] = window scans.  +inf where there was none.
 * A decision can only differ between two correct implementations of the reference's arithmetic if its margin is
 * below their rounding differences (libm log: a few ulp, i.e. ~1e-10 on a gain). */
int orc_statsplit_audit(const double *x, int n, int min_width, int max_width, int window_width, double min_gain,
                        double *out)
{
    split_ctx S;
    const int cap = n / (min_width > 0 ? min_width : 1) * 2 + 16;
    double *c = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *c2 = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    int *bp = (int *)malloc(sizeof(int) * (size_t)cap);
    if (!c || !c2 || !bp) { free(c); free(c2); free(bp); return -1; }
    orc_cumsum(x, n, c, c2);
    memset(&S, 0, sizeof S);
    S.c = c; S.c2 = c2;
    S.min_width = min_width; S.max_width = max_width; S.window_width = window_width;
    S.min_gain = min_gain;
    S.bp = bp; S.cap = cap;
    S.min_split_margin = INFINITY; S.min_nosplit_margin = INFINITY;
    recursive_split(&S, 0, n);
    out[0] = S.min_split_margin; out[1] = S.min_nosplit_margin; out[2] = (double)S.nscan;
    free(c); free(c2); free(bp);
    return S.overflow ? -1 : 0;
}

/* Same on many events of one trace; events are independent
 * (SURVEY 8e), so the generous multi-core CPU baseline is an OpenMP loop.
 * bp_off[e] .. bp_off[e+1] bounds event e's slice of bp (capacity per event =
 * len/min_width + 2).  nbp[e] receives the count.  Returns 0 or -1. */
int orc_statsplit_events(const double *x, const int64_t *ev_start,
                         const int64_t *ev_len, int64_t n_events, int min_width,
                         int max_width, int window_width, double min_gain,
                         int *bp, const int64_t *bp_off, int *nbp,
                         int64_t *ncand_total, int threads)
{
    int bad = 0;
    int64_t ncand = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+:ncand) reduction(|:bad)
#endif
    for (int64_t e = 0; e < n_events; ++e) {
        int64_t st[2];
        int r = orc_statsplit(x + ev_start[e], (int)ev_len[e], min_width, max_width,
                              window_width, min_gain, bp + bp_off[e],
                              (int)(bp_off[e + 1] - bp_off[e]), st, NULL, NULL);
        if (r < 0) { bad |= 1; nbp[e] = 0; }
        else nbp[e] = r;
        ncand += st[0];
    }
    if (ncand_total) *ncand_total = ncand;
    return bad ? -1 : 0;
}

/* ------------------------------------------------------------------------ */
/* Event.filter: PyPore/DataTypes.py:270-271 calls scipy.signal.filtfilt     */
/* (scipy is not vendored in the reference; this restates the published      */
/* algorithm of scipy 1.x filtfilt(method='pad', padtype='odd',              */
/* padlen=3*max(len(a),len(b))) with lfilter_zi initial conditions and a     */
/* direct-form-II-transposed lfilter).                      */
/* b, a have ncoef entries each (a[0] must be 1 after normalisation), zi has */
/* ncoef-1 entries.  Returns 0, or -1 if n <= padlen (scipy raises           */
/* ValueError) or on allocation failure.                                     */
/* ------------------------------------------------------------------------ */
static void lfilter_df2t(const double *b, const double *a, int nc,
                         const double *x, int64_t n, double *z, double *y)
{
    for (int64_t i = 0; i < n; ++i) {
        double xi = x[i];
        double yi = b[0] * xi + (nc > 1 ? z[0] : 0.0);
        for (int k = 1; k < nc - 1; ++k)
            z[k - 1] = b[k] * xi + z[k] - a[k] * yi;
        if (nc > 1)
            z[nc - 2] = b[nc - 1] * xi - a[nc - 1] * yi;
        y[i] = yi;
    }
}

int orc_filtfilt(const double *b, const double *a, const double *zi, int ncoef,
                 const double *x, int64_t n, double *out)
{
    const int64_t padlen = 3 * (int64_t)ncoef;
    if (n <= padlen) return -1;
    const int64_t m = n + 2 * padlen;
    double *ext = (double *)malloc(sizeof(double) * (size_t)m);
    double *y = (double *)malloc(sizeof(double) * (size_t)m);
    double z[32];
    if (!ext || !y || ncoef > 32) { free(ext); free(y); return -1; }
    /* odd extension: 2*x[0] - x[padlen..1], x, 2*x[-1] - x[-2..-(padlen+1)] */
    for (int64_t i = 0; i < padlen; ++i) {
        ext[i] = 2.0 * x[0] - x[padlen - i];
        ext[padlen + n + i] = 2.0 * x[n - 1] - x[n - 2 - i];
    }
    memcpy(ext + padlen, x, sizeof(double) * (size_t)n);
    /* forward */
    for (int k = 0; k < ncoef - 1; ++k) z[k] = zi[k] * ext[0];
    lfilter_df2t(b, a, ncoef, ext, m, z, y);
    /* backward: reverse, filter with zi * y[-1], reverse */
    for (int64_t i = 0; i < m / 2; ++i) { double t = y[i]; y[i] = y[m - 1 - i]; y[m - 1 - i] = t; }
    for (int k = 0; k < ncoef - 1; ++k) z[k] = zi[k] * y[0];
    lfilter_df2t(b, a, ncoef, y, m, z, ext);
    for (int64_t i = 0; i < n; ++i) out[i] = ext[m - 1 - padlen - i];
    free(ext); free(y);
    return 0;
}
