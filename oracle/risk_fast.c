/*
 * risk_fast.c -- the CPU baseline of the risk stages, written the way a performance-minded CPU author would:
 * SIMD across trials, register-blocked over portfolios, threshold-prefiltered selection.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h); PARITY UNPINNED like risk_oracle.c (the reference has no such code).
 *
 * risk_oracle.c states DESIGN.md section 3 literally (one dependent fma chain per valuation: latency-bound, ~4e7
 * valuations/s per core) and stays the CHECKER.  This file computes the same results fast and is what bench.py times
 * as `cpu_baseline` / `--impl reference`, so that GPU/CPU ratios are quoted against an honest CPU implementation:
 *
 *   valuation  each SIMD lane runs the k-ascending fma chain of ONE trial (acc = +0.0; acc = fma(E[p][k], S[n][k], acc)),
 *              so every V[p][n] is bit-identical to orc_value's; 4 portfolios x 2 vectors of trials are in flight per
 *              inner loop (8 independent chains hide the FMA latency), scenarios are re-tiled once to
 *              [trial block][k][lanes] so the k loop streams one contiguous 4 KB tile;
 *   moments    shifted one-pass sums in double, folded in while a tile's values are still in L1 (tested at 1e-9 against
 *              the long-double oracle);
 *   VaR/tail   a threshold from the row's first 4 096 trials, candidates (~1.5 t) collected on the fly, exact selection
 *              among them on (loss, trial); rows whose threshold was too tight are redone from a full V row; V itself is
 *              never materialised and the scenario tiles are walked in L2-sized chunks shared by 32 portfolios;
 *   CVaR       the oracle's own summation order (ascending trial), so it is bit-identical too.
 *
 * AVX2+FMA always (the library is built -march=x86-64-v3); an AVX-512 body is chosen at run time when the host has it.
 */
#include "oracle.h"

#include <immintrin.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double loss; int32_t trial; } lt_pair;

static inline int lt_less(const lt_pair *a, const lt_pair *b)
{
    return a->loss < b->loss || (a->loss == b->loss && a->trial < b->trial);
}

static void lt_select(lt_pair *a, int64_t n, int64_t kth)
{
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        const int64_t mid = lo + (hi - lo) / 2;
        if (lt_less(&a[mid], &a[lo])) { lt_pair t = a[mid]; a[mid] = a[lo]; a[lo] = t; }
        if (lt_less(&a[hi], &a[lo])) { lt_pair t = a[hi]; a[hi] = a[lo]; a[lo] = t; }
        if (lt_less(&a[hi], &a[mid])) { lt_pair t = a[hi]; a[hi] = a[mid]; a[mid] = t; }
        const lt_pair pivot = a[mid];
        int64_t i = lo, j = hi;
        while (i <= j) {
            while (lt_less(&a[i], &pivot)) ++i;
            while (lt_less(&pivot, &a[j])) --j;
            if (i <= j) { lt_pair t = a[i]; a[i] = a[j]; a[j] = t; ++i; --j; }
        }
        if (kth <= j) hi = j;
        else if (kth >= i) lo = i;
        else return;
    }
}

static int cmp_i32(const void *a, const void *b)
{
    const int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* ------------------------------------------------------------------ scenario tiles
 * T[blk][k][TB] = S[blk*TB + lane][k], zero-padded past N (TB = 16 trials per tile). */
#define TB 16

static double *tile_scenarios(const double *S, int64_t N, int32_t F, int64_t *nblk_out)
{
    const int64_t nblk = (N + TB - 1) / TB;
    double *T = (double *)aligned_alloc(64, (size_t)(nblk > 0 ? nblk : 1) * F * TB * sizeof(double));
    if (!T) return NULL;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < nblk; ++b)
        for (int32_t k = 0; k < F; ++k)
            for (int l = 0; l < TB; ++l) {
                const int64_t n = b * TB + l;
                T[(b * F + k) * TB + l] = n < N ? S[n * F + k] : 0.0;
            }
    *nblk_out = nblk;
    return T;
}

/* ------------------------------------------------------------------ valuation microkernels
 * V rows of `np` (1..4) portfolios over every tile; Vrow[q] has room for nblk*TB doubles. */
static void value_rows_avx2(const double *E, int32_t F, const double *T, int64_t nblk, int np, double *const Vrow[4])
{
    for (int64_t b = 0; b < nblk; ++b) {
        const double *t = T + b * F * TB;
        for (int h = 0; h < TB; h += 8) { /* two 4-lane vectors at a time: 8 chains for 4 portfolios */
            __m256d a00 = _mm256_setzero_pd(), a01 = a00, a10 = a00, a11 = a00, a20 = a00, a21 = a00, a30 = a00, a31 = a00;
            for (int32_t k = 0; k < F; ++k) {
                const __m256d s0 = _mm256_load_pd(t + k * TB + h), s1 = _mm256_load_pd(t + k * TB + h + 4);
                const __m256d e0 = _mm256_broadcast_sd(E + k);
                a00 = _mm256_fmadd_pd(e0, s0, a00); a01 = _mm256_fmadd_pd(e0, s1, a01);
                if (np > 1) { const __m256d e1 = _mm256_broadcast_sd(E + F + k); a10 = _mm256_fmadd_pd(e1, s0, a10); a11 = _mm256_fmadd_pd(e1, s1, a11); }
                if (np > 2) { const __m256d e2 = _mm256_broadcast_sd(E + 2 * F + k); a20 = _mm256_fmadd_pd(e2, s0, a20); a21 = _mm256_fmadd_pd(e2, s1, a21); }
                if (np > 3) { const __m256d e3 = _mm256_broadcast_sd(E + 3 * F + k); a30 = _mm256_fmadd_pd(e3, s0, a30); a31 = _mm256_fmadd_pd(e3, s1, a31); }
            }
            const int64_t o = b * TB + h;
            _mm256_storeu_pd(Vrow[0] + o, a00); _mm256_storeu_pd(Vrow[0] + o + 4, a01);
            if (np > 1) { _mm256_storeu_pd(Vrow[1] + o, a10); _mm256_storeu_pd(Vrow[1] + o + 4, a11); }
            if (np > 2) { _mm256_storeu_pd(Vrow[2] + o, a20); _mm256_storeu_pd(Vrow[2] + o + 4, a21); }
            if (np > 3) { _mm256_storeu_pd(Vrow[3] + o, a30); _mm256_storeu_pd(Vrow[3] + o + 4, a31); }
        }
    }
}

__attribute__((target("avx512f")))
static void value_rows_avx512(const double *E, int32_t F, const double *T, int64_t nblk, int np, double *const Vrow[4])
{
    for (int64_t b = 0; b < nblk; ++b) {
        const double *t = T + b * F * TB;
        __m512d a00 = _mm512_setzero_pd(), a01 = a00, a10 = a00, a11 = a00, a20 = a00, a21 = a00, a30 = a00, a31 = a00;
        for (int32_t k = 0; k < F; ++k) {
            const __m512d s0 = _mm512_load_pd(t + k * TB), s1 = _mm512_load_pd(t + k * TB + 8);
            const __m512d e0 = _mm512_set1_pd(E[k]);
            a00 = _mm512_fmadd_pd(e0, s0, a00); a01 = _mm512_fmadd_pd(e0, s1, a01);
            if (np > 1) { const __m512d e1 = _mm512_set1_pd(E[F + k]); a10 = _mm512_fmadd_pd(e1, s0, a10); a11 = _mm512_fmadd_pd(e1, s1, a11); }
            if (np > 2) { const __m512d e2 = _mm512_set1_pd(E[2 * F + k]); a20 = _mm512_fmadd_pd(e2, s0, a20); a21 = _mm512_fmadd_pd(e2, s1, a21); }
            if (np > 3) { const __m512d e3 = _mm512_set1_pd(E[3 * F + k]); a30 = _mm512_fmadd_pd(e3, s0, a30); a31 = _mm512_fmadd_pd(e3, s1, a31); }
        }
        const int64_t o = b * TB;
        _mm512_storeu_pd(Vrow[0] + o, a00); _mm512_storeu_pd(Vrow[0] + o + 8, a01);
        if (np > 1) { _mm512_storeu_pd(Vrow[1] + o, a10); _mm512_storeu_pd(Vrow[1] + o + 8, a11); }
        if (np > 2) { _mm512_storeu_pd(Vrow[2] + o, a20); _mm512_storeu_pd(Vrow[2] + o + 8, a21); }
        if (np > 3) { _mm512_storeu_pd(Vrow[3] + o, a30); _mm512_storeu_pd(Vrow[3] + o + 8, a31); }
    }
}

typedef void (*value_rows_fn)(const double *, int32_t, const double *, int64_t, int, double *const[4]);

static value_rows_fn pick_value_rows(void)
{
    const char *force = getenv("ORC_FAST_ISA"); /* "avx2" pins the narrow body (tests run both) */
    if (force && strcmp(force, "avx2") == 0) return value_rows_avx2;
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f") ? value_rows_avx512 : value_rows_avx2;
}

const char *orc_fast_isa(void) { return pick_value_rows() == value_rows_avx512 ? "avx512f" : "avx2+fma"; }

void orc_value_fast(const double *E, const double *S, int64_t P, int64_t N, int32_t F, double *V, int nthreads)
{
    int64_t nblk = 0;
#ifdef _OPENMP
    if (nthreads > 0) { omp_set_dynamic(0); omp_set_num_threads(nthreads); }
#endif
    double *T = tile_scenarios(S, N, F, &nblk);
    const value_rows_fn fn = pick_value_rows();
#pragma omp parallel
    {
        double *buf = (double *)aligned_alloc(64, (size_t)4 * (nblk > 0 ? nblk : 1) * TB * sizeof(double));
#pragma omp for schedule(dynamic, 4)
        for (int64_t p = 0; p < P; p += 4) {
            const int np = (int)(P - p < 4 ? P - p : 4);
            double *rows[4] = { buf, buf + nblk * TB, buf + 2 * nblk * TB, buf + 3 * nblk * TB };
            fn(E + p * F, F, T, nblk, np, rows);
            for (int q = 0; q < np; ++q) memcpy(V + (p + q) * N, rows[q], (size_t)N * sizeof(double));
        }
        free(buf);
    }
    free(T);
}

/* ------------------------------------------------------------------ fused streaming statistics
 * orc_risk_fast never materialises a V row.  Per task a thread owns a GROUP of portfolios and walks the scenario tiles in
 * L2-sized chunks; inside a chunk every quad of portfolios runs the microkernel over the chunk's tiles and, while the 16
 * values of a tile are still in a 128-byte line of L1, folds them into the row's shifted moment sums and compares them with
 * the row's tail threshold (candidates are appended scalar: ~1.5 % of the values).  The threshold is the r-th largest loss
 * of the row's first n0 trials (phase A), r chosen so that ~1.5 t candidates are expected; a row that ends with fewer than
 * t candidates (or overflows its buffer) is redone exactly from a full V row (`row_exact`). */
#define GROUP 32      /* portfolios per task                                  */
#define N0    4096    /* trials of the threshold sample (phase A)             */

static double nth_largest(double *a, int64_t n, int64_t r) /* r-th largest (0-based) by quickselect; a is permuted */
{
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        const double pivot = a[lo + (hi - lo) / 2];
        int64_t i = lo, j = hi;
        while (i <= j) {
            while (a[i] > pivot) ++i;
            while (a[j] < pivot) --j;
            if (i <= j) { const double t = a[i]; a[i] = a[j]; a[j] = t; ++i; --j; }
        }
        if (r <= j) hi = j;
        else if (r >= i) lo = i;
        else break;
    }
    return a[r];
}

typedef struct {
    __m256d vsd, vsdd;      /* lane sums of (v - shift) and (v - shift)^2        */
    double shift;
    double vthr;            /* candidates: V <= vthr  (<=> loss >= -vthr)         */
    int64_t cnt;
    int overflow;
    lt_pair *cand;
} row_state;

/* one tile (TB values of one row, trials n0..n0+TB, zero-padded past N) into the row's running statistics */
static inline void fold_tile(row_state *r, const double *v, int64_t n0, int64_t N, int64_t cap)
{
    const __m256d sh = _mm256_set1_pd(r->shift), th = _mm256_set1_pd(r->vthr);
    if (n0 + TB <= N) {
        __m256d sd = r->vsd, sdd = r->vsdd;
        int mask = 0;
        for (int h = 0; h < TB; h += 4) {
            const __m256d x = _mm256_load_pd(v + h), d = _mm256_sub_pd(x, sh);
            sd = _mm256_add_pd(sd, d);
            sdd = _mm256_fmadd_pd(d, d, sdd);
            mask |= _mm256_movemask_pd(_mm256_cmp_pd(x, th, _CMP_LE_OQ)) << h;
        }
        r->vsd = sd; r->vsdd = sdd;
        while (mask) {
            const int l = __builtin_ctz(mask);
            mask &= mask - 1;
            if (r->cnt < cap) { r->cand[r->cnt].loss = 0.0 - v[l]; r->cand[r->cnt].trial = (int32_t)(n0 + l); ++r->cnt; }
            else r->overflow = 1;
        }
        return;
    }
    /* the last, partial tile: lane 0 of the sums takes the scalars */
    double sd = 0.0, sdd = 0.0;
    for (int l = 0; l < (int)(N - n0); ++l) {
        const double d = v[l] - r->shift;
        sd += d; sdd += d * d;
        if (v[l] <= r->vthr) {
            if (r->cnt < cap) { r->cand[r->cnt].loss = 0.0 - v[l]; r->cand[r->cnt].trial = (int32_t)(n0 + l); ++r->cnt; }
            else r->overflow = 1;
        }
    }
    r->vsd = _mm256_add_pd(r->vsd, _mm256_set_pd(0.0, 0.0, 0.0, sd));
    r->vsdd = _mm256_add_pd(r->vsdd, _mm256_set_pd(0.0, 0.0, 0.0, sdd));
}

static inline double hsum4(__m256d v)
{
    double l[4];
    _mm256_storeu_pd(l, v);
    return (l[0] + l[1]) + (l[2] + l[3]);
}

/* the t largest (loss, trial) pairs of a[0..cnt): VaR pair, breach list (ascending trial), CVaR.
 * bits: N/64 + 1 words of scratch, pref: as many int32. */
static void tail_from_pairs(lt_pair *a, int64_t cnt, int64_t t, int64_t N, double *var_loss, int32_t *var_trial, double *cvar,
                            int32_t *idx, double *byrank, uint64_t *bits, int32_t *pref)
{
    lt_select(a, cnt, cnt - t);
    *var_loss = a[cnt - t].loss;
    *var_trial = a[cnt - t].trial;
    /* ascending trial order by counting: one bit per tail trial, popcount prefix per word, rank = prefix + bits below */
    const int64_t nw = (N + 63) / 64;
    memset(bits, 0, (size_t)nw * sizeof(uint64_t));
    const lt_pair *tl = a + (cnt - t);
    for (int64_t i = 0; i < t; ++i) bits[tl[i].trial >> 6] |= 1ull << (tl[i].trial & 63);
    int32_t run = 0;
    for (int64_t w = 0; w < nw; ++w) { pref[w] = run; run += __builtin_popcountll(bits[w]); }
    for (int64_t i = 0; i < t; ++i) {
        const int32_t tr = tl[i].trial;
        const int32_t rk = pref[tr >> 6] + __builtin_popcountll(bits[tr >> 6] & ((1ull << (tr & 63)) - 1));
        idx[rk] = tr; byrank[rk] = tl[i].loss;
    }
    double c = 0.0; /* plain double sum of the tail losses in ASCENDING TRIAL order (the oracle's order) */
    for (int64_t i = 0; i < t; ++i) c += byrank[i];
    *cvar = c / (double)t;
}

int64_t orc_risk_fast(const double *E, const double *S, int64_t P, int64_t N, int32_t F, double alpha,
                      int64_t p0, int64_t p1,
                      double *mean, double *variance, double *var_loss, int32_t *var_trial, double *cvar,
                      int32_t *breach_idx, int nthreads)
{
    (void)P;
    const int64_t t = orc_tail_size(N, alpha);
    if (N <= 0 || p1 <= p0) return t;
#ifdef _OPENMP
    if (nthreads > 0) { omp_set_dynamic(0); omp_set_num_threads(nthreads); }
#endif
    int64_t nblk = 0;
    double *T = tile_scenarios(S, N, F, &nblk);
    const value_rows_fn fn = pick_value_rows();
    const int64_t n0 = N < N0 ? N : N0, nblk0 = (n0 + TB - 1) / TB;
    /* threshold rank in the sample: expected tail share + 3 sigma + 2 */
    const double ts = (double)t * (double)n0 / (double)N;
    int64_t r0 = (int64_t)(ts + 3.0 * sqrt(ts) + 2.0);
    const int filtered = N >= 4 * N0 && r0 * 4 < n0; /* small problems: every trial is a candidate */
    if (r0 >= n0) r0 = n0 - 1;
    const int64_t cap = filtered ? (int64_t)(3.0 * (double)r0 / (double)n0 * (double)N) + 4 * t + 1024 : N;
    /* chunk of tiles that stays in L2 next to the candidate buffers: ~256 KB of scenarios */
    int64_t chunk = (256 * 1024) / ((int64_t)F * TB * 8);
    if (chunk < 1) chunk = 1;
    const int64_t ngroups = (p1 - p0 + GROUP - 1) / GROUP;
#pragma omp parallel
    {
        double *v0 = (double *)aligned_alloc(64, (size_t)4 * (nblk0 > chunk ? nblk0 : chunk) * TB * sizeof(double));
        double *full = NULL; /* a whole V row, only for rows that need the exact redo */
        double *smp = (double *)malloc((size_t)n0 * sizeof(double));
        lt_pair *cands = (lt_pair *)malloc((size_t)GROUP * cap * sizeof(lt_pair));
        double *byrank = (double *)malloc((size_t)t * sizeof(double));
        uint64_t *bits = (uint64_t *)malloc((size_t)(N / 64 + 2) * sizeof(uint64_t));
        int32_t *pref = (int32_t *)malloc((size_t)(N / 64 + 2) * sizeof(int32_t));
        int32_t *idx = (int32_t *)malloc((size_t)t * sizeof(int32_t));
        row_state st[GROUP];
#pragma omp for schedule(dynamic, 1)
        for (int64_t g = 0; g < ngroups; ++g) {
            const int64_t gp0 = p0 + g * GROUP;
            const int gn = (int)(p1 - gp0 < GROUP ? p1 - gp0 : GROUP);
            /* ---- phase A: shift and threshold of every row from its first n0 trials */
            for (int q = 0; q < gn; q += 4) {
                const int np = gn - q < 4 ? gn - q : 4;
                double *rows[4] = { v0, v0 + nblk0 * TB, v0 + 2 * nblk0 * TB, v0 + 3 * nblk0 * TB };
                fn(E + (gp0 + q) * F, F, T, nblk0, np, rows);
                for (int j = 0; j < np; ++j) {
                    row_state *r = &st[q + j];
                    double s = 0.0;
                    for (int64_t n = 0; n < n0; ++n) s += rows[j][n];
                    r->shift = s / (double)n0; r->vsd = r->vsdd = _mm256_setzero_pd(); r->cnt = 0; r->overflow = 0;
                    r->cand = cands + (int64_t)(q + j) * cap;
                    if (filtered) {
                        for (int64_t n = 0; n < n0; ++n) smp[n] = 0.0 - rows[j][n];
                        r->vthr = 0.0 - nth_largest(smp, n0, r0);
                    } else
                        r->vthr = INFINITY;
                }
            }
            /* ---- phase B: all trials, chunk by chunk (the chunk's tiles stay in L2 for the whole group) */
            for (int64_t b0 = 0; b0 < nblk; b0 += chunk) {
                const int64_t nb = nblk - b0 < chunk ? nblk - b0 : chunk;
                for (int q = 0; q < gn; q += 4) {
                    const int np = gn - q < 4 ? gn - q : 4;
                    double *rows[4] = { v0, v0 + nb * TB, v0 + 2 * nb * TB, v0 + 3 * nb * TB };
                    fn(E + (gp0 + q) * F, F, T + b0 * F * TB, nb, np, rows);
                    for (int j = 0; j < np; ++j)
                        for (int64_t b = 0; b < nb; ++b) fold_tile(&st[q + j], rows[j] + b * TB, (b0 + b) * TB, N, cap);
                }
            }
            /* ---- results */
            for (int q = 0; q < gn; ++q) {
                row_state *r = &st[q];
                const int64_t o = gp0 - p0 + q;
                lt_pair *a = r->cand;
                int64_t cnt = r->cnt;
                const double sd = hsum4(r->vsd), sdd = hsum4(r->vsdd);
                double mu = r->shift + sd / (double)N;
                double var = (sdd - sd * sd / (double)N) / (double)N;
                if (cnt < t || r->overflow || !(var == var)) {
                    /* the sample misjudged this row (adversarial trial order, NaNs): exact redo from the full row */
                    if (!full) full = (double *)aligned_alloc(64, (size_t)nblk * TB * sizeof(double));
                    double *rows[4] = { full, full, full, full };
                    fn(E + (gp0 + q) * F, F, T, nblk, 1, rows);
                    a = (lt_pair *)malloc((size_t)N * sizeof(lt_pair));
                    for (int64_t n = 0; n < N; ++n) { a[n].loss = 0.0 - full[n]; a[n].trial = (int32_t)n; }
                    cnt = N;
                }
                double vl, cv;
                int32_t vtr;
                tail_from_pairs(a, cnt, t, N, &vl, &vtr, &cv, idx, byrank, bits, pref);
                if (a != r->cand) free(a);
                if (var < 0.0) var = 0.0;
                if (mean) mean[o] = mu;
                if (variance) variance[o] = var;
                if (var_loss) var_loss[o] = vl;
                if (var_trial) var_trial[o] = vtr;
                if (cvar) cvar[o] = cv;
                if (breach_idx) memcpy(breach_idx + o * t, idx, (size_t)t * sizeof(int32_t));
            }
        }
        free(v0); free(full); free(smp); free(cands); free(byrank); free(bits); free(pref); free(idx);
    }
    free(T);
    return t;
}
