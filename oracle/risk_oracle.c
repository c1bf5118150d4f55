/*
 * risk_oracle.c -- CPU oracle for valuation / moments / VaR / CVaR / breach lists.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * PARITY UNPINNED: the reference (mapr-demos/monte-carlo-portfolios) contains no
 * valuation, statistics or breach code at all (no file even reads macro.json;
 * SURVEY.md section 0).  This file therefore implements DESIGN.md section 3, the
 * specification this repository wrote for those stages, in the most literal way:
 * a scalar fma chain, an exact order statistic by selection on (loss, trial), and
 * long-double moments.  The CUDA path is checked against it; neither is checked
 * against reference behaviour because there is none.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* launchers such as torchrun export OMP_NUM_THREADS=1; an explicit request must win over the environment */
static void orc_use_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) { omp_set_dynamic(0); omp_set_num_threads(n); }
#else
    (void)n;
#endif
}

/* DESIGN.md 3.1:  acc = +0.0; for k ascending: acc = fma(E[p][k], S[n][k], acc) */
static inline double value_one(const double *e, const double *s, int32_t F)
{
    double acc = 0.0;
    for (int32_t k = 0; k < F; ++k) acc = fma(e[k], s[k], acc);
    return acc;
}

void orc_value(const double *E, const double *S, int64_t P, int64_t N, int32_t F, double *V, int nthreads)
{
orc_use_threads(nthreads);
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t p = 0; p < P; ++p)
        for (int64_t n = 0; n < N; ++n) V[p * N + n] = value_one(E + p * F, S + n * F, F);
}

/* DESIGN.md 3.4:  k = ceil(alpha*N - 1e-7) clamped to [1,N];  tail size t = N-k+1 */
int64_t orc_tail_size(int64_t N, double alpha)
{
    if (N <= 0) return 0;
    int64_t k = (int64_t)ceil(alpha * (double)N - 1e-7);
    if (k < 1) k = 1;
    if (k > N) k = N;
    return N - k + 1;
}

typedef struct { double loss; int32_t trial; } lt_pair;

static inline int lt_less(const lt_pair *a, const lt_pair *b)
{
    return a->loss < b->loss || (a->loss == b->loss && a->trial < b->trial);
}

/* quickselect: afterwards a[kth] is the kth smallest (0-based) and a[kth..] >= it */
static void lt_select(lt_pair *a, int64_t n, int64_t kth)
{
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        const int64_t mid = lo + (hi - lo) / 2;
        /* median of three to the middle */
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

int64_t orc_risk(const double *E, const double *S, int64_t P, int64_t N, int32_t F, double alpha,
                 int64_t p0, int64_t p1,
                 double *mean, double *variance, double *var_loss, int32_t *var_trial, double *cvar,
                 int32_t *breach_idx, int nthreads)
{
    (void)P;
    const int64_t t = orc_tail_size(N, alpha);
    const int64_t k = N - t + 1; /* 1-based rank of the VaR order statistic */
orc_use_threads(nthreads);
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        double *V = (double *)malloc((size_t)N * sizeof(double));
        lt_pair *a = (lt_pair *)malloc((size_t)N * sizeof(lt_pair));
        int32_t *idx = (int32_t *)malloc((size_t)t * sizeof(int32_t));
#pragma omp for schedule(dynamic, 4)
        for (int64_t p = p0; p < p1; ++p) {
            const int64_t q = p - p0;
            for (int64_t n = 0; n < N; ++n) V[n] = value_one(E + p * F, S + n * F, F);
            /* DESIGN.md 3.3 moments of V (population variance), long double two-pass */
            long double s = 0.0L;
            for (int64_t n = 0; n < N; ++n) s += (long double)V[n];
            const long double mu = s / (long double)N;
            long double m2 = 0.0L;
            for (int64_t n = 0; n < N; ++n) { const long double d = (long double)V[n] - mu; m2 += d * d; }
            if (mean) mean[q] = (double)mu;
            if (variance) variance[q] = (double)(m2 / (long double)N);
            /* DESIGN.md 3.2 loss = 0.0 - V (canonical +0.0) */
            for (int64_t n = 0; n < N; ++n) { a[n].loss = 0.0 - V[n]; a[n].trial = (int32_t)n; }
            lt_select(a, N, k - 1);
            if (var_loss) var_loss[q] = a[k - 1].loss;
            if (var_trial) var_trial[q] = a[k - 1].trial;
            for (int64_t i = 0; i < t; ++i) idx[i] = a[k - 1 + i].trial;
            qsort(idx, (size_t)t, sizeof(int32_t), cmp_i32);
            double c = 0.0; /* plain double sum in ascending trial order */
            for (int64_t i = 0; i < t; ++i) c += 0.0 - V[idx[i]];
            if (cvar) cvar[q] = c / (double)t;
            if (breach_idx) memcpy(breach_idx + q * t, idx, (size_t)t * sizeof(int32_t));
        }
        free(V); free(a); free(idx);
    }
    return t;
}

/* ------------------------------------------------------------ Philox4x32-10
 * Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11).  Checked
 * in tests against the Random123 known-answer vectors.
 */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/*
 * synthetic matrix element (row, col) of tensor `tensor_id`:
 * twelve 32-bit Philox outputs (three blocks, counter = {row_lo, row_hi, col,
 * 4*tensor_id + j}) are summed exactly in 64-bit integers; x = (sum - 6*2^32)*2^-32
 * is an Irwin-Hall(12) variate with mean 0, variance 1 on [-6, 6); the result is
 * fma(x, scale, shift).  Every step is exact or a single IEEE rounding, so the CUDA
 * generator reproduces it bit for bit.
 */
void orc_gen_normal(double *out, int64_t rows, int32_t cols, uint64_t seed, uint32_t tensor_id,
                    double scale, double shift, int64_t row0, int nthreads)
{
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
orc_use_threads(nthreads);
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t r = 0; r < rows; ++r) {
        const uint64_t gr = (uint64_t)(row0 + r);
        for (int32_t c = 0; c < cols; ++c) {
            uint64_t sum = 0;
            for (uint32_t j = 0; j < 3; ++j) {
                const uint32_t ctr[4] = { (uint32_t)gr, (uint32_t)(gr >> 32), (uint32_t)c, 4u * tensor_id + j };
                uint32_t o[4];
                orc_philox4x32_10(ctr, key, o);
                sum += (uint64_t)o[0] + o[1] + o[2] + o[3];
            }
            const double x = (double)((int64_t)sum - (int64_t)(6ull << 32)) * 0x1p-32;
            out[r * cols + c] = fma(x, scale, shift);
        }
    }
}

/* trial n of list `list` breaches iff philox({n>>2, list_lo, list_hi, 0xB4EAC400}, seed)[n&3] < rate*2^32 */
int64_t orc_gen_breach_list(int32_t *out, int64_t list, int32_t N, double rate, uint64_t seed)
{
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    const uint32_t thresh = (uint32_t)(rate * 4294967296.0);
    int64_t cnt = 0;
    for (int32_t n0 = 0; n0 < N; n0 += 4) {
        const uint32_t ctr[4] = { (uint32_t)(n0 >> 2), (uint32_t)list, (uint32_t)((uint64_t)list >> 32), 0xB4EAC400u };
        uint32_t o[4];
        orc_philox4x32_10(ctr, key, o);
        for (int j = 0; j < 4 && n0 + j < N; ++j)
            if (o[j] < thresh) out[cnt++] = n0 + j;
    }
    return cnt;
}


/*
 * Ingest ([builder-defined]: the reference holds (instrument, weight) positions per portfolio,
 * LoadPortfolios.java:352-379, and no instrument -> factor model).  Exposure of portfolio p to factor f:
 *     acc = +0.0; for every position j of p in FILE ORDER: acc = fma(weight_j, B[instrument_j][f], acc)
 * with B the n_inst x F loading table.  Duplicate instruments simply contribute twice.  Returns -1 if an instrument
 * id is outside [0, n_inst).
 */
int orc_build_exposures(const int32_t *inst, const double *w, const int64_t *off, int64_t P, const double *B, int64_t n_inst,
                        int32_t F, double *E)
{
    for (int64_t p = 0; p < P; ++p)
        for (int64_t j = off[p]; j < off[p + 1]; ++j)
            if (inst[j] < 0 || inst[j] >= n_inst) return -1;
    for (int64_t p = 0; p < P; ++p)
        for (int32_t f = 0; f < F; ++f) {
            double acc = 0.0;
            for (int64_t j = off[p]; j < off[p + 1]; ++j) acc = fma(w[j], B[(int64_t)inst[j] * F + f], acc);
            E[p * F + f] = acc;
        }
    return 0;
}
