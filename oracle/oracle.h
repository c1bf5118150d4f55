/*
 * oracle.h -- CPU oracle for the monte-carlo-portfolios hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it, and only as the checker / CPU baseline.
 *
 * Two halves:
 *   codec_oracle.c -- restatement of the vendored JavaFastPFOR 0.1.8-SNAPSHOT
 *                     codecs that the reference's loader calls
 *                     (maprdb-portfolio/.../LoadPortfolios.java:608-662,725-726)
 *                     and that north_star names (FastPFOR.java, Delta.java, ...).
 *                     PINNED by: the stored-document known-answer vector
 *                     setup.txt:47-48 <-> portfolios.json:1, bit-packing golden
 *                     vectors obtained by interpreting the reference's own
 *                     BitPacking.java / IntegratedBitPacking.java method bodies
 *                     (tests/golden/make_bitpacking_golden.py), and the
 *                     reference's JUnit round-trip matrix.
 *   risk_oracle.c  -- valuation / moments / VaR / CVaR / breach list.  The
 *                     reference contains NO code for these stages,
 *                     so this half follows DESIGN.md's own specification:
 *                     PARITY UNPINNED by the reference.
 */
#ifndef MCP_ORACLE_H
#define MCP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* codec ids: identical to include/mcp.h */
enum {
    ORC_CODEC_BP32_VB        = 0, /* Composition(BinaryPacking, VariableByte)            LoadPortfolios.java:725 */
    ORC_CODEC_FASTPFOR256_VB = 1, /* Composition(FastPFOR, VariableByte)                 FastPFOR.java:22         */
    ORC_CODEC_FASTPFOR128_VB = 2, /* Composition(FastPFOR128, VariableByte)              FastPFOR128.java:33      */
    ORC_CODEC_IBP32_IVB      = 3, /* IntegratedComposition(IntegratedBinaryPacking, IntegratedVariableByte) */
    ORC_CODEC_SK_IBP32_IVB   = 4, /* IntegratedIntCompressor framing  IntegratedIntCompressor.java:38-62 */
    ORC_CODEC_VB             = 5, /* bare VariableByte                                   VariableByte.java:32 */
    ORC_CODEC_SK_BP32_VB     = 6, /* IntCompressor framing             IntCompressor.java:38-61 */
    ORC_CODEC_SK_FASTPFOR256_VB = 7, /* IntCompressor(SkippableComposition(FastPFOR, VariableByte)): n, then the headless stream
                                         IntCompressor.java:38-61, SkippableComposition.java:37-54, FastPFOR.java:92-105 */
    ORC_CODEC_SK_FASTPFOR128_VB = 8, /* the same over FastPFOR128 (FastPFOR128.java:33) */
    ORC_CODEC_XOR128_IVB     = 9, /* IntegratedComposition(XorBinaryPacking, IntegratedVariableByte)   differential/XorBinaryPacking.java:26-139 */
    ORC_CODEC_OPTPFD_VB      = 10, /* Composition(OptPFD, VariableByte)                   OptPFD.java:35-200, S16.java */
    ORC_CODEC_COUNT          = 11,
    ORC_CODEC_MASK           = 0xff,
    ORC_FLAG_DELTA           = 0x100 /* Delta.delta(int[]) before compress / inverseDelta after (Delta.java:24-28,84-88) */
};

/* ---- primitives (exported so the tests can pin them one by one) ---- */
int  orc_bits(uint32_t v);                                                 /* Util.java:101-103 */
int  orc_maxbits(const uint32_t *in, int n);                               /* Util.java:28-33   */
int  orc_maxdiffbits(uint32_t init, const uint32_t *in, int n);            /* Util.java:85-92   */
void orc_fastpack(const uint32_t *in, uint32_t *out, int b);               /* BitPacking.java:42-148 (masked)   */
void orc_fastpackwithoutmask(const uint32_t *in, uint32_t *out, int b);    /* BitPacking.java:1706-1812         */
void orc_fastunpack(const uint32_t *in, uint32_t *out, int b);             /* BitPacking.java:3021-3127         */
void orc_integratedpack(uint32_t init, const uint32_t *in, uint32_t *out, int b);   /* IntegratedBitPacking.java:39-145    */
void orc_integratedunpack(uint32_t init, const uint32_t *in, uint32_t *out, int b); /* IntegratedBitPacking.java:1815-1921 */
void orc_delta(uint32_t *data, size_t n);                                  /* Delta.java:24-28  */
void orc_inverse_delta(uint32_t *data, size_t n);                          /* Delta.java:84-88  */

/*
 * Whole-list compress with a FRESH codec instance:  codec.compress(in, 0, n, out, 0).
 * `out` must hold orc_max_compressed_words(n) words.  Returns the number of 32-bit
 * words written (Java's outpos), or -1 on a bad codec id.
 * With ORC_FLAG_DELTA the input is copied and Delta.delta() applied first.
 */
int64_t orc_compress(int codec, const int32_t *in, int64_t n, int32_t *out);

/*
 * Whole-list uncompress:  codec.uncompress(in, 0, inlen, out, 0).
 * `n_expected` is only used by the SK_* codecs' sanity check (they carry n in
 * word 0) and to size nothing; pass -1 if unknown.  `out_cap` bounds writes.
 * Returns the number of ints produced, or -1 on error/overrun.
 */
int64_t orc_uncompress(int codec, const int32_t *in, int64_t inlen, int32_t *out, int64_t out_cap);

int64_t orc_max_compressed_words(int64_t n);

/* Multi-call FastPFOR instance (stale exception-buffer semantics, FastPFOR.java:52-59,143). */
typedef struct orc_fastpfor orc_fastpfor;
orc_fastpfor *orc_fastpfor_new(int block_size);
void          orc_fastpfor_free(orc_fastpfor *);
/* FastPFOR.compress(in,0,n,out,0) on this instance; returns words written */
int64_t       orc_fastpfor_compress(orc_fastpfor *, const int32_t *in, int64_t n, int32_t *out);
int64_t       orc_fastpfor_uncompress(orc_fastpfor *, const int32_t *in, int64_t inlen, int32_t *out, int64_t *consumed);

/*
 * App framing (LoadPortfolios.java:608-623, BitManipulationHelper.java:29-66):
 * compress, then big-endian bytes of outpos+1 words (one trailing zero int).
 * Returns byte length.  bytes must hold 4*(orc_max_compressed_words(n)+1).
 */
int64_t orc_get_compressed_instruments(int codec, const int32_t *in, int64_t n, uint8_t *bytes);
/* LoadPortfolios.java:625-662: bytesToInts (big-endian) then uncompress with inlength = all words */
int64_t orc_get_decompressed_instruments(int codec, const uint8_t *bytes, int64_t nbytes, int32_t *out, int64_t out_cap);
/* BitManipulationHelper.java:17-27 : raw big-endian ints ("none" fallback) */
void    orc_integers_to_bytes(const int32_t *in, int64_t n, uint8_t *bytes);

/* Batched, OpenMP over lists (CPU baseline).  CSR in, CSR out.
 * framing: 0 = native-endian words, 1 = app framing (big-endian + trailing zero int).
 * out_off[i] = i * stride_bytes (caller-chosen slots); out_len[i] = bytes written. */
int orc_batch_compress(int codec, int framing, const int32_t *vals, const int64_t *list_off, int64_t nlists,
                       uint8_t *out, int64_t stride_bytes, int64_t *out_len, int nthreads);
int orc_batch_uncompress(int codec, int framing, const uint8_t *in, int64_t stride_bytes, const int64_t *in_len,
                         int64_t nlists, int32_t *out, const int64_t *out_off, int nthreads);

/* ---------------- weights column (weights_oracle.c) ----------------
 * Raw Snappy format: the decoder takes any valid stream (snappy-java's included); the encoder is this repository's own
 * greedy matcher -- valid Snappy, not snappy-java's byte choices (un-vendored dependency: parity unpinned). */
int64_t orc_snappy_compress(const uint8_t *in, int64_t len, uint8_t *out);               /* out == NULL: size only */
int64_t orc_snappy_uncompress(const uint8_t *in, int64_t inlen, uint8_t *out, int64_t out_cap);
int64_t orc_get_compressed_weights(const double *w, int64_t n, uint8_t *out);            /* LoadPortfolios.java:563-595 */
int64_t orc_get_decompressed_weights(const uint8_t *in, int64_t inlen, double *out, int64_t out_cap); /* :480,:580 */
void    orc_doubles_to_bytes(const double *w, int64_t n, uint8_t *bytes);                /* BitManipulationHelper.java:193-208 */

/* ---------------- risk half (builder-defined spec) ---------------- */

/* V[p*N+n] = fma-chain over k ascending of E[p*F+k]*S[n*F+k], acc starts at +0.0 */
int orc_build_exposures(const int32_t *inst, const double *w, const int64_t *off, int64_t P, const double *B, int64_t n_inst,
                        int32_t F, double *E);
void orc_value(const double *E, const double *S, int64_t P, int64_t N, int32_t F, double *V, int nthreads);

/*
 * Per-portfolio statistics and breach list for portfolios [p0, p1).
 *   loss L = 0.0 - V;  k = ceil(alpha*N) clamped to [1,N];  t = N-k+1.
 *   var_loss/var_trial = k-th smallest (L, trial) in lexicographic order.
 *   breach list = the t trials with (L,trial) >= (var_loss,var_trial), ascending trial.
 *   cvar = sum of breach losses in ascending trial order / t.
 *   mean = sum(V)/N, variance = sum((V-mean)^2)/N   (long double accumulation).
 * breach_idx is [ (p1-p0) * t ] row-major.  Returns t.
 */
int64_t orc_risk(const double *E, const double *S, int64_t P, int64_t N, int32_t F, double alpha,
                 int64_t p0, int64_t p1,
                 double *mean, double *variance, double *var_loss, int32_t *var_trial, double *cvar,
                 int32_t *breach_idx, int nthreads);

int64_t orc_tail_size(int64_t N, double alpha);

/* risk_fast.c: the same results computed fast (SIMD across trials, portfolio-blocked, threshold-prefiltered select) --
 * the CPU BASELINE bench.py times.  V, var_loss, var_trial, the breach lists and cvar are bit-identical to orc_value /
 * orc_risk (asserted in tests/test_oracle_risk.py); mean / variance agree to 1e-9 (double two-pass instead of long double). */
void orc_value_fast(const double *E, const double *S, int64_t P, int64_t N, int32_t F, double *V, int nthreads);
int64_t orc_risk_fast(const double *E, const double *S, int64_t P, int64_t N, int32_t F, double alpha,
                      int64_t p0, int64_t p1,
                      double *mean, double *variance, double *var_loss, int32_t *var_trial, double *cvar,
                      int32_t *breach_idx, int nthreads);
const char *orc_fast_isa(void); /* "avx512f" or "avx2+fma": the body picked on this host */

/* Philox4x32-10 synthetic generators (shared definition with the CUDA side) */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_gen_normal(double *out, int64_t rows, int32_t cols, uint64_t seed, uint32_t tensor_id,
                    double scale, double shift, int64_t row0, int nthreads);
/* ascending indices of a Bernoulli(rate) mask over N trials for lists [l0,l1): returns count per list */
int64_t orc_gen_breach_list(int32_t *out, int64_t list, int32_t N, double rate, uint64_t seed);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
