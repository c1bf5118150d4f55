/*
 * mcp.h -- C ABI of libmcp.so, the B200-native replacement for the one hot path of
 * mapr-demos/monte-carlo-portfolios:
 *
 *   portfolio x scenario valuation -> per-portfolio loss/tail statistics ->
 *   sorted breach-trial index lists -> JavaFastPFOR-format encoding (and decoding)
 *
 * The reference is pure Java with NO native seam (no JNI, no `native` methods).  The
 * Java-side interfaces this ABI sits behind are
 *
 *   IntegerCODEC.compress / uncompress
 *       JavaFastPFOR/src/main/java/me/lemire/integercompression/IntegerCODEC.java:36-58
 *   SkippableIntegerCODEC.headlessCompress / headlessUncompress
 *       .../SkippableIntegerCODEC.java:43-66
 *   LoadPortfolios.getCompressedInstruments / getDecompressedInstruments
 *       maprdb-portfolio/src/main/java/com/mapr/db/data/LoadPortfolios.java:608-623, 625-662
 *   the codec instance chosen at LoadPortfolios.java:725-726
 *
 * INTEGRATION.md shows the JNI / Panama-FFM stubs a maintainer would add.  Everything
 * here is plain pointers and sizes: JNI `GetPrimitiveArrayCritical` / FFM
 * `MemorySegment` friendly, no callbacks, no ownership transfer, never throws.
 *
 * There is NO CPU fallback: every entry point that computes runs sm_100a kernels and
 * fails with MCP_E_CUDA if no usable GPU is present.
 *
 * Threading: an mcp_ctx is bound to one GPU and one CUDA stream and must be used by
 * one thread at a time (the same rule the reference states for FastPFOR instances,
 * FastPFOR.java:36-37).  Use one ctx per thread / per GPU.
 */
#ifndef MCP_H
#define MCP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCP_VERSION 100 /* 0.1.0 */

/* ------------------------------------------------------------------ errors */
enum {
    MCP_OK            = 0,
    MCP_E_ARG         = -1, /* null pointer, negative size, unknown codec id ...            */
    MCP_E_CAP         = -2, /* caller's output buffer too small (Java: ArrayIndexOutOfBounds) */
    MCP_E_CUDA        = -3, /* CUDA runtime error / no device                               */
    MCP_E_NOMEM       = -4, /* device or host allocation failed                             */
    MCP_E_FORMAT      = -5, /* malformed compressed stream                                  */
    MCP_E_STATE       = -6, /* call sequence error (run before upload, ...)                 */
    MCP_E_UNSUPPORTED = -7, /* valid request this build does not implement  */
    MCP_E_NCCL        = -8
};
typedef struct mcp_ctx mcp_ctx;

const char *mcp_strerror(int err);
/* detail of the last error on this ctx (CUDA error string etc.); never NULL */
const char *mcp_last_error(const mcp_ctx *ctx);

/* ------------------------------------------------------------------ codecs */
/* codec_id = one of MCP_CODEC_* optionally OR-ed with MCP_FLAG_DELTA */
enum {
    MCP_CODEC_BP32_VB        = 0, /* new Composition(new BinaryPacking(), new VariableByte())   LoadPortfolios.java:725 ("bpacking+variablebyte") */
    MCP_CODEC_FASTPFOR256_VB = 1, /* new Composition(new FastPFOR(), new VariableByte())        FastPFOR.java:22         */
    MCP_CODEC_FASTPFOR128_VB = 2, /* new Composition(new FastPFOR128(), new VariableByte())     FastPFOR128.java:33      */
    MCP_CODEC_IBP32_IVB      = 3, /* new IntegratedComposition(new IntegratedBinaryPacking(), new IntegratedVariableByte()) */
    MCP_CODEC_SK_IBP32_IVB   = 4, /* new IntegratedIntCompressor()  (n, then headless stream)   IntegratedIntCompressor.java:38-62 */
    MCP_CODEC_VB             = 5, /* new VariableByte()                                         VariableByte.java:32
                                     (| MCP_FLAG_DELTA = new IntegratedVariableByte(): gaps from 0, the same bytes --
                                     differential/IntegratedVariableByte.java:35-76)                                     */
    MCP_CODEC_SK_BP32_VB     = 6, /* new IntCompressor()            (n, then headless stream)   IntCompressor.java:38-61 */
    MCP_CODEC_SK_FASTPFOR256_VB = 7, /* new IntCompressor(new SkippableComposition(new FastPFOR(), new VariableByte()))
                                        (n, then the headless stream: FastPFOR pages without their count word, VariableByte tail)
                                        IntCompressor.java:27-61, SkippableComposition.java:37-54, FastPFOR.java:92-105,222-231 */
    MCP_CODEC_SK_FASTPFOR128_VB = 8, /* the same over new FastPFOR128()                          FastPFOR128.java:33      */
    MCP_CODEC_XOR128_IVB     = 9, /* new IntegratedComposition(new XorBinaryPacking(), new IntegratedVariableByte())
                                     differential/XorBinaryPacking.java:26-139 (128-value blocks, XOR with the predecessor) */
    MCP_CODEC_OPTPFD_VB      = 10, /* new Composition(new OptPFD(), new VariableByte())        OptPFD.java:35-200, S16.java
                                      (128-value blocks, Simple16-coded exceptions; 9 and 10 run in the warp-per-list kernels) */
    MCP_CODEC_COUNT          = 11,
    MCP_CODEC_MASK           = 0xff,
    MCP_FLAG_DELTA           = 0x100 /* Delta.delta(int[]) before compress, inverse after uncompress (Delta.java:24-28,84-88) */
};

/* byte framing of a compressed list */
enum {
    MCP_FRAME_WORDS = 0, /* the codec's int[] output, native (little-endian) 32-bit words = Java int[0..outpos) */
    MCP_FRAME_APP   = 1  /* LoadPortfolios framing: big-endian bytes of outpos+1 ints (one trailing zero int)
                            LoadPortfolios.java:622, BitManipulationHelper.java:29-66 */
};

/* ------------------------------------------------------------------ context */
int  mcp_create(int device, mcp_ctx **out);
void mcp_destroy(mcp_ctx *ctx);
int  mcp_device_info(mcp_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *hbm_bytes);
int  mcp_sync(mcp_ctx *ctx);
/* Tuning / test switches.  Keys: "force_dense" (1 = value every trial into a dense tile and select once, the
 * path used for factor counts without a streaming kernel), "force_sort_finalize" (1 = breach list by bitonic
 * sort instead of the trial bitmap), "value_mma" (0 = mcp_value through the scalar-fma kernel instead of the FP64
 * tensor-core kernel), "fused_moments" (1 = mean/variance by a per-trial reduction fused into the valuation kernel
 * instead of by linearity, mean = E.mean(S), var = E^T Cov(S) E), "optimistic" (0 = always use the guaranteed
 * geometric trial schedule; 1 (default) = two-step schedule with an optimistic filter bound and an exact fallback
 * for the rows where the bet fails), "fast_codec" (0 = warp-per-list codec kernels only), "short_list_codec" (0 / 1 / 2:
 * thread-per-list codec kernels off / where they are the faster ones / for every codec), "flat_codec" (1 (default) = short
 * lists that are all VariableByte -- FastPFOR / FastPFOR128 below one block, VariableByte alone -- take the flat kernels: a
 * thread per 32 consecutive values, size pass + tile scan + emit pass; 0 = the thread-per-list kernels), "reg_select" / "small_select" /
 * "warp_select" (0 = dense-step selection staged in shared memory / 256-thread selection CTAs for small tails / no warp-per-row
 * selection of the sparse step for rows of <= 512 candidates), "dense_floor"
 * (smallest dense sample in trials, default 1024; process-wide), "dense_mult" (tail sizes in the dense sample of the
 * optimistic schedule, default 4; process-wide), "radix_finalize" (0 = tails of >= 1 024 trials over a trial range too long
 * for the bitmap are put in order by the bitonic sort instead of the counting kernel), "ts_exchange" (see
 * mcp_run_trial_sharded), "overlap_tiles" (> 1 = portfolio tiles alternate between
 * two streams), "item_trials" (trials per valuation work item), "upload_split" (0 = mcp_simulate uploads the scenario
 * tail in one part), "simulate_turns" (mcp_simulate calls in flight on one GPU -- one context and host thread each -- take
 * turns on the SMs: a call enqueues its uploads, takes the device's token, makes its kernels wait for the event behind the
 * previous call's last kernel, enqueues them, records its own event and hands the token on -- before its first host sync,
 * so the next call's kernels are queued when this call's finish; downloads happen outside; 1 / 0 = always / never,
 * -1 (default) = when a call uploads 48 MB or more: without turns the kernels of two long calls interleave, the calls
 * fall into step and copy together with the SMs idle), "ts_slack" (trial-sharded test hook).  Trial indices, breach lists and byte streams are identical
 * either way, moments agree to 1e-9; only the kernels differ. */
int  mcp_set_option(mcp_ctx *ctx, const char *key, int64_t value);

/* Upper bound on the encoded size in BYTES of one list of n ints under `framing`
 * (the analogue of the reference's `new int[4*n + 1024]`, LoadPortfolios.java:611). */
int64_t mcp_codec_max_bytes(int codec_id, int framing, int64_t n);

/* ----------------------------------------------- single-list drop-in calls
 * IntegerCODEC.compress(in, inpos, inlength, out, outpos) with a fresh codec:
 *   in/out are the Java arrays already offset by inpos/outpos; *out_words receives the
 *   number of ints written (what Java adds to outpos); out_cap is out.length - outpos.
 *   All inlength ints are consumed (Java adds inlength to inpos).
 * These run the same kernels as the batched calls with nlists = 1: correct, but a GPU
 * round trip per list -- use the batched calls for throughput.
 */
int mcp_codec_compress(mcp_ctx *ctx, int codec_id, const int32_t *in, int32_t inlength,
                       int32_t *out, int32_t out_cap, int32_t *out_words);
/* SkippableIntegerCODEC.headlessCompress(in, inpos, inlength, out, outpos) (SkippableIntegerCODEC.java:43-48) of the
 * composition behind one of the MCP_CODEC_SK_* ids (SkippableComposition(BinaryPacking | FastPFOR | FastPFOR128, VariableByte),
 * SkippableIntegratedComposition(IntegratedBinaryPacking, IntegratedVariableByte) with initvalue 0): the stream WITHOUT
 * any count word -- what IntCompressor.compress writes after compressed[0].  MCP_E_ARG for the other ids. */
int mcp_codec_headless_compress(mcp_ctx *ctx, int codec_id, const int32_t *in, int32_t inlength,
                                int32_t *out, int32_t out_cap, int32_t *out_words);
/* SkippableIntegerCODEC.headlessUncompress(in, inpos, inlength, out, outpos, num) (:58-66): exactly `num` ints are
 * produced (out must hold them); *in_words receives the ints consumed (what Java adds to inpos).  The consumed count is
 * obtained by re-encoding the decoded values, so it is exact for streams one of these encoders (or the reference) wrote;
 * for a foreign stream that decodes to the same values through a different layout it may differ. */
int mcp_codec_headless_uncompress(mcp_ctx *ctx, int codec_id, const int32_t *in, int32_t inlength,
                                  int32_t *out, int32_t num, int32_t *in_words);
/* IntegerCODEC.uncompress(in, inpos, inlength, out, outpos): *out_count = ints produced */
int mcp_codec_uncompress(mcp_ctx *ctx, int codec_id, const int32_t *in, int32_t inlength,
                         int32_t *out, int32_t out_cap, int32_t *out_count);
/* LoadPortfolios.getCompressedInstruments(ArrayList<Integer>) -> byte[] (MCP_FRAME_APP) */
int mcp_get_compressed_instruments(mcp_ctx *ctx, int codec_id, const int32_t *instruments, int32_t n,
                                   uint8_t *bytes, int64_t cap, int64_t *nbytes);
/* LoadPortfolios.getDecompressedInstruments(byte[], numberOfIntegers) -> int[] */
int mcp_get_decompressed_instruments(mcp_ctx *ctx, int codec_id, const uint8_t *bytes, int64_t nbytes,
                                     int32_t number_of_integers, int32_t *out);

/* ------------------------------------------------------- batched codec, host
 * CSR in, CSR out.  vals[list_off[i] .. list_off[i+1]) is list i.  On return
 * out_off[0..nlists] are byte offsets into out (out_off[0] = 0) and *out_len the total.
 * Copies host->device, runs the kernels, copies back.  MCP_E_CAP if cap is too small
 * (then *out_len holds the required size).
 */
int mcp_codec_encode(mcp_ctx *ctx, int codec_id, int framing,
                     const int32_t *vals, const int64_t *list_off, int64_t nlists,
                     int64_t *out_off, uint8_t *out, int64_t cap, int64_t *out_len);
/* in[in_off[i] .. in_off[i+1]) is the compressed list i; out_off gives where its
 * decoded ints go (out_off[i+1]-out_off[i] = the list's length, which the reference
 * stores beside the blob as "uncompressedIntegerSize", LoadPortfolios.java:600). */
int mcp_codec_decode(mcp_ctx *ctx, int codec_id, int framing,
                     const uint8_t *in, const int64_t *in_off, int64_t nlists,
                     int32_t *out, const int64_t *out_off);

/* diagnostics: encode calls served by the batched single-pass kernels since the context was created, and calls those
 * kernels handed back to the warp-per-list kernels because a batch did not fit their staging areas */
int mcp_codec_stats(mcp_ctx *ctx, int64_t *fast_encodes, int64_t *fast_fallbacks);
/* same for the thread-per-list kernels that take lists averaging <= 144 values (BASELINE config C4's ~100-int breach
 * lists): encode + decode calls they served, and calls they handed on to the batched kernels because a list did not
 * fit (a FastPFOR list of a whole block or more).  mcp_set_option("short_list_codec", v): 0 = off, 1 (default) = where
 * they are the faster kernels (the encoder for every codec, the decoder for FastPFOR / VariableByte), 2 = everywhere.
 * Bytes are identical whichever kernels run. */
int mcp_codec_stats_short(mcp_ctx *ctx, int64_t *short_calls, int64_t *short_fallbacks);

/* ----------------------------------------------------- batched codec, device
 * Same, but every pointer is a DEVICE pointer on ctx's GPU (for callers whose lists
 * are already in HBM).  total_out is a HOST pointer.  d_out_off is written.
 */
int mcp_codec_encode_dev(mcp_ctx *ctx, int codec_id, int framing,
                         const int32_t *d_vals, const int64_t *d_list_off, int64_t nlists, int64_t total_vals,
                         int64_t *d_out_off, uint8_t *d_out, int64_t cap, int64_t *total_out);
int mcp_codec_decode_dev(mcp_ctx *ctx, int codec_id, int framing,
                         const uint8_t *d_in, const int64_t *d_in_off, int64_t nlists,
                         int32_t *d_out, const int64_t *d_out_off);

/* ------------------------------------------------------------ weights column
 * The loader's second column: Snappy.compress(double[]) (LoadPortfolios.getCompressedWeights, LoadPortfolios.java:563-595;
 * Snappy.uncompressDoubleArray, :480,:580) = the RAW SNAPPY FORMAT over the doubles' little-endian bytes, stored if smaller
 * than 8 n bytes, else big-endian doubles under algo "none" (:270-278; the host-side mirror applies that rule).
 * snappy-java 1.1.10.4 is an un-vendored dependency of the reference (maprdb-portfolio/pom.xml:50-56): the decoder here
 * accepts every valid raw-Snappy stream (all four element kinds), so blobs the reference wrote decode; the encoder is a
 * greedy matcher of this library's own (oracle/weights_oracle.c states it) -- valid Snappy for any Snappy decoder, NOT
 * snappy-java's byte choices (parity unpinned).  CSR in / CSR out like the integer codecs; list i = weights[list_off[i] ..
 * list_off[i+1]).  MCP_E_CAP: *out_len holds the size needed.  MCP_E_FORMAT: malformed stream or a length that is not
 * 8 * (out_off[i+1] - out_off[i]). */
int64_t mcp_weights_max_bytes(int64_t n);
int mcp_weights_encode(mcp_ctx *ctx, const double *weights, const int64_t *list_off, int64_t nlists,
                       int64_t *out_off, uint8_t *out, int64_t cap, int64_t *out_len);
int mcp_weights_decode(mcp_ctx *ctx, const uint8_t *in, const int64_t *in_off, int64_t nlists,
                       double *out, const int64_t *out_off);
/* LoadPortfolios.getCompressedWeights(ArrayList<Double>) -> byte[] / Snappy.uncompressDoubleArray(byte[]) -> double[] */
int mcp_get_compressed_weights(mcp_ctx *ctx, const double *weights, int32_t n, uint8_t *bytes, int64_t cap, int64_t *nbytes);
int mcp_get_decompressed_weights(mcp_ctx *ctx, const uint8_t *bytes, int64_t nbytes, int32_t number_of_doubles, double *out);

/* -------------------------------------------------------- risk path: inputs
 * Exposure matrix E: row-major P x F doubles.  Scenario matrix S: row-major N x F
 * doubles, row n = trial n.  (The reference defines neither: portfolios.json holds
 * (instrument, weight) lists, macro.json holds 10 factor rows, and no Java code joins
 * them -- SURVEY.md section 0; DESIGN.md section 3 is the specification.)
 */
int mcp_upload_exposures(mcp_ctx *ctx, const double *E, int64_t P, int32_t F);
int mcp_upload_scenarios(mcp_ctx *ctx, const double *S, int64_t N, int32_t F);
/* Ingest: build E on the device from what the loader parses (LoadPortfolios.java:352-379: per portfolio the
 * `instruments` array of {instrument, weight}), CSR: positions pos_off[p] .. pos_off[p+1] belong to portfolio p.
 * E[p][f] = fma chain in FILE ORDER of weight * loadings[instrument][f], loadings = n_instruments x F row-major
 * (the instrument -> factor table; the reference ships none).  MCP_E_ARG on an id outside the table. */
int mcp_build_exposures(mcp_ctx *ctx, const int32_t *instruments, const double *weights, const int64_t *pos_off, int64_t P,
                        const double *loadings, int64_t n_instruments, int32_t F);
/* Counter-based on-device generation (Philox4x32-10, Irwin-Hall(12) variates;
 * value = fma(x, scale, shift); bit-identical to oracle/risk_oracle.c orc_gen_normal).
 * row0 = global index of the first row (for sharded generation). */
int mcp_generate_exposures(mcp_ctx *ctx, int64_t P, int32_t F, uint64_t seed, double scale, double shift, int64_t row0);
int mcp_generate_scenarios(mcp_ctx *ctx, int64_t N, int32_t F, uint64_t seed, double scale, double shift, int64_t row0);
int mcp_download_exposures(mcp_ctx *ctx, double *E);
int mcp_download_scenarios(mcp_ctx *ctx, double *S);

/* V[p0..p1) x [0..N) -> host, row-major (debug / parity: bit-identical to the oracle).
 * The tensor-core kernel (mma.sync.m8n8k4.f64) equals the k-ascending fma chain bit for bit on B200 -- a measured
 * property of this GPU (profiles/microbench/dmma_probe.cu), not an architectural guarantee: re-run `pytest -m gpu` on a
 * new GPU stepping or driver.  mcp_set_option("value_mma", 0) + ("force_dense", 1) is the scalar-fma escape hatch. */
int mcp_value(mcp_ctx *ctx, int64_t p0, int64_t p1, double *V);

/* tail size t = N - ceil(alpha*N - 1e-7) + 1 */
int64_t mcp_tail_size(int64_t N, double alpha);

/* ----------------------------------------------------------- risk path: run
 * Values every uploaded portfolio under every uploaded scenario, reduces, selects,
 * compacts and encodes ON THE DEVICE; results stay in HBM until fetched.
 *   trial_base: global index of scenario row 0 (trial-sharded runs), else 0.
 */
int mcp_run(mcp_ctx *ctx, double alpha, int codec_id, int framing);

/* sizes of the last run (P = portfolios whose results the context holds: all of them, or the rank's slice after a
 * trial-sharded run; N = trials the tail was taken over) */
int mcp_run_sizes(mcp_ctx *ctx, int64_t *P, int64_t *N, int64_t *tail, int64_t *enc_bytes);
/* diagnostics of the last run: *fallback_rows = portfolios whose optimistic filter bound was too tight and that
 * were redone with the guaranteed schedule (0 for exchangeable trial orders; results are exact either way) */
int mcp_run_stats(mcp_ctx *ctx, int64_t *fallback_rows);

/* Fetch results of the last mcp_run (any pointer may be NULL):
 *   mean, variance (population), var_loss, cvar : double[P];  var_trial : int32[P]
 *   breach_idx : int32[P * tail] ascending trial indices per portfolio
 *   enc_off : int64[P+1] byte offsets;  enc : uint8[enc_cap]
 */
int mcp_fetch(mcp_ctx *ctx, double *mean, double *variance, double *var_loss, int32_t *var_trial, double *cvar,
              int32_t *breach_idx, int64_t *enc_off, uint8_t *enc, int64_t enc_cap);

/* One-call host-buffer version (upload + run + fetch): the end-to-end path.  The call returns when the results are in the
 * caller's buffers.  For throughput keep two calls in flight: contexts are independent, so two contexts on one GPU driven
 * by two host threads (one context per thread at a time) let the uploads and downloads of one step ride under the kernels
 * of the other -- measured on C2: 2.85 ms per step against 3.23 ms back to back (kernels alone: 2.81 ms).  The Python
 * mirror is mcp_b200.SimulatePipeline; the Java shim's SimulatePortfolios would hold two McpNative handles the same way. */
int mcp_simulate(mcp_ctx *ctx, const double *E, int64_t P, const double *S, int64_t N, int32_t F,
                 double alpha, int codec_id, int framing,
                 double *mean, double *variance, double *var_loss, int32_t *var_trial, double *cvar,
                 int32_t *breach_idx, int64_t *enc_off, uint8_t *enc, int64_t enc_cap, int64_t *enc_len);

/* ------------------------------------------------- trial-sharded mode (optional)
 * BASELINE config C5: every rank holds ALL exposure rows and a contiguous shard of the scenario rows (trials
 * [trial_base, trial_base + N_local) of n_global, balanced: N_local <= ceil(n_global / world)).  Per portfolio the
 * global tail is found from the ranks' local top-m candidates (m ~ tail/world + 6 sigma), exchanged ONCE: a rank merges its
 * own portfolio slice, so mcp_run_trial_sharded sends it only that slice of every other rank's candidate block (one group
 * of ncclSend / ncclRecv; option "ts_exchange" 0 = ncclAllGather of the whole blocks, world times the bytes; the phase
 * calls below leave the transport to the caller and work with whole gathered blocks or with the slices in place)
 * (plus one byte per portfolio saying whether that was provably enough); rows for which it was not (adversarial
 * trial orders) go through a second round with the full local tails, so results are exact for any input.  After the run, rank g holds the results of its portfolio slice
 * [P g / world, P (g+1) / world) (same split as portfolio sharding): mcp_run_sizes / mcp_fetch return that slice.
 *
 * (1) with NCCL inside the library: rank 0 calls mcp_comm_unique_id and hands the 128 bytes to the other ranks
 *     through whatever channel the host application has; every rank calls mcp_comm_init, then mcp_run_trial_sharded.
 *     NCCL is bound at run time (dlopen "libnccl.so.2"); MCP_E_NCCL if it is not installed.
 */
int mcp_comm_unique_id(uint8_t id[128]);
int mcp_comm_init(mcp_ctx *ctx, const uint8_t id[128], int rank, int world);
int mcp_comm_destroy(mcp_ctx *ctx);
int mcp_run_trial_sharded(mcp_ctx *ctx, double alpha, int codec_id, int framing, int64_t n_global, int64_t trial_base);
/* (2) with the caller's own transport (MPI, sockets, a JVM-side shuffle ...), phase by phase.  All block pointers are
 *     DEVICE pointers; "gathered" = the world blocks of one round concatenated in rank order.
 *       mcp_ts_local   : local candidates + moments -> this rank's block (owned by ctx, *block_bytes each)
 *       mcp_ts_merge   : merge THIS RANK'S portfolio slice from the gathered blocks and test it for exactness; returns
 *                        the slice's flag block (one byte per row, ceil(P / world) bytes, owned by ctx).  The gathered
 *                        candidate buffer must stay valid until mcp_ts_finish.
 *       mcp_ts_flags   : the ranks' flag blocks gathered in rank order -> *flagged = rows needing the second round (the
 *                        same on every rank)
 *       mcp_ts_local2 / mcp_ts_merge2 : second round, only if *flagged > 0
 *       mcp_ts_finish  : breach lists, CVaR, pooled moments and encoding of this rank's portfolio slice
 */
int mcp_ts_local(mcp_ctx *ctx, double alpha, int64_t n_global, int64_t trial_base, int world, void **d_block, int64_t *block_bytes);
int mcp_ts_merge(mcp_ctx *ctx, const void *d_gathered, int rank, void **d_flags, int64_t *flag_bytes);
int mcp_ts_flags(mcp_ctx *ctx, const void *d_gathered_flags, int64_t *flagged);
int mcp_ts_local2(mcp_ctx *ctx, void **d_block, int64_t *block_bytes);
int mcp_ts_merge2(mcp_ctx *ctx, const void *d_gathered2);
int mcp_ts_finish(mcp_ctx *ctx, int rank, int world, int codec_id, int framing);
/* candidates per rank and portfolio the first round exchanges (for sizing / reporting) */
int64_t mcp_ts_candidates(int64_t n_global, double alpha, int world);

/* ------------------------------------------------------------- measurement
 * Kernel timing with CUDA events on ctx's own stream.  Classes: */
enum {
    MCP_K_VALUE = 0,   /* valuation (+ fused moments / threshold filter) */
    MCP_K_SELECT = 1,  /* radix-select / prune / finalize               */
    MCP_K_ENCODE = 2,  /* codec encode (size + emit)                    */
    MCP_K_DECODE = 3,  /* codec decode                                  */
    MCP_K_SCAN = 4,    /* prefix sums / bookkeeping                     */
    MCP_K_GEN = 5,     /* generators, FP64 peak probe                   */
    MCP_K_MOMENTS = 6, /* mean / variance (on a side stream: overlaps the valuation) */
    MCP_K_COMM = 7,    /* NCCL exchange of the trial-sharded mode               */
    MCP_K_COUNT = 8
};
int mcp_profile_enable(mcp_ctx *ctx, int on);
int mcp_profile_reset(mcp_ctx *ctx);
/* per class: number of launches and summed device ms since the last reset (syncs the stream) */
int mcp_profile_get(mcp_ctx *ctx, int64_t launches[MCP_K_COUNT], double ms[MCP_K_COUNT]);
/* whole-region timing on ctx's stream */
int mcp_timer_start(mcp_ctx *ctx);
int mcp_timer_stop(mcp_ctx *ctx, double *ms);
/* dependent-free DFMA issue-rate probe: returns achieved FP64 TFLOP/s (2 flop per FMA) */
int mcp_fp64_peak(mcp_ctx *ctx, double *tflops);
/* the same for the FP64 tensor-core instruction the valuation kernel issues (mma.sync.m8n8k4.f64, SASS DMMA) */
int mcp_fp64_mma_peak(mcp_ctx *ctx, double *tflops);
/* write `bytes` of zeros to a scratch buffer (> L2) to flush the L2 between timed steps */
int mcp_flush_l2(mcp_ctx *ctx);

/* Synthetic breach lists on device (Bernoulli(rate) mask over n_trials per list; Philox;
 * bit-identical to oracle orc_gen_breach_list).  Returns device pointers owned by ctx
 * (valid until the next call / destroy).  first_list = global id of list 0. */
int mcp_generate_breach_lists(mcp_ctx *ctx, int64_t nlists, int32_t n_trials, double rate, uint64_t seed,
                              int64_t first_list, const int32_t **d_vals, const int64_t **d_list_off, int64_t *total_vals);

/* raw device memory helpers for callers without a CUDA runtime of their own */
int mcp_dev_alloc(mcp_ctx *ctx, size_t bytes, void **dptr);
int mcp_dev_free(mcp_ctx *ctx, void *dptr);
int mcp_memcpy_h2d(mcp_ctx *ctx, void *dptr, const void *hptr, size_t bytes);
int mcp_memcpy_d2h(mcp_ctx *ctx, void *hptr, const void *dptr, size_t bytes);
int mcp_host_alloc_pinned(size_t bytes, void **hptr);
int mcp_host_free_pinned(void *hptr);

#ifdef __cplusplus
}
#endif
#endif /* MCP_H */
