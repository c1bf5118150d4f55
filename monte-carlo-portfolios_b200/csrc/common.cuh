// common.cuh -- context, error plumbing and device-buffer helpers shared by the
// translation units of libmcp.so.  Plain CUDA runtime; no torch types anywhere.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mcp.h"

namespace mcp {

constexpr int kWarp = 32;

// A grow-only device buffer (cudaMalloc is slow; steps reuse their scratch).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Profile {
    bool on = false;
    struct Rec { int cls; cudaEvent_t a, b; };
    std::vector<Rec> recs;          // recorded, not yet resolved
    std::vector<cudaEvent_t> pool;  // free events
    int64_t launches[MCP_K_COUNT] = {0};
    double ms[MCP_K_COUNT] = {0};
};

} // namespace mcp

struct mcp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    std::string last_error;

    // ---- risk path state
    mcp::DevBuf dE, dS;            // exposures P x F, scenarios N x F (doubles)
    mcp::DevBuf dEpad, dSpad;      // zero-padded copies for factor counts without a tensor-core kernel
    mcp::DevBuf dSfrag;            // scenarios in tensor-core B-fragment order (built lazily from dS)
    bool sfragValid = false;
    int64_t P = 0, N = 0;
    int32_t F = 0;
    bool haveE = false, haveS = false;

    // results of the last run (device)
    mcp::DevBuf dMean, dVar, dVarLoss, dVarTrial, dCvar, dBreach, dEncOff, dEnc;
    int64_t runP = 0, runN = 0, runTail = 0, runEncBytes = 0;
    int64_t runRow0 = 0;           // first portfolio of the result arrays that the last run produced (trial-sharded: the rank's slice)
    bool haveRun = false;

    // scratch
    mcp::DevBuf dLoss;             // dense loss tile
    mcp::DevBuf dPart;             // moment partials
    mcp::DevBuf dSizes, dScanTmp;  // codec size/scan scratch
    mcp::DevBuf dIo0, dIo1, dIo2, dIo3; // host-API staging (vals, offsets, out, out offsets)
    mcp::DevBuf dErr;              // int flags (error, multi-page count, shadow ticket)
    mcp::DevBuf dShadow;           // FastPFOR stale-buffer emulation for multi-page lists
    int planMultiPage = 0;
    mcp::DevBuf dFlush;            // L2 flush buffer
    mcp::DevBuf dGenVals, dGenOff; // synthetic breach lists
    mcp::DevBuf dCandKey, dCandCnt, dTau, dKeptKey, dKeptIdx; // streaming tail filter
    mcp::DevBuf dKept2Key, dKept2Idx, dFinKey, dFinIdx;
    mcp::DevBuf dFail;             // [0] = number of rows that failed the optimistic bound, [4..] = their indices
    mcp::DevBuf dFbE, dFbOut, dFbKey, dFbIdx; // fallback problem (failed rows, guaranteed schedule)
    bool optimistic = true;        // two-step schedule with the optimistic filter bound (risk_kernels.cu)
    int64_t itemTrials = 2048;     // target trials per valuation CTA work item (amortises the exposure-tile prologue)
    bool failPending = false;      // the optimistic run's failed-row count is in flight to hPinned + 64
    int64_t runFallbackRows = 0;   // rows the last run had to redo with the guaranteed schedule
    // trial-sharded mode (risk_kernels.cu): state between the phase calls
    int64_t tsT = 0, tsM = 0, tsNGlobal = 0, tsTrialBase = 0, tsNFlag = 0, tsSlackOverride = -1;
    bool tsSliceExchange = true;   // trial-sharded: ranks exchange portfolio SLICES of their blocks (grouped send / recv), not whole blocks
    int tsWorld = 0, tsRank = -1;
    const unsigned char *tsGath = nullptr; // first-round gathered blocks (caller- or library-owned), needed until ts_finish
    std::vector<int32_t> tsFlagRows;
    mcp::DevBuf dTsSend, dTsRecv, dTsRecv2, dTsFlags, dTsFlagsAll;
    void *nccl = nullptr;          // NcclComm* (nccl_dyn.cu), null until mcp_comm_init
    cudaStream_t stream3 = nullptr; // second tile stream (risk_stream_core), `cur` = the stream the tile helpers launch on
    cudaStream_t cur = nullptr;
    cudaEvent_t evTileFork = nullptr, evTileJoin = nullptr, evSfrag = nullptr;
    int64_t overlapTiles = 1;      // >1: that many portfolio tiles alternate between two streams (selection of one tile under the valuation of the next); with 8-warp valuation CTAs: C2 2.79 -> 2.70 ms, C3 6.85 -> 6.70 ms per 125 000 rows (round 1, 16-warp CTAs: -2 % at best).  Off by default: the two-calls-in-flight pipeline of mcp_simulate gets the same overlap across steps (2.75 ms e2e), and per-kernel event times stay those of kernels running alone
    mcp::DevBuf dCandKey2, dCandCnt2;
    mcp::DevBuf dOvf;                  // [2][tile rows + 16] rows the warp-per-row selection handed back (count, then row indices)
    cudaStream_t stream2 = nullptr; // side stream: the moments (independent of the selection chain) overlap the valuation
    cudaEvent_t evFork = nullptr, evJoin = nullptr, evTail = nullptr, evTail1 = nullptr;
    bool sTailPending = false;     // pipelined upload: scenario rows >= sHeadRows are still arriving on stream2 (evTail)
    int64_t sHeadRows = 0;
    bool uploadSplit = true;       // mcp_simulate: scenario tail in two parts, sparse valuation in two launches
    int64_t sMidRows = 0;          // > 0: the tail arrives in two parts, rows [sHeadRows, sMidRows) first (evTail1), then the rest (evTail)
    void *earlyBreachDst = nullptr; // host buffer the raw breach lists are copied to under the encode (mcp_simulate)
    // host buffers the per-portfolio statistics are copied to under the encode (mcp_simulate); any may be null
    double *earlyMean = nullptr, *earlyVar = nullptr, *earlyVarLoss = nullptr, *earlyCvar = nullptr;
    int32_t *earlyVarTrial = nullptr;
    bool forceSortFinalize = false; // debug: breach list by bitonic sort instead of the trial bitmap
    int selCap = 2304;
    uint32_t mmaConfigured = 0;
    bool valueMma = true;          // mcp_value through the tensor-core kernel when F allows (else scalar fma)
    bool fusedMoments = false;     // true: per-trial moment reduction fused in the valuation kernel (scalar FP64)
    mcp::DevBuf dGram;
    int64_t fastEncodeCalls = 0, fastEncodeFallbacks = 0; // encodes served by the batched kernels / handed back to the warp-per-list ones
    bool smallSelect = true;       // 64-thread selection / compaction CTAs for small tails
    int selSlack = 0;              // extra entries of the sparse steps' staging area (experiment knob)
    bool radixFinalize = true;     // large tails over long trial ranges: counting finalize (k_finalize_radix) instead of the bitonic sort
    int simulateTurns = -1;        // mcp_simulate calls in flight on one GPU take turns on the SMs: 0 no, 1 yes, -1 by upload size
    cudaEvent_t evTurn = nullptr;  // ... "this call's kernels are done", what the next call's kernels wait for
    bool turnHeld = false;         // this context holds the device's turn token (between simulate_turn_begin and _end)
    bool warpSelect = true;        // small rows (dense samples of <= 1 024 values, <= 512 staged candidates): one warp per row
    bool regSelect = true;         // dense-step selection from registers when the row fits (risk_kernels.cu)
    bool fastCodec = true;         // batched single-pass kernels for the BinaryPacking family (codec_kernels.cu)
    int flatCodec = 1;             // 1 = flat (thread per 32 values) VariableByte kernels for short lists that are all VariableByte (default); 0 = thread-per-list kernels
    int shortListCodec = 1;        // thread-per-list kernels for lists averaging <= 144 values: 0 off, 1 where they win, 2 every codec
    int64_t shortListCalls = 0, shortListFallbacks = 0; // encode + decode calls they served / handed on to the batched kernels
    bool forceDense = false;       // debug: use the generic dense path even when F == 32
    void *hPinned = nullptr;       // small pinned staging (scalars)

    mcp::Profile prof;
    cudaEvent_t tA = nullptr, tB = nullptr;
};

namespace mcp {

inline int fail(mcp_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess)
{
    if (ctx) {
        char buf[512];
        if (e != cudaSuccess)
            snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
        else
            snprintf(buf, sizeof buf, "%s", what);
        ctx->last_error = buf;
    }
    return code;
}

#define MCP_CUDA(ctx, expr)                                                          \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            cudaGetLastError();                                                      \
            return ::mcp::fail((ctx), _e == cudaErrorMemoryAllocation ? MCP_E_NOMEM : MCP_E_CUDA, #expr, _e); \
        }                                                                            \
    } while (0)

#define MCP_TRY(expr)                 \
    do {                              \
        int _rc = (expr);             \
        if (_rc != MCP_OK) return _rc; \
    } while (0)

// RAII-ish scope that brackets kernel launches of one class with CUDA events on the
// ctx stream when profiling is enabled.
struct KScope {
    mcp_ctx *ctx;
    int cls;
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t st;
    KScope(mcp_ctx *c, int k, int nlaunch = 1, cudaStream_t s = nullptr) : ctx(c), cls(k), st(s ? s : (c->cur ? c->cur : c->stream))
    {
        ctx->prof.launches[cls] += nlaunch;
        if (!ctx->prof.on) return;
        auto get = [&]() {
            cudaEvent_t e;
            if (!ctx->prof.pool.empty()) { e = ctx->prof.pool.back(); ctx->prof.pool.pop_back(); }
            else cudaEventCreate(&e);
            return e;
        };
        a = get();
        b = get();
        cudaEventRecord(a, st);
    }
    ~KScope()
    {
        if (!a) return;
        cudaEventRecord(b, st);
        ctx->prof.recs.push_back({cls, a, b});
    }
};

inline cudaStream_t cur_stream(mcp_ctx *ctx) { return ctx->cur ? ctx->cur : ctx->stream; }

inline int check_launch(mcp_ctx *ctx, const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, MCP_E_CUDA, what, e);
    return MCP_OK;
}

// ---- launch-side entry points implemented in the .cu files ----

// codec_kernels.cu
int codec_encode_plan(mcp_ctx *ctx, int codec_id, int framing, const int32_t *d_vals, const int64_t *d_list_off,
                      int64_t uniform_len, int64_t nlists, int64_t *d_out_off, int64_t *total_out);
int codec_encode_emit(mcp_ctx *ctx, int codec_id, int framing, const int32_t *d_vals, const int64_t *d_list_off,
                      int64_t uniform_len, int64_t nlists, const int64_t *d_out_off, uint8_t *d_out);
int codec_decode_counts(mcp_ctx *ctx, int codec_id, int framing, const uint8_t *d_in, const int64_t *d_in_off,
                        int64_t nlists, int64_t *d_counts);
int codec_decode_device(mcp_ctx *ctx, int codec_id, int framing, const uint8_t *d_in, const int64_t *d_in_off,
                        int64_t nlists, int32_t *d_out, const int64_t *d_out_off);
bool codec_fast_applicable(int codec_id, int64_t total_vals, int64_t nlists, bool encode);
void simulate_turn_end(mcp_ctx *ctx); // (mcp_api.cu) called before the first host sync behind a call's last kernel
int64_t codec_fast_max_bytes(int codec_id, int64_t total_vals, int64_t nlists);
int codec_fast_encode(mcp_ctx *ctx, int codec_id, int framing, const int32_t *d_vals, const int64_t *d_list_off,
                      int64_t uniform_len, int64_t nlists, int64_t total_vals, int64_t *d_out_off, uint8_t *d_out, int64_t cap,
                      int64_t *total_out);
int codec_fast_decode(mcp_ctx *ctx, int codec_id, int framing, const uint8_t *d_in, const int64_t *d_in_off, int64_t nlists,
                      int64_t total_out_vals, int32_t *d_out, const int64_t *d_out_off);
int exclusive_scan_i64(mcp_ctx *ctx, const int64_t *d_in, int64_t n, int64_t *d_out /* n+1 */);
int64_t codec_max_words(int codec, int64_t n);

// weights_kernels.cu: the loader's weights column (raw Snappy format over the doubles' bytes)
int weights_encode_plan(mcp_ctx *ctx, const double *d_w, const int64_t *d_off, int64_t nlists, int64_t *d_out_off, int64_t *total);
int weights_encode_emit(mcp_ctx *ctx, const double *d_w, const int64_t *d_off, int64_t nlists, const int64_t *d_out_off, uint8_t *d_out);
int weights_decode_device(mcp_ctx *ctx, const uint8_t *d_in, const int64_t *d_in_off, int64_t nlists, double *d_out, const int64_t *d_out_off);

// risk_kernels.cu
int risk_run(mcp_ctx *ctx, double alpha, int codec_id, int framing);
int risk_value(mcp_ctx *ctx, int64_t p0, int64_t p1, double *d_V);
int gen_matrix(mcp_ctx *ctx, double *d_out, int64_t rows, int32_t cols, uint64_t seed, uint32_t tensor_id, double scale,
               double shift, int64_t row0);
int gen_breach_lists(mcp_ctx *ctx, int64_t nlists, int32_t n_trials, double rate, uint64_t seed, int64_t first_list,
                     int64_t *total);
int64_t risk_first_step_len(int64_t N, int64_t t);
void risk_set_dense_floor(int64_t v);
void risk_set_dense_mult(int64_t v);
int build_exposures(mcp_ctx *ctx, const int32_t *d_inst, const double *d_w, const int64_t *d_off, int64_t P, const double *d_B,
                    int64_t n_inst, int32_t F, double *d_E);
int fp64_peak(mcp_ctx *ctx, double *tflops);
int fp64_mma_peak(mcp_ctx *ctx, double *tflops);
int64_t ts_candidates_per_rank(int64_t n_global, int64_t t, int world);
int risk_ts_local(mcp_ctx *ctx, double alpha, int64_t n_global, int64_t trial_base, int world, void **d_block, int64_t *block_bytes);
int risk_ts_merge(mcp_ctx *ctx, const void *d_gathered, int rank, void **d_flags, int64_t *flag_bytes);
int risk_ts_flags(mcp_ctx *ctx, const void *d_gathered_flags, int64_t *nflag_out);
int risk_ts_local2(mcp_ctx *ctx, void **d_block, int64_t *block_bytes);
int risk_ts_merge2(mcp_ctx *ctx, const void *d_gathered2);
int risk_ts_finish(mcp_ctx *ctx, int rank, int world, int codec_id, int framing);

// nccl_dyn.cu: NCCL bound at run time (dlopen), so that libmcp.so loads on boxes without it
int nccl_unique_id(unsigned char id[128], std::string *err);
int nccl_comm_init(mcp_ctx *ctx, const unsigned char id[128], int rank, int world);
void nccl_comm_destroy(mcp_ctx *ctx);
int nccl_all_gather(mcp_ctx *ctx, const void *send, void *recv, size_t bytes);
bool nccl_has_send_recv();
// One region of the ranks' exchange blocks: `head` bytes at `off` that every peer gets as they are, or rows of `row_bytes`
// bytes from `off` on, of which peer d gets rows [p0[d], p0[d + 1]).
struct ExchRegion { size_t off, row_bytes, head; };
int nccl_slice_exchange(mcp_ctx *ctx, const unsigned char *send_block, unsigned char *recv_base, size_t block_bytes,
                        const ExchRegion *regions, int nregions, const int64_t *p0 /* [world + 1] */);
int risk_ts_exchange_regions(mcp_ctx *ctx, int world, ExchRegion regions[5], int64_t *p0 /* [world + 1] */);
int nccl_rank_world(mcp_ctx *ctx, int *rank, int *world);

} // namespace mcp
