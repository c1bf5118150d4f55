// mcp_api.cu -- the extern "C" surface declared in include/mcp.h.
// Host-side argument checking, host<->device staging, and nothing else: all compute is
// in codec_kernels.cu / risk_kernels.cu.  There is no CPU implementation of any stage.
#include <math.h>

#include <new>

#include "common.cuh"

#include <mutex>

// (see simulate_turn_begin)
struct TurnSlot {
    std::mutex m;
    cudaEvent_t last = nullptr; // recorded behind the last kernel of the latest call that took a turn on this device
};
static TurnSlot g_turn[64];

using namespace mcp;

namespace {

int set_dev(mcp_ctx *ctx)
{
    if (!ctx) return MCP_E_ARG;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail(ctx, MCP_E_CUDA, "cudaSetDevice", e);
    return MCP_OK;
}

int h2d(mcp_ctx *ctx, void *d, const void *h, size_t bytes)
{
    if (bytes == 0) return MCP_OK;
    MCP_CUDA(ctx, cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return MCP_OK;
}

int d2h(mcp_ctx *ctx, void *h, const void *d, size_t bytes)
{
    if (bytes == 0) return MCP_OK;
    MCP_CUDA(ctx, cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return MCP_OK;
}

int sync(mcp_ctx *ctx)
{
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MCP_OK;
}

bool codec_ok(int codec_id)
{
    const int c = codec_id & MCP_CODEC_MASK;
    return c >= 0 && c < MCP_CODEC_COUNT && !(codec_id & ~(MCP_CODEC_MASK | MCP_FLAG_DELTA));
}

void resolve_profile(mcp_ctx *ctx)
{
    for (auto &r : ctx->prof.recs) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) ctx->prof.ms[r.cls] += ms;
        ctx->prof.pool.push_back(r.a);
        ctx->prof.pool.push_back(r.b);
    }
    ctx->prof.recs.clear();
}

} // namespace

static int simulate_turn_begin(mcp_ctx *ctx)
{
    TurnSlot &slot = g_turn[ctx->device & 63];
    slot.m.lock();
    ctx->turnHeld = true;
    if (!ctx->evTurn) MCP_CUDA(ctx, cudaEventCreateWithFlags(&ctx->evTurn, cudaEventDisableTiming));
    if (slot.last && slot.last != ctx->evTurn) MCP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, slot.last, 0));
    return MCP_OK;
}

namespace mcp {
void simulate_turn_end(mcp_ctx *ctx)
{
    if (!ctx->turnHeld) return;
    TurnSlot &slot = g_turn[ctx->device & 63];
    if (ctx->evTurn && cudaEventRecord(ctx->evTurn, ctx->stream) == cudaSuccess) slot.last = ctx->evTurn;
    ctx->turnHeld = false;
    slot.m.unlock();
}
} // namespace mcp


extern "C" {

const char *mcp_strerror(int err)
{
    switch (err) {
    case MCP_OK: return "ok";
    case MCP_E_ARG: return "invalid argument";
    case MCP_E_CAP: return "output buffer too small";
    case MCP_E_CUDA: return "CUDA error";
    case MCP_E_NOMEM: return "out of memory";
    case MCP_E_FORMAT: return "malformed compressed stream";
    case MCP_E_STATE: return "call sequence error";
    case MCP_E_UNSUPPORTED: return "not supported by this build";
    case MCP_E_NCCL: return "NCCL error";
    }
    return "unknown error";
}

const char *mcp_last_error(const mcp_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int mcp_create(int device, mcp_ctx **out)
{
    if (!out) return MCP_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return MCP_E_CUDA; // no GPU: there is deliberately no CPU fallback
    }
    if (device < 0 || device >= ndev) return MCP_E_ARG;
    mcp_ctx *ctx = new (std::nothrow) mcp_ctx();
    if (!ctx) return MCP_E_NOMEM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream3, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evTileFork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evTileJoin, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evSfrag, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evJoin, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evTail, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evTail1, cudaEventDisableTiming) != cudaSuccess ||
        cudaMallocHost(&ctx->hPinned, 4096) != cudaSuccess || cudaEventCreate(&ctx->tA) != cudaSuccess ||
        cudaEventCreate(&ctx->tB) != cudaSuccess) {
        cudaGetLastError();
        mcp_destroy(ctx);
        return MCP_E_CUDA;
    }
    *out = ctx;
    return MCP_OK;
}

void mcp_destroy(mcp_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
    if (ctx->stream3) cudaStreamSynchronize(ctx->stream3);
    resolve_profile(ctx);
    nccl_comm_destroy(ctx);
    for (auto e : ctx->prof.pool) cudaEventDestroy(e);
    mcp::DevBuf *bufs[] = {&ctx->dE, &ctx->dS, &ctx->dMean, &ctx->dVar, &ctx->dVarLoss, &ctx->dVarTrial, &ctx->dCvar,
                           &ctx->dBreach, &ctx->dEncOff, &ctx->dEnc, &ctx->dLoss, &ctx->dPart, &ctx->dSizes,
                           &ctx->dScanTmp, &ctx->dIo0, &ctx->dIo1, &ctx->dIo2, &ctx->dIo3, &ctx->dErr, &ctx->dFlush,
                           &ctx->dGenVals, &ctx->dGenOff, &ctx->dCandKey, &ctx->dCandCnt, &ctx->dTau,
                           &ctx->dKeptKey, &ctx->dKeptIdx, &ctx->dShadow, &ctx->dKept2Key, &ctx->dKept2Idx,
                           &ctx->dFinKey, &ctx->dFinIdx, &ctx->dSfrag, &ctx->dGram, &ctx->dFail, &ctx->dFbE, &ctx->dFbOut,
                           &ctx->dFbKey, &ctx->dFbIdx, &ctx->dEpad, &ctx->dSpad, &ctx->dCandKey2, &ctx->dCandCnt2, &ctx->dTsSend, &ctx->dTsRecv, &ctx->dTsRecv2, &ctx->dTsFlags, &ctx->dTsFlagsAll};
    for (auto b : bufs) b->release();
    if (ctx->hPinned) cudaFreeHost(ctx->hPinned);
    if (ctx->tA) cudaEventDestroy(ctx->tA);
    if (ctx->tB) cudaEventDestroy(ctx->tB);
    if (ctx->evTurn) { // nobody may wait for an event that is about to go
        TurnSlot &slot = g_turn[ctx->device & 63];
        std::lock_guard<std::mutex> g(slot.m);
        if (slot.last == ctx->evTurn) slot.last = nullptr;
        cudaEventDestroy(ctx->evTurn);
    }
    if (ctx->evFork) cudaEventDestroy(ctx->evFork);
    if (ctx->evJoin) cudaEventDestroy(ctx->evJoin);
    if (ctx->evTail) cudaEventDestroy(ctx->evTail);
    if (ctx->evTail1) cudaEventDestroy(ctx->evTail1);
    if (ctx->evTileFork) cudaEventDestroy(ctx->evTileFork);
    if (ctx->evTileJoin) cudaEventDestroy(ctx->evTileJoin);
    if (ctx->evSfrag) cudaEventDestroy(ctx->evSfrag);
    if (ctx->stream3) cudaStreamDestroy(ctx->stream3);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int mcp_device_info(mcp_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *hbm_bytes)
{
    if (!ctx) return MCP_E_ARG;
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (cc_major) *cc_major = ctx->prop.major;
    if (cc_minor) *cc_minor = ctx->prop.minor;
    if (hbm_bytes) *hbm_bytes = ctx->prop.totalGlobalMem;
    return MCP_OK;
}

int mcp_set_option(mcp_ctx *ctx, const char *key, int64_t value)
{
    if (!ctx || !key) return MCP_E_ARG;
    if (!strcmp(key, "force_dense")) ctx->forceDense = value != 0;
    else if (!strcmp(key, "force_sort_finalize")) ctx->forceSortFinalize = value != 0;
    else if (!strcmp(key, "value_mma")) ctx->valueMma = value != 0;
    else if (!strcmp(key, "fused_moments")) ctx->fusedMoments = value != 0;
    else if (!strcmp(key, "optimistic")) ctx->optimistic = value != 0;
    else if (!strcmp(key, "fast_codec")) ctx->fastCodec = value != 0;
    else if (!strcmp(key, "flat_codec")) ctx->flatCodec = value != 0;
    else if (!strcmp(key, "short_list_codec")) ctx->shortListCodec = (int)(value < 0 ? 0 : value > 2 ? 2 : value);
    else if (!strcmp(key, "reg_select")) ctx->regSelect = value != 0;
    else if (!strcmp(key, "warp_select")) ctx->warpSelect = value != 0;
    else if (!strcmp(key, "sel_slack")) ctx->selSlack = (int)value;
    else if (!strcmp(key, "radix_finalize")) ctx->radixFinalize = value != 0;
    else if (!strcmp(key, "simulate_turns")) ctx->simulateTurns = (int)(value < 0 ? -1 : value > 0 ? 1 : 0);
    else if (!strcmp(key, "upload_split")) ctx->uploadSplit = value != 0;
    else if (!strcmp(key, "small_select")) ctx->smallSelect = value != 0;
    else if (!strcmp(key, "dense_floor")) risk_set_dense_floor(value);
    else if (!strcmp(key, "dense_mult")) risk_set_dense_mult(value);
    else if (!strcmp(key, "overlap_tiles")) ctx->overlapTiles = value < 1 ? 1 : value;
    else if (!strcmp(key, "ts_exchange")) ctx->tsSliceExchange = value != 0;
    else if (!strcmp(key, "ts_slack")) ctx->tsSlackOverride = value; // trial-sharded test hook: m = ceil(t / world) + value (-1 = automatic)
    else if (!strcmp(key, "item_trials")) ctx->itemTrials = value < 64 ? 64 : value;
    else return fail(ctx, MCP_E_ARG, "set_option: unknown key");
    return MCP_OK;
}

int mcp_sync(mcp_ctx *ctx)
{
    MCP_TRY(set_dev(ctx));
    return sync(ctx);
}

int64_t mcp_codec_max_bytes(int codec_id, int framing, int64_t n)
{
    if (!codec_ok(codec_id) || n < 0) return -1;
    return 4 * (codec_max_words(codec_id & MCP_CODEC_MASK, n) + (framing == MCP_FRAME_APP ? 1 : 0));
}

int64_t mcp_tail_size(int64_t N, double alpha)
{
    if (N <= 0) return 0;
    int64_t k = (int64_t)ceil(alpha * (double)N - 1e-7);
    if (k < 1) k = 1;
    if (k > N) k = N;
    return N - k + 1;
}

// ------------------------------------------------------------ batched codec
int mcp_codec_encode_dev(mcp_ctx *ctx, int codec_id, int framing, const int32_t *d_vals, const int64_t *d_list_off,
                         int64_t nlists, int64_t total_vals, int64_t *d_out_off, uint8_t *d_out, int64_t cap,
                         int64_t *total_out)
{
    MCP_TRY(set_dev(ctx));
    if (nlists < 0 || total_vals < 0 || !d_out_off || (nlists > 0 && !d_list_off)) return fail(ctx, MCP_E_ARG, "encode_dev: bad argument");
    if (!codec_ok(codec_id)) return fail(ctx, MCP_E_ARG, "unknown codec id");
    if (framing != MCP_FRAME_WORDS && framing != MCP_FRAME_APP) return fail(ctx, MCP_E_ARG, "unknown framing");
    int64_t total = 0;
    if (ctx->fastCodec && codec_fast_applicable(codec_id, total_vals, nlists, true)) {
        const int rc = codec_fast_encode(ctx, codec_id, framing, d_vals, d_list_off, 0, nlists, total_vals, d_out_off, d_out, cap, &total);
        if (rc != MCP_E_UNSUPPORTED) {
            if (total_out) *total_out = total;
            return rc;
        }
    }
    MCP_TRY(codec_encode_plan(ctx, codec_id, framing, d_vals, d_list_off, 0, nlists, d_out_off, &total));
    if (total_out) *total_out = total;
    if (total > cap || (total > 0 && !d_out)) return fail(ctx, MCP_E_CAP, "encode: output buffer too small");
    return codec_encode_emit(ctx, codec_id, framing, d_vals, d_list_off, 0, nlists, d_out_off, d_out);
}

int mcp_codec_decode_dev(mcp_ctx *ctx, int codec_id, int framing, const uint8_t *d_in, const int64_t *d_in_off,
                         int64_t nlists, int32_t *d_out, const int64_t *d_out_off)
{
    MCP_TRY(set_dev(ctx));
    if (nlists < 0 || (nlists > 0 && (!d_in_off || !d_out_off))) return fail(ctx, MCP_E_ARG, "decode_dev: bad argument");
    if (!codec_ok(codec_id)) return fail(ctx, MCP_E_ARG, "unknown codec id");
    if (framing != MCP_FRAME_WORDS && framing != MCP_FRAME_APP) return fail(ctx, MCP_E_ARG, "unknown framing");
    if (ctx->fastCodec && nlists > 0 && codec_fast_applicable(codec_id, 0, nlists, false)) {
        // the batch size follows the average list length: two offsets from the device
        MCP_CUDA(ctx, cudaMemcpyAsync(ctx->hPinned, d_out_off, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        MCP_CUDA(ctx, cudaMemcpyAsync(reinterpret_cast<char *>(ctx->hPinned) + 8, d_out_off + nlists, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const int64_t tv = reinterpret_cast<int64_t *>(ctx->hPinned)[1] - reinterpret_cast<int64_t *>(ctx->hPinned)[0];
        if (tv >= 0 && codec_fast_applicable(codec_id, tv, nlists, false)) {
            const int rc = codec_fast_decode(ctx, codec_id, framing, d_in, d_in_off, nlists, tv, d_out, d_out_off);
            if (rc != MCP_E_UNSUPPORTED) return rc;
        }
    }
    return codec_decode_device(ctx, codec_id, framing, d_in, d_in_off, nlists, d_out, d_out_off);
}

int mcp_codec_encode(mcp_ctx *ctx, int codec_id, int framing, const int32_t *vals, const int64_t *list_off, int64_t nlists,
                     int64_t *out_off, uint8_t *out, int64_t cap, int64_t *out_len)
{
    MCP_TRY(set_dev(ctx));
    if (nlists < 0 || !list_off || !out_off || cap < 0) return fail(ctx, MCP_E_ARG, "encode: bad argument");
    if (!codec_ok(codec_id)) return fail(ctx, MCP_E_ARG, "unknown codec id");
    const int64_t total_vals = list_off[nlists] - list_off[0];
    if (total_vals < 0 || (total_vals > 0 && !vals)) return fail(ctx, MCP_E_ARG, "encode: bad list offsets");
    for (int64_t i = 0; i < nlists; ++i)
        if (list_off[i + 1] < list_off[i]) return fail(ctx, MCP_E_ARG, "encode: list offsets not monotone");
    MCP_CUDA(ctx, ctx->dIo0.reserve((size_t)(total_vals + 1) * sizeof(int32_t)));
    MCP_CUDA(ctx, ctx->dIo1.reserve((size_t)(nlists + 1) * sizeof(int64_t)));
    MCP_CUDA(ctx, ctx->dIo3.reserve((size_t)(nlists + 1) * sizeof(int64_t)));
    MCP_TRY(h2d(ctx, ctx->dIo0.p, vals + list_off[0], (size_t)total_vals * sizeof(int32_t)));
    // offsets are rebased to the first list on the device side
    if (list_off[0] == 0) {
        MCP_TRY(h2d(ctx, ctx->dIo1.p, list_off, (size_t)(nlists + 1) * sizeof(int64_t)));
    } else {
        std::vector<int64_t> rb((size_t)nlists + 1);
        for (int64_t i = 0; i <= nlists; ++i) rb[i] = list_off[i] - list_off[0];
        MCP_TRY(h2d(ctx, ctx->dIo1.p, rb.data(), rb.size() * sizeof(int64_t)));
        MCP_TRY(sync(ctx));
    }
    int64_t total = 0;
    bool fast = ctx->fastCodec && (framing == MCP_FRAME_WORDS || framing == MCP_FRAME_APP) && codec_fast_applicable(codec_id, total_vals, nlists, true);
    if (fast) {
        // single pass: encode into a worst-case device buffer, then copy back what was produced
        const int64_t bound = codec_fast_max_bytes(codec_id, total_vals, nlists);
        MCP_CUDA(ctx, ctx->dIo2.reserve((size_t)bound + 16));
        const bool sizing = !out || cap <= 0; // size query: nothing is written
        const int rc = codec_fast_encode(ctx, codec_id, framing, ctx->dIo0.as<int32_t>(), ctx->dIo1.as<int64_t>(), 0, nlists, total_vals,
                                         ctx->dIo3.as<int64_t>(), sizing ? nullptr : ctx->dIo2.as<uint8_t>(), sizing ? 0 : (cap < bound ? cap : bound), &total);
        if (rc == MCP_E_UNSUPPORTED) fast = false; // a batch did not fit the FastPFOR kernel: warp-per-list kernels below
        else {
            if (out_len) *out_len = total;
            if (rc != MCP_OK && rc != MCP_E_CAP) return rc;
            MCP_TRY(d2h(ctx, out_off, ctx->dIo3.p, (size_t)(nlists + 1) * sizeof(int64_t)));
            if (rc == MCP_E_CAP || (total > 0 && sizing)) {
                sync(ctx);
                return fail(ctx, MCP_E_CAP, "encode: output buffer too small");
            }
            MCP_TRY(d2h(ctx, out, ctx->dIo2.p, (size_t)total));
            return sync(ctx);
        }
    }
    MCP_TRY(codec_encode_plan(ctx, codec_id, framing, ctx->dIo0.as<int32_t>(), ctx->dIo1.as<int64_t>(), 0, nlists,
                              ctx->dIo3.as<int64_t>(), &total));
    if (out_len) *out_len = total;
    MCP_TRY(d2h(ctx, out_off, ctx->dIo3.p, (size_t)(nlists + 1) * sizeof(int64_t)));
    if (total > cap || (total > 0 && !out)) {
        sync(ctx);
        return fail(ctx, MCP_E_CAP, "encode: output buffer too small");
    }
    MCP_CUDA(ctx, ctx->dIo2.reserve((size_t)total + 16));
    MCP_TRY(codec_encode_emit(ctx, codec_id, framing, ctx->dIo0.as<int32_t>(), ctx->dIo1.as<int64_t>(), 0, nlists,
                              ctx->dIo3.as<int64_t>(), ctx->dIo2.as<uint8_t>()));
    MCP_TRY(d2h(ctx, out, ctx->dIo2.p, (size_t)total));
    return sync(ctx);
}

int mcp_codec_decode(mcp_ctx *ctx, int codec_id, int framing, const uint8_t *in, const int64_t *in_off, int64_t nlists,
                     int32_t *out, const int64_t *out_off)
{
    MCP_TRY(set_dev(ctx));
    if (nlists < 0 || !in_off || !out_off) return fail(ctx, MCP_E_ARG, "decode: bad argument");
    if (!codec_ok(codec_id)) return fail(ctx, MCP_E_ARG, "unknown codec id");
    if (nlists == 0) return MCP_OK;
    if (in_off[0] != 0 || out_off[0] != 0) return fail(ctx, MCP_E_ARG, "decode: offsets must start at 0");
    const int64_t in_bytes = in_off[nlists], out_n = out_off[nlists];
    if (in_bytes < 0 || out_n < 0 || (in_bytes > 0 && !in) || (out_n > 0 && !out)) return fail(ctx, MCP_E_ARG, "decode: bad sizes");
    for (int64_t i = 0; i < nlists; ++i) {
        if (in_off[i + 1] < in_off[i] || out_off[i + 1] < out_off[i]) return fail(ctx, MCP_E_ARG, "decode: offsets not monotone");
        if ((in_off[i] & 3) != 0) return fail(ctx, MCP_E_FORMAT, "decode: byte offsets must be multiples of 4");
    }
    if (in_bytes & 3) return fail(ctx, MCP_E_FORMAT, "decode: byte length not a multiple of 4");
    MCP_CUDA(ctx, ctx->dIo0.reserve((size_t)in_bytes + 16));
    MCP_CUDA(ctx, ctx->dIo1.reserve((size_t)(nlists + 1) * sizeof(int64_t)));
    MCP_CUDA(ctx, ctx->dIo3.reserve((size_t)(nlists + 1) * sizeof(int64_t)));
    MCP_CUDA(ctx, ctx->dIo2.reserve((size_t)(out_n + 1) * sizeof(int32_t)));
    MCP_TRY(h2d(ctx, ctx->dIo0.p, in, (size_t)in_bytes));
    MCP_TRY(h2d(ctx, ctx->dIo1.p, in_off, (size_t)(nlists + 1) * sizeof(int64_t)));
    MCP_TRY(h2d(ctx, ctx->dIo3.p, out_off, (size_t)(nlists + 1) * sizeof(int64_t)));
    int rc = MCP_E_UNSUPPORTED;
    if (ctx->fastCodec && (framing == MCP_FRAME_WORDS || framing == MCP_FRAME_APP) && codec_fast_applicable(codec_id, out_n, nlists, false))
        rc = codec_fast_decode(ctx, codec_id, framing, ctx->dIo0.as<uint8_t>(), ctx->dIo1.as<int64_t>(), nlists, out_n,
                               ctx->dIo2.as<int32_t>(), ctx->dIo3.as<int64_t>());
    if (rc == MCP_E_UNSUPPORTED)
        rc = codec_decode_device(ctx, codec_id, framing, ctx->dIo0.as<uint8_t>(), ctx->dIo1.as<int64_t>(), nlists,
                                 ctx->dIo2.as<int32_t>(), ctx->dIo3.as<int64_t>());
    MCP_TRY(rc);
    MCP_TRY(d2h(ctx, out, ctx->dIo2.p, (size_t)out_n * sizeof(int32_t)));
    return sync(ctx);
}

// ------------------------------------------------------- single-list drop-ins
int mcp_codec_compress(mcp_ctx *ctx, int codec_id, const int32_t *in, int32_t inlength, int32_t *out, int32_t out_cap,
                       int32_t *out_words)
{
    if (!ctx || inlength < 0 || out_cap < 0 || !out_words) return ctx ? fail(ctx, MCP_E_ARG, "compress: bad argument") : MCP_E_ARG;
    const int64_t lo[2] = {0, inlength};
    int64_t oo[2] = {0, 0}, len = 0;
    const int rc = mcp_codec_encode(ctx, codec_id, MCP_FRAME_WORDS, in, lo, 1, oo, reinterpret_cast<uint8_t *>(out),
                                    (int64_t)out_cap * 4, &len);
    *out_words = (int32_t)(len / 4);
    return rc;
}

static int decode_single(mcp_ctx *ctx, int codec_id, int framing, const uint8_t *bytes, int64_t nbytes, int32_t *out,
                         int64_t out_cap, int64_t *out_count)
{
    MCP_TRY(set_dev(ctx));
    if (nbytes < 0 || out_cap < 0 || (nbytes > 0 && !bytes)) return fail(ctx, MCP_E_ARG, "uncompress: bad argument");
    if (!codec_ok(codec_id)) return fail(ctx, MCP_E_ARG, "unknown codec id");
    if (nbytes & 3) return fail(ctx, MCP_E_FORMAT, "uncompress: byte length not a multiple of 4");
    *out_count = 0;
    const int64_t in_off[2] = {0, nbytes};
    MCP_CUDA(ctx, ctx->dIo0.reserve((size_t)nbytes + 16));
    MCP_CUDA(ctx, ctx->dIo1.reserve(2 * sizeof(int64_t)));
    MCP_CUDA(ctx, ctx->dIo3.reserve(2 * sizeof(int64_t)));
    MCP_TRY(h2d(ctx, ctx->dIo0.p, bytes, (size_t)nbytes));
    MCP_TRY(h2d(ctx, ctx->dIo1.p, in_off, sizeof in_off));
    // length of the list, from the stream's own headers
    MCP_TRY(codec_decode_counts(ctx, codec_id, framing, ctx->dIo0.as<uint8_t>(), ctx->dIo1.as<int64_t>(), 1,
                                ctx->dIo3.as<int64_t>() + 1));
    int64_t n = 0;
    MCP_TRY(d2h(ctx, &n, ctx->dIo3.as<int64_t>() + 1, sizeof n));
    MCP_TRY(sync(ctx));
    if (n > out_cap) return fail(ctx, MCP_E_CAP, "uncompress: output array too small");
    const int64_t out_off[2] = {0, n};
    MCP_TRY(h2d(ctx, ctx->dIo3.p, out_off, sizeof out_off));
    MCP_CUDA(ctx, ctx->dIo2.reserve((size_t)(n + 1) * sizeof(int32_t)));
    MCP_TRY(codec_decode_device(ctx, codec_id, framing, ctx->dIo0.as<uint8_t>(), ctx->dIo1.as<int64_t>(), 1,
                                ctx->dIo2.as<int32_t>(), ctx->dIo3.as<int64_t>()));
    MCP_TRY(d2h(ctx, out, ctx->dIo2.p, (size_t)n * sizeof(int32_t)));
    MCP_TRY(sync(ctx));
    *out_count = n;
    return MCP_OK;
}

int mcp_codec_uncompress(mcp_ctx *ctx, int codec_id, const int32_t *in, int32_t inlength, int32_t *out, int32_t out_cap,
                         int32_t *out_count)
{
    if (!ctx || inlength < 0 || !out_count) return ctx ? fail(ctx, MCP_E_ARG, "uncompress: bad argument") : MCP_E_ARG;
    int64_t n = 0;
    const int rc = decode_single(ctx, codec_id, MCP_FRAME_WORDS, reinterpret_cast<const uint8_t *>(in), (int64_t)inlength * 4,
                                 out, out_cap, &n);
    *out_count = (int32_t)n;
    return rc;
}

// SkippableIntegerCODEC.headlessCompress / headlessUncompress (SkippableIntegerCODEC.java:43-66) of the composition behind a
// MCP_CODEC_SK_* id: IntCompressor's stream is [n][headless stream] (IntCompressor.java:38-46), so the headless stream is
// that stream without its first word, and decoding it is decoding [num] + stream.
static bool sk_codec(int codec_id)
{
    const int c = codec_id & MCP_CODEC_MASK;
    return c == MCP_CODEC_SK_BP32_VB || c == MCP_CODEC_SK_IBP32_IVB || c == MCP_CODEC_SK_FASTPFOR256_VB || c == MCP_CODEC_SK_FASTPFOR128_VB;
}

int mcp_codec_headless_compress(mcp_ctx *ctx, int codec_id, const int32_t *in, int32_t inlength, int32_t *out, int32_t out_cap,
                                int32_t *out_words)
{
    if (!ctx || inlength < 0 || out_cap < 0 || !out_words) return ctx ? fail(ctx, MCP_E_ARG, "headlessCompress: bad argument") : MCP_E_ARG;
    if (!codec_ok(codec_id) || !sk_codec(codec_id)) return fail(ctx, MCP_E_ARG, "headlessCompress: not a skippable codec id");
    *out_words = 0;
    if (inlength == 0) return MCP_OK; // (SkippableComposition writes nothing for an empty input)
    const int64_t lo[2] = {0, inlength};
    int64_t oo[2] = {0, 0}, len = 0;
    std::vector<int32_t> tmp((size_t)out_cap + 1);
    const int rc = mcp_codec_encode(ctx, codec_id, MCP_FRAME_WORDS, in, lo, 1, oo, reinterpret_cast<uint8_t *>(tmp.data()),
                                    ((int64_t)out_cap + 1) * 4, &len);
    if (rc != MCP_OK) return rc;
    *out_words = (int32_t)(len / 4 - 1);
    if (*out_words > 0) memcpy(out, tmp.data() + 1, (size_t)*out_words * sizeof(int32_t));
    return MCP_OK;
}

int mcp_codec_headless_uncompress(mcp_ctx *ctx, int codec_id, const int32_t *in, int32_t inlength, int32_t *out, int32_t num,
                                  int32_t *in_words)
{
    if (!ctx || inlength < 0 || num < 0) return ctx ? fail(ctx, MCP_E_ARG, "headlessUncompress: bad argument") : MCP_E_ARG;
    if (!codec_ok(codec_id) || !sk_codec(codec_id)) return fail(ctx, MCP_E_ARG, "headlessUncompress: not a skippable codec id");
    if (in_words) *in_words = 0;
    if (num == 0) return MCP_OK;
    MCP_TRY(set_dev(ctx));
    // [num] + stream through the batched decoder with the count known; the words consumed come from a re-encode-free
    // walk of the headers: decode, then measure what an encoder would have written for exactly these ints
    std::vector<int32_t> tmp((size_t)inlength + 1);
    tmp[0] = num;
    if (inlength > 0) memcpy(tmp.data() + 1, in, (size_t)inlength * sizeof(int32_t));
    const int64_t io[2] = {0, ((int64_t)inlength + 1) * 4}, oo[2] = {0, num};
    MCP_TRY(mcp_codec_decode(ctx, codec_id, MCP_FRAME_WORDS, reinterpret_cast<const uint8_t *>(tmp.data()), io, 1, out, oo));
    if (in_words) {
        const int64_t lo[2] = {0, num};
        int64_t eo[2] = {0, 0}, len = 0;
        const int rc = mcp_codec_encode(ctx, codec_id, MCP_FRAME_WORDS, out, lo, 1, eo, nullptr, 0, &len);
        if (rc != MCP_OK && rc != MCP_E_CAP) return rc;
        *in_words = (int32_t)(len / 4 - 1);
    }
    return MCP_OK;
}

int mcp_get_compressed_instruments(mcp_ctx *ctx, int codec_id, const int32_t *instruments, int32_t n, uint8_t *bytes,
                                   int64_t cap, int64_t *nbytes)
{
    if (!ctx || n < 0 || !nbytes) return ctx ? fail(ctx, MCP_E_ARG, "getCompressedInstruments: bad argument") : MCP_E_ARG;
    const int64_t lo[2] = {0, n};
    int64_t oo[2] = {0, 0};
    return mcp_codec_encode(ctx, codec_id, MCP_FRAME_APP, instruments, lo, 1, oo, bytes, cap, nbytes);
}

int mcp_get_decompressed_instruments(mcp_ctx *ctx, int codec_id, const uint8_t *bytes, int64_t nbytes,
                                     int32_t number_of_integers, int32_t *out)
{
    if (!ctx || number_of_integers < 0) return ctx ? fail(ctx, MCP_E_ARG, "getDecompressedInstruments: bad argument") : MCP_E_ARG;
    int64_t n = 0;
    MCP_TRY(decode_single(ctx, codec_id, MCP_FRAME_APP, bytes, nbytes, out, number_of_integers, &n));
    for (int64_t i = n; i < number_of_integers; ++i) out[i] = 0; // Java: new int[numberOfIntegers] is zero-filled
    return MCP_OK;
}

// ------------------------------------------------------------ weights column
int64_t mcp_weights_max_bytes(int64_t n)
{
    if (n < 0) return -1;
    return 32 + 8 * n + (8 * n) / 6; // snappy's MaxCompressedLength over the 8 n bytes
}

int mcp_weights_encode(mcp_ctx *ctx, const double *weights, const int64_t *list_off, int64_t nlists, int64_t *out_off, uint8_t *out,
                       int64_t cap, int64_t *out_len)
{
    MCP_TRY(set_dev(ctx));
    if (nlists < 0 || !list_off || !out_off || !out_len || cap < 0) return fail(ctx, MCP_E_ARG, "weights encode: bad argument");
    *out_len = 0;
    out_off[0] = 0;
    if (nlists == 0) return MCP_OK;
    if (list_off[0] != 0) return fail(ctx, MCP_E_ARG, "weights encode: offsets must start at 0");
    for (int64_t i = 0; i < nlists; ++i)
        if (list_off[i + 1] < list_off[i]) return fail(ctx, MCP_E_ARG, "weights encode: offsets not monotone");
    const int64_t n = list_off[nlists];
    if (n > 0 && !weights) return fail(ctx, MCP_E_ARG, "weights encode: null weights");
    MCP_CUDA(ctx, ctx->dIo0.reserve((size_t)(n + 2) * sizeof(double)));
    MCP_CUDA(ctx, ctx->dIo1.reserve((size_t)(nlists + 1) * sizeof(int64_t)));
    MCP_CUDA(ctx, ctx->dIo3.reserve((size_t)(nlists + 1) * sizeof(int64_t)));
    MCP_TRY(h2d(ctx, ctx->dIo0.p, weights, (size_t)n * sizeof(double)));
    MCP_TRY(h2d(ctx, ctx->dIo1.p, list_off, (size_t)(nlists + 1) * sizeof(int64_t)));
    int64_t total = 0;
    MCP_TRY(weights_encode_plan(ctx, ctx->dIo0.as<double>(), ctx->dIo1.as<int64_t>(), nlists, ctx->dIo3.as<int64_t>(), &total));
    *out_len = total;
    MCP_TRY(d2h(ctx, out_off, ctx->dIo3.p, (size_t)(nlists + 1) * sizeof(int64_t)));
    if (total > cap || (total > 0 && !out)) { MCP_TRY(sync(ctx)); return fail(ctx, MCP_E_CAP, "weights encode: output buffer too small"); }
    MCP_CUDA(ctx, ctx->dIo2.reserve((size_t)total + 16));
    MCP_TRY(weights_encode_emit(ctx, ctx->dIo0.as<double>(), ctx->dIo1.as<int64_t>(), nlists, ctx->dIo3.as<int64_t>(), ctx->dIo2.as<uint8_t>()));
    MCP_TRY(d2h(ctx, out, ctx->dIo2.p, (size_t)total));
    return sync(ctx);
}

int mcp_weights_decode(mcp_ctx *ctx, const uint8_t *in, const int64_t *in_off, int64_t nlists, double *out, const int64_t *out_off)
{
    MCP_TRY(set_dev(ctx));
    if (nlists < 0 || !in_off || !out_off) return fail(ctx, MCP_E_ARG, "weights decode: bad argument");
    if (nlists == 0) return MCP_OK;
    if (in_off[0] != 0 || out_off[0] != 0) return fail(ctx, MCP_E_ARG, "weights decode: offsets must start at 0");
    for (int64_t i = 0; i < nlists; ++i)
        if (in_off[i + 1] < in_off[i] || out_off[i + 1] < out_off[i]) return fail(ctx, MCP_E_ARG, "weights decode: offsets not monotone");
    const int64_t in_bytes = in_off[nlists], n = out_off[nlists];
    if ((in_bytes > 0 && !in) || (n > 0 && !out)) return fail(ctx, MCP_E_ARG, "weights decode: null buffer");
    MCP_CUDA(ctx, ctx->dIo0.reserve((size_t)in_bytes + 16));
    MCP_CUDA(ctx, ctx->dIo1.reserve((size_t)(nlists + 1) * sizeof(int64_t)));
    MCP_CUDA(ctx, ctx->dIo3.reserve((size_t)(nlists + 1) * sizeof(int64_t)));
    MCP_CUDA(ctx, ctx->dIo2.reserve((size_t)(n + 2) * sizeof(double)));
    MCP_TRY(h2d(ctx, ctx->dIo0.p, in, (size_t)in_bytes));
    MCP_TRY(h2d(ctx, ctx->dIo1.p, in_off, (size_t)(nlists + 1) * sizeof(int64_t)));
    MCP_TRY(h2d(ctx, ctx->dIo3.p, out_off, (size_t)(nlists + 1) * sizeof(int64_t)));
    MCP_TRY(weights_decode_device(ctx, ctx->dIo0.as<uint8_t>(), ctx->dIo1.as<int64_t>(), nlists, ctx->dIo2.as<double>(), ctx->dIo3.as<int64_t>()));
    MCP_TRY(d2h(ctx, out, ctx->dIo2.p, (size_t)n * sizeof(double)));
    return sync(ctx);
}

int mcp_get_compressed_weights(mcp_ctx *ctx, const double *weights, int32_t n, uint8_t *bytes, int64_t cap, int64_t *nbytes)
{
    if (!ctx || n < 0 || !nbytes) return ctx ? fail(ctx, MCP_E_ARG, "getCompressedWeights: bad argument") : MCP_E_ARG;
    const int64_t lo[2] = {0, n};
    int64_t oo[2] = {0, 0};
    return mcp_weights_encode(ctx, weights, lo, 1, oo, bytes, cap, nbytes);
}

int mcp_get_decompressed_weights(mcp_ctx *ctx, const uint8_t *bytes, int64_t nbytes, int32_t number_of_doubles, double *out)
{
    if (!ctx || number_of_doubles < 0 || nbytes < 0) return ctx ? fail(ctx, MCP_E_ARG, "getDecompressedWeights: bad argument") : MCP_E_ARG;
    const int64_t io[2] = {0, nbytes}, oo[2] = {0, number_of_doubles};
    return mcp_weights_decode(ctx, bytes, io, 1, out, oo);
}

// ------------------------------------------------------------------ risk path
static int upload_matrix(mcp_ctx *ctx, mcp::DevBuf &buf, const double *h, int64_t rows, int32_t F)
{
    if (rows < 0 || F <= 0 || (rows > 0 && !h)) return fail(ctx, MCP_E_ARG, "upload: bad argument");
    MCP_CUDA(ctx, buf.reserve((size_t)rows * F * sizeof(double) + 16));
    MCP_TRY(h2d(ctx, buf.p, h, (size_t)rows * F * sizeof(double)));
    return sync(ctx);
}

int mcp_upload_exposures(mcp_ctx *ctx, const double *E, int64_t P, int32_t F)
{
    MCP_TRY(set_dev(ctx));
    if (ctx->haveS && ctx->F != F) ctx->haveS = false; // a new factor count invalidates the other matrix
    MCP_TRY(upload_matrix(ctx, ctx->dE, E, P, F));
    ctx->P = P; ctx->F = F; ctx->haveE = true; ctx->haveRun = false;
    return MCP_OK;
}

int mcp_build_exposures(mcp_ctx *ctx, const int32_t *instruments, const double *weights, const int64_t *pos_off, int64_t P,
                        const double *loadings, int64_t n_instruments, int32_t F)
{
    MCP_TRY(set_dev(ctx));
    if (P < 0 || F <= 0 || n_instruments <= 0 || !pos_off || !loadings) return fail(ctx, MCP_E_ARG, "build_exposures: bad argument");
    if (pos_off[0] != 0) return fail(ctx, MCP_E_ARG, "build_exposures: offsets must start at 0");
    for (int64_t p = 0; p < P; ++p)
        if (pos_off[p + 1] < pos_off[p]) return fail(ctx, MCP_E_ARG, "build_exposures: offsets not monotone");
    const int64_t npos = pos_off[P];
    if (npos > 0 && (!instruments || !weights)) return fail(ctx, MCP_E_ARG, "build_exposures: null positions");
    if (ctx->haveS && ctx->F != F) ctx->haveS = false;
    ctx->haveE = false;
    MCP_CUDA(ctx, ctx->dE.reserve((size_t)P * F * sizeof(double) + 16));
    MCP_CUDA(ctx, ctx->dIo0.reserve((size_t)(npos + 1) * sizeof(int32_t)));
    MCP_CUDA(ctx, ctx->dIo1.reserve((size_t)(P + 1) * sizeof(int64_t)));
    MCP_CUDA(ctx, ctx->dIo2.reserve((size_t)(npos + 1) * sizeof(double)));
    MCP_CUDA(ctx, ctx->dIo3.reserve((size_t)n_instruments * F * sizeof(double)));
    MCP_TRY(h2d(ctx, ctx->dIo0.p, instruments, (size_t)npos * sizeof(int32_t)));
    MCP_TRY(h2d(ctx, ctx->dIo1.p, pos_off, (size_t)(P + 1) * sizeof(int64_t)));
    MCP_TRY(h2d(ctx, ctx->dIo2.p, weights, (size_t)npos * sizeof(double)));
    MCP_TRY(h2d(ctx, ctx->dIo3.p, loadings, (size_t)n_instruments * F * sizeof(double)));
    MCP_TRY(build_exposures(ctx, ctx->dIo0.as<int32_t>(), ctx->dIo2.as<double>(), ctx->dIo1.as<int64_t>(), P, ctx->dIo3.as<double>(),
                            n_instruments, F, ctx->dE.as<double>()));
    ctx->P = P; ctx->F = F; ctx->haveE = true; ctx->haveRun = false;
    return MCP_OK;
}

int mcp_upload_scenarios(mcp_ctx *ctx, const double *S, int64_t N, int32_t F)
{
    MCP_TRY(set_dev(ctx));
    if (ctx->haveE && ctx->F != F) ctx->haveE = false;
    MCP_TRY(upload_matrix(ctx, ctx->dS, S, N, F));
    ctx->N = N; ctx->F = F; ctx->haveS = true; ctx->haveRun = false; ctx->sfragValid = false;
    return MCP_OK;
}

int mcp_generate_exposures(mcp_ctx *ctx, int64_t P, int32_t F, uint64_t seed, double scale, double shift, int64_t row0)
{
    MCP_TRY(set_dev(ctx));
    if (P < 0 || F <= 0) return fail(ctx, MCP_E_ARG, "generate_exposures: bad size");
    if (ctx->haveS && ctx->F != F) ctx->haveS = false;
    MCP_CUDA(ctx, ctx->dE.reserve((size_t)P * F * sizeof(double) + 16));
    MCP_TRY(gen_matrix(ctx, ctx->dE.as<double>(), P, F, seed, 1u, scale, shift, row0));
    ctx->P = P; ctx->F = F; ctx->haveE = true; ctx->haveRun = false;
    return MCP_OK;
}

int mcp_generate_scenarios(mcp_ctx *ctx, int64_t N, int32_t F, uint64_t seed, double scale, double shift, int64_t row0)
{
    MCP_TRY(set_dev(ctx));
    if (N < 0 || F <= 0) return fail(ctx, MCP_E_ARG, "generate_scenarios: bad size");
    if (ctx->haveE && ctx->F != F) ctx->haveE = false;
    MCP_CUDA(ctx, ctx->dS.reserve((size_t)N * F * sizeof(double) + 16));
    MCP_TRY(gen_matrix(ctx, ctx->dS.as<double>(), N, F, seed, 2u, scale, shift, row0));
    ctx->N = N; ctx->F = F; ctx->haveS = true; ctx->haveRun = false; ctx->sfragValid = false;
    return MCP_OK;
}

int mcp_download_exposures(mcp_ctx *ctx, double *E)
{
    MCP_TRY(set_dev(ctx));
    if (!ctx->haveE || !E) return fail(ctx, MCP_E_STATE, "download_exposures: nothing uploaded");
    MCP_TRY(d2h(ctx, E, ctx->dE.p, (size_t)ctx->P * ctx->F * sizeof(double)));
    return sync(ctx);
}

int mcp_download_scenarios(mcp_ctx *ctx, double *S)
{
    MCP_TRY(set_dev(ctx));
    if (!ctx->haveS || !S) return fail(ctx, MCP_E_STATE, "download_scenarios: nothing uploaded");
    MCP_TRY(d2h(ctx, S, ctx->dS.p, (size_t)ctx->N * ctx->F * sizeof(double)));
    return sync(ctx);
}

int mcp_value(mcp_ctx *ctx, int64_t p0, int64_t p1, double *V)
{
    MCP_TRY(set_dev(ctx));
    if (!ctx->haveE || !ctx->haveS) return fail(ctx, MCP_E_STATE, "value: upload exposures and scenarios first");
    if (p0 < 0 || p1 > ctx->P || p1 < p0 || !V) return fail(ctx, MCP_E_ARG, "value: bad range");
    const size_t bytes = (size_t)(p1 - p0) * ctx->N * sizeof(double);
    MCP_CUDA(ctx, ctx->dLoss.reserve(bytes + 16));
    MCP_TRY(risk_value(ctx, p0, p1, ctx->dLoss.as<double>()));
    MCP_TRY(d2h(ctx, V, ctx->dLoss.p, bytes));
    return sync(ctx);
}

int mcp_run(mcp_ctx *ctx, double alpha, int codec_id, int framing)
{
    MCP_TRY(set_dev(ctx));
    if (!ctx->haveE || !ctx->haveS) return fail(ctx, MCP_E_STATE, "run: upload exposures and scenarios first");
    if (!(alpha > 0.0 && alpha <= 1.0)) return fail(ctx, MCP_E_ARG, "run: alpha must be in (0,1]");
    if (!codec_ok(codec_id)) return fail(ctx, MCP_E_ARG, "unknown codec id");
    if (framing != MCP_FRAME_WORDS && framing != MCP_FRAME_APP) return fail(ctx, MCP_E_ARG, "unknown framing");
    return risk_run(ctx, alpha, codec_id, framing);
}

int mcp_run_sizes(mcp_ctx *ctx, int64_t *P, int64_t *N, int64_t *tail, int64_t *enc_bytes)
{
    if (!ctx) return MCP_E_ARG;
    if (!ctx->haveRun) return fail(ctx, MCP_E_STATE, "run_sizes: no completed run");
    if (P) *P = ctx->runP;
    if (N) *N = ctx->runN;
    if (tail) *tail = ctx->runTail;
    if (enc_bytes) *enc_bytes = ctx->runEncBytes;
    return MCP_OK;
}

int mcp_run_stats(mcp_ctx *ctx, int64_t *fallback_rows)
{
    if (!ctx) return MCP_E_ARG;
    if (!ctx->haveRun) return fail(ctx, MCP_E_STATE, "run_stats: no completed run");
    if (fallback_rows) *fallback_rows = ctx->runFallbackRows;
    return MCP_OK;
}

int mcp_codec_stats(mcp_ctx *ctx, int64_t *fast_encodes, int64_t *fast_fallbacks)
{
    if (!ctx) return MCP_E_ARG;
    if (fast_encodes) *fast_encodes = ctx->fastEncodeCalls;
    if (fast_fallbacks) *fast_fallbacks = ctx->fastEncodeFallbacks;
    return MCP_OK;
}

int mcp_codec_stats_short(mcp_ctx *ctx, int64_t *short_calls, int64_t *short_fallbacks)
{
    if (!ctx) return MCP_E_ARG;
    if (short_calls) *short_calls = ctx->shortListCalls;
    if (short_fallbacks) *short_fallbacks = ctx->shortListFallbacks;
    return MCP_OK;
}

int mcp_fetch(mcp_ctx *ctx, double *mean, double *variance, double *var_loss, int32_t *var_trial, double *cvar,
              int32_t *breach_idx, int64_t *enc_off, uint8_t *enc, int64_t enc_cap)
{
    MCP_TRY(set_dev(ctx));
    if (!ctx->haveRun) return fail(ctx, MCP_E_STATE, "fetch: no completed run");
    const size_t P = (size_t)ctx->runP, r0 = (size_t)ctx->runRow0; // r0 > 0: the slice a trial-sharded run left on this rank
    if (enc && enc_cap < ctx->runEncBytes) return fail(ctx, MCP_E_CAP, "fetch: encoded buffer too small");
    if (mean) MCP_TRY(d2h(ctx, mean, ctx->dMean.as<double>() + r0, P * sizeof(double)));
    if (variance) MCP_TRY(d2h(ctx, variance, ctx->dVar.as<double>() + r0, P * sizeof(double)));
    if (var_loss) MCP_TRY(d2h(ctx, var_loss, ctx->dVarLoss.as<double>() + r0, P * sizeof(double)));
    if (var_trial) MCP_TRY(d2h(ctx, var_trial, ctx->dVarTrial.as<int32_t>() + r0, P * sizeof(int32_t)));
    if (cvar) MCP_TRY(d2h(ctx, cvar, ctx->dCvar.as<double>() + r0, P * sizeof(double)));
    if (breach_idx) MCP_TRY(d2h(ctx, breach_idx, ctx->dBreach.as<int32_t>() + r0 * (size_t)ctx->runTail, P * (size_t)ctx->runTail * sizeof(int32_t)));
    if (enc_off) MCP_TRY(d2h(ctx, enc_off, ctx->dEncOff.p, (P + 1) * sizeof(int64_t)));
    if (enc) MCP_TRY(d2h(ctx, enc, ctx->dEnc.p, (size_t)ctx->runEncBytes));
    return sync(ctx);
}

static bool mma_factor_count(int32_t F) { return F == 16 || F == 32 || F == 48 || F == 64; }

// Calls in flight on one GPU (one context and one host thread each, so that a step's copies ride under another step's
// kernels) drift into step when their kernels are allowed to interleave: both then compute at half speed, finish together
// and copy together with the SMs idle.  With turns the kernels of the calls run one call after the other ON THE GPU: a call
// enqueues its uploads, takes the device's token, makes its stream wait for the event the previous holder recorded behind
// its last kernel, enqueues its own kernels, records its own event and hands the token on -- all before its first host
// sync, so the token is held for the ~0.1 ms of enqueueing, not for the milliseconds the kernels take, and the next
// call's kernels are already queued when this call's finish.  Downloads happen outside.
static constexpr double kSimulateTurnBytes = 48e6; // uploads per call from which turns pay (measured: profiles/r2/optimization_log.md)

static bool simulate_takes_turns(const mcp_ctx *ctx, int64_t P, int64_t N, int32_t F)
{
    if (ctx->simulateTurns >= 0) return ctx->simulateTurns != 0;
    return (double)(P + N) * F * sizeof(double) >= kSimulateTurnBytes;
}

int mcp_simulate(mcp_ctx *ctx, const double *E, int64_t P, const double *S, int64_t N, int32_t F, double alpha,
                 int codec_id, int framing, double *mean, double *variance, double *var_loss, int32_t *var_trial,
                 double *cvar, int32_t *breach_idx, int64_t *enc_off, uint8_t *enc, int64_t enc_cap, int64_t *enc_len)
{
    MCP_TRY(set_dev(ctx));
    if (P < 0 || N < 0 || F <= 0 || (P > 0 && !E) || (N > 0 && !S)) return fail(ctx, MCP_E_ARG, "simulate: bad argument");
    if (!(alpha > 0.0 && alpha <= 1.0)) return fail(ctx, MCP_E_ARG, "run: alpha must be in (0,1]");
    if (!codec_ok(codec_id)) return fail(ctx, MCP_E_ARG, "unknown codec id");
    if (framing != MCP_FRAME_WORDS && framing != MCP_FRAME_APP) return fail(ctx, MCP_E_ARG, "unknown framing");
    ctx->haveE = ctx->haveS = false; // (only once the arguments are known to be good: a rejected call keeps earlier uploads)
    const int64_t t = mcp_tail_size(N, alpha);
    const int64_t head = N > 0 ? risk_first_step_len(N, t) : 0;
    if (!mma_factor_count(F) || ctx->forceDense || head >= N || P == 0) {
        // plain sequence: upload, run, fetch
        MCP_TRY(mcp_upload_exposures(ctx, E, P, F));
        MCP_TRY(mcp_upload_scenarios(ctx, S, N, F));
        {
            int rc = simulate_takes_turns(ctx, P, N, F) ? simulate_turn_begin(ctx) : MCP_OK;
            if (rc == MCP_OK) rc = mcp_run(ctx, alpha, codec_id, framing);
            simulate_turn_end(ctx); // (if the run did not get as far as its own hand-over)
            if (rc != MCP_OK) return rc;
        }
        if (enc_len) *enc_len = ctx->runEncBytes;
        return mcp_fetch(ctx, mean, variance, var_loss, var_trial, cvar, breach_idx, enc_off, enc, enc_cap);
    }
    // Pipelined: the dense first step needs E and only the first `head` scenario rows, so the bulk of S (96 % at C2)
    // is uploaded on the side stream while that step and its selection run; on the way out the raw breach lists (if
    // wanted) leave on the side stream under the encode.  (Host buffers should be pinned for the overlap to happen.)
    MCP_CUDA(ctx, ctx->dE.reserve((size_t)P * F * sizeof(double) + 16));
    MCP_CUDA(ctx, ctx->dS.reserve((size_t)N * F * sizeof(double) + 16));
    MCP_TRY(h2d(ctx, ctx->dE.p, E, (size_t)P * F * sizeof(double)));
    MCP_TRY(h2d(ctx, ctx->dS.p, S, (size_t)head * F * sizeof(double)));
    // the tail in two parts: the sparse valuation is launched in two halves, the first as soon as ITS scenario rows are
    // there (the whole tail takes longer to arrive than the dense step and its selection take to run)
    int64_t mid = (head + (N - head) / 2) & ~(int64_t)7;
    if (mid <= head || N - head < 16384 || !ctx->uploadSplit) mid = 0;
    const int64_t split = mid > 0 ? mid : N;
    MCP_CUDA(ctx, cudaMemcpyAsync(ctx->dS.as<double>() + head * F, S + head * F, (size_t)(split - head) * F * sizeof(double),
                                  cudaMemcpyHostToDevice, ctx->stream2));
    if (mid > 0) {
        MCP_CUDA(ctx, cudaEventRecord(ctx->evTail1, ctx->stream2));
        MCP_CUDA(ctx, cudaMemcpyAsync(ctx->dS.as<double>() + mid * F, S + mid * F, (size_t)(N - mid) * F * sizeof(double),
                                      cudaMemcpyHostToDevice, ctx->stream2));
    }
    MCP_CUDA(ctx, cudaEventRecord(ctx->evTail, ctx->stream2));
    ctx->P = P; ctx->N = N; ctx->F = F;
    ctx->haveE = ctx->haveS = true; ctx->haveRun = false; ctx->sfragValid = false;
    ctx->sTailPending = true; ctx->sHeadRows = head; ctx->sMidRows = mid;
    ctx->earlyBreachDst = breach_idx;
    ctx->earlyMean = mean; ctx->earlyVar = variance; ctx->earlyVarLoss = var_loss; ctx->earlyVarTrial = var_trial; ctx->earlyCvar = cvar;
    int rc = simulate_takes_turns(ctx, P, N, F) ? simulate_turn_begin(ctx) : MCP_OK;
    if (rc == MCP_OK) rc = mcp_run(ctx, alpha, codec_id, framing);
    simulate_turn_end(ctx); // (if the run did not get as far as its own hand-over)
    ctx->earlyBreachDst = nullptr;
    ctx->earlyMean = ctx->earlyVar = ctx->earlyVarLoss = ctx->earlyCvar = nullptr;
    ctx->earlyVarTrial = nullptr;
    ctx->sMidRows = 0;
    if (ctx->sTailPending) { // the run failed before it consumed the tail: drain the copy, keep the state consistent
        ctx->sTailPending = false;
        cudaStreamSynchronize(ctx->stream2);
    }
    if (rc != MCP_OK) { cudaStreamSynchronize(ctx->stream2); return rc; }
    if (enc_len) *enc_len = ctx->runEncBytes;
    // (the statistics left under the encode.)  Whatever the fetch says, the side stream's copies into the caller's
    // buffers must have landed before the call returns: the caller may free or reuse them right away.
    const int frc = mcp_fetch(ctx, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, enc_off, enc, enc_cap);
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream2)); // the statistics and the raw lists
    return frc;
}

// ---------------------------------------------------------- trial-sharded mode
int mcp_comm_unique_id(uint8_t id[128])
{
    if (!id) return MCP_E_ARG;
    return nccl_unique_id(id, nullptr);
}

int mcp_comm_init(mcp_ctx *ctx, const uint8_t id[128], int rank, int world)
{
    MCP_TRY(set_dev(ctx));
    if (!id) return fail(ctx, MCP_E_ARG, "comm_init: null id");
    return nccl_comm_init(ctx, id, rank, world);
}

int mcp_comm_destroy(mcp_ctx *ctx)
{
    MCP_TRY(set_dev(ctx));
    nccl_comm_destroy(ctx);
    return MCP_OK;
}

static int ts_check_args(mcp_ctx *ctx, double alpha, int codec_id, int framing)
{
    if (!ctx->haveE || !ctx->haveS) return fail(ctx, MCP_E_STATE, "trial-sharded run: upload exposures and scenarios first");
    if (!(alpha > 0.0 && alpha <= 1.0)) return fail(ctx, MCP_E_ARG, "run: alpha must be in (0,1]");
    if (!codec_ok(codec_id)) return fail(ctx, MCP_E_ARG, "unknown codec id");
    if (framing != MCP_FRAME_WORDS && framing != MCP_FRAME_APP) return fail(ctx, MCP_E_ARG, "unknown framing");
    return MCP_OK;
}

int mcp_ts_local(mcp_ctx *ctx, double alpha, int64_t n_global, int64_t trial_base, int world, void **d_block, int64_t *block_bytes)
{
    MCP_TRY(set_dev(ctx));
    MCP_TRY(ts_check_args(ctx, alpha, 0, 0));
    MCP_TRY(risk_ts_local(ctx, alpha, n_global, trial_base, world, d_block, block_bytes));
    return sync(ctx);
}

int mcp_ts_merge(mcp_ctx *ctx, const void *d_gathered, int rank, void **d_flags, int64_t *flag_bytes)
{
    MCP_TRY(set_dev(ctx));
    MCP_TRY(risk_ts_merge(ctx, d_gathered, rank, d_flags, flag_bytes));
    return sync(ctx);
}

int mcp_ts_flags(mcp_ctx *ctx, const void *d_gathered_flags, int64_t *flagged)
{
    MCP_TRY(set_dev(ctx));
    return risk_ts_flags(ctx, d_gathered_flags, flagged);
}

int mcp_ts_local2(mcp_ctx *ctx, void **d_block, int64_t *block_bytes)
{
    MCP_TRY(set_dev(ctx));
    MCP_TRY(risk_ts_local2(ctx, d_block, block_bytes));
    return sync(ctx);
}

int mcp_ts_merge2(mcp_ctx *ctx, const void *d_gathered2)
{
    MCP_TRY(set_dev(ctx));
    return risk_ts_merge2(ctx, d_gathered2);
}

int mcp_ts_finish(mcp_ctx *ctx, int rank, int world, int codec_id, int framing)
{
    MCP_TRY(set_dev(ctx));
    if (!codec_ok(codec_id)) return fail(ctx, MCP_E_ARG, "unknown codec id");
    if (framing != MCP_FRAME_WORDS && framing != MCP_FRAME_APP) return fail(ctx, MCP_E_ARG, "unknown framing");
    return risk_ts_finish(ctx, rank, world, codec_id, framing);
}

int64_t mcp_ts_candidates(int64_t n_global, double alpha, int world)
{
    if (n_global <= 0 || world < 1) return -1;
    return ts_candidates_per_rank(n_global, mcp_tail_size(n_global, alpha), world);
}

int mcp_run_trial_sharded(mcp_ctx *ctx, double alpha, int codec_id, int framing, int64_t n_global, int64_t trial_base)
{
    MCP_TRY(set_dev(ctx));
    MCP_TRY(ts_check_args(ctx, alpha, codec_id, framing));
    int rank = 0, world = 1;
    MCP_TRY(nccl_rank_world(ctx, &rank, &world));
    void *blk = nullptr;
    int64_t bytes = 0, flagged = 0;
    MCP_TRY(risk_ts_local(ctx, alpha, n_global, trial_base, world, &blk, &bytes));
    MCP_CUDA(ctx, ctx->dTsRecv.reserve((size_t)bytes * world));
    {
        // the one exchange of the common case.  A rank merges only ITS portfolio slice, so it needs only that slice of every
        // other rank's block: a grouped send / recv of slices (1/world of the all-gather's bytes, same addresses) unless the
        // option or the NCCL at hand says otherwise
        KScope ks(ctx, MCP_K_COMM);
        if (ctx->tsSliceExchange && world > 1 && world <= 64 && nccl_has_send_recv()) {
            ExchRegion regions[5];
            int64_t p0[65];
            MCP_TRY(risk_ts_exchange_regions(ctx, world, regions, p0));
            MCP_TRY(nccl_slice_exchange(ctx, reinterpret_cast<const unsigned char *>(blk), ctx->dTsRecv.as<unsigned char>(), (size_t)bytes, regions, 5, p0));
        } else
            MCP_TRY(nccl_all_gather(ctx, blk, ctx->dTsRecv.p, (size_t)bytes));
    }
    void *flags = nullptr;
    int64_t fbytes = 0;
    MCP_TRY(risk_ts_merge(ctx, ctx->dTsRecv.p, rank, &flags, &fbytes)); // this rank's portfolio slice
    MCP_CUDA(ctx, ctx->dTsFlagsAll.reserve((size_t)fbytes * world + 16));
    {
        KScope ks(ctx, MCP_K_COMM);
        MCP_TRY(nccl_all_gather(ctx, flags, ctx->dTsFlagsAll.p, (size_t)fbytes)); // one byte per row: which rows need round two
    }
    MCP_TRY(risk_ts_flags(ctx, ctx->dTsFlagsAll.p, &flagged));
    if (flagged > 0) { // same count on every rank: the collective below is entered by all of them
        MCP_TRY(risk_ts_local2(ctx, &blk, &bytes));
        MCP_CUDA(ctx, ctx->dTsRecv2.reserve((size_t)bytes * world));
        {
            KScope ks(ctx, MCP_K_COMM);
            MCP_TRY(nccl_all_gather(ctx, blk, ctx->dTsRecv2.p, (size_t)bytes));
        }
        MCP_TRY(risk_ts_merge2(ctx, ctx->dTsRecv2.p));
    }
    return risk_ts_finish(ctx, rank, world, codec_id, framing);
}

// ---------------------------------------------------------------- measurement
int mcp_profile_enable(mcp_ctx *ctx, int on)
{
    if (!ctx) return MCP_E_ARG;
    ctx->prof.on = on != 0;
    return MCP_OK;
}

int mcp_profile_reset(mcp_ctx *ctx)
{
    MCP_TRY(set_dev(ctx));
    MCP_TRY(sync(ctx));
    resolve_profile(ctx);
    for (int i = 0; i < MCP_K_COUNT; ++i) { ctx->prof.launches[i] = 0; ctx->prof.ms[i] = 0; }
    return MCP_OK;
}

int mcp_profile_get(mcp_ctx *ctx, int64_t launches[MCP_K_COUNT], double ms[MCP_K_COUNT])
{
    MCP_TRY(set_dev(ctx));
    MCP_TRY(sync(ctx));
    resolve_profile(ctx);
    for (int i = 0; i < MCP_K_COUNT; ++i) {
        if (launches) launches[i] = ctx->prof.launches[i];
        if (ms) ms[i] = ctx->prof.ms[i];
    }
    return MCP_OK;
}

int mcp_timer_start(mcp_ctx *ctx)
{
    MCP_TRY(set_dev(ctx));
    MCP_CUDA(ctx, cudaEventRecord(ctx->tA, ctx->stream));
    return MCP_OK;
}

int mcp_timer_stop(mcp_ctx *ctx, double *ms)
{
    MCP_TRY(set_dev(ctx));
    MCP_CUDA(ctx, cudaEventRecord(ctx->tB, ctx->stream));
    MCP_CUDA(ctx, cudaEventSynchronize(ctx->tB));
    float f = 0;
    MCP_CUDA(ctx, cudaEventElapsedTime(&f, ctx->tA, ctx->tB));
    if (ms) *ms = f;
    return MCP_OK;
}

int mcp_fp64_peak(mcp_ctx *ctx, double *tflops)
{
    MCP_TRY(set_dev(ctx));
    if (!tflops) return fail(ctx, MCP_E_ARG, "fp64_peak: null");
    return fp64_peak(ctx, tflops);
}

int mcp_fp64_mma_peak(mcp_ctx *ctx, double *tflops)
{
    MCP_TRY(set_dev(ctx));
    if (!tflops) return fail(ctx, MCP_E_ARG, "fp64_mma_peak: null");
    return fp64_mma_peak(ctx, tflops);
}

int mcp_flush_l2(mcp_ctx *ctx)
{
    MCP_TRY(set_dev(ctx));
    const size_t bytes = (size_t)256 << 20; // 2x the 126 MB L2
    MCP_CUDA(ctx, ctx->dFlush.reserve(bytes));
    MCP_CUDA(ctx, cudaMemsetAsync(ctx->dFlush.p, 0, bytes, ctx->stream));
    return MCP_OK;
}

int mcp_generate_breach_lists(mcp_ctx *ctx, int64_t nlists, int32_t n_trials, double rate, uint64_t seed,
                              int64_t first_list, const int32_t **d_vals, const int64_t **d_list_off, int64_t *total_vals)
{
    MCP_TRY(set_dev(ctx));
    if (nlists < 0 || n_trials < 0 || !(rate >= 0.0 && rate < 1.0) || !d_vals || !d_list_off || !total_vals)
        return fail(ctx, MCP_E_ARG, "generate_breach_lists: bad argument");
    int64_t total = 0;
    MCP_TRY(gen_breach_lists(ctx, nlists, n_trials, rate, seed, first_list, &total));
    *d_vals = ctx->dGenVals.as<int32_t>();
    *d_list_off = ctx->dGenOff.as<int64_t>();
    *total_vals = total;
    return MCP_OK;
}

int mcp_dev_alloc(mcp_ctx *ctx, size_t bytes, void **dptr)
{
    MCP_TRY(set_dev(ctx));
    if (!dptr) return fail(ctx, MCP_E_ARG, "dev_alloc: null");
    MCP_CUDA(ctx, cudaMalloc(dptr, bytes ? bytes : 16));
    return MCP_OK;
}

int mcp_dev_free(mcp_ctx *ctx, void *dptr)
{
    MCP_TRY(set_dev(ctx));
    MCP_CUDA(ctx, cudaFree(dptr));
    return MCP_OK;
}

int mcp_memcpy_h2d(mcp_ctx *ctx, void *dptr, const void *hptr, size_t bytes)
{
    MCP_TRY(set_dev(ctx));
    MCP_TRY(h2d(ctx, dptr, hptr, bytes));
    return sync(ctx);
}

int mcp_memcpy_d2h(mcp_ctx *ctx, void *hptr, const void *dptr, size_t bytes)
{
    MCP_TRY(set_dev(ctx));
    MCP_TRY(d2h(ctx, hptr, dptr, bytes));
    return sync(ctx);
}

int mcp_host_alloc_pinned(size_t bytes, void **hptr)
{
    if (!hptr) return MCP_E_ARG;
    if (cudaMallocHost(hptr, bytes ? bytes : 16) != cudaSuccess) {
        cudaGetLastError();
        return MCP_E_NOMEM;
    }
    return MCP_OK;
}

int mcp_host_free_pinned(void *hptr)
{
    if (cudaFreeHost(hptr) != cudaSuccess) {
        cudaGetLastError();
        return MCP_E_CUDA;
    }
    return MCP_OK;
}

} // extern "C"
