// weights_kernels.cu -- the loader's WEIGHTS column.
//
// The reference stores a portfolio's weights as Snappy.compress(double[]) -- the raw Snappy format over the
// little-endian bytes of the doubles (LoadPortfolios.java:563-595) -- when that is smaller than 8 n bytes, else as
// big-endian doubles under algo "none" (LoadPortfolios.java:270-278, BitManipulationHelper.java:193-208).
// snappy-java is an un-vendored dependency (maprdb-portfolio/pom.xml:50-56): PARITY UNPINNED for the compressed bytes.
// What is kept is the FORMAT: the decoder takes every valid raw-Snappy stream (literals, 1- / 2- / 4-byte-offset
// copies), so blobs snappy-java wrote decode here; the encoder is a greedy matcher stated in oracle/weights_oracle.c
// (1 024-entry table, h = x * 0x1e35a7bd >> 22, one-byte steps, snappy's own copy splitting) and its streams are valid
// Snappy that any decoder reads.  One warp per list: the element walk is serial by nature (lane 0), the other lanes
// only clear the table -- weights are off the timed hot path, correctness and format compatibility are the point.
#include "common.cuh"

namespace mcp {

namespace {

constexpr int WK_WARPS = 4;
constexpr int WK_TABLE = 1024;

__device__ __forceinline__ uint32_t ld32u(const uint8_t *base, int64_t pos)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(base); // lists start 8-byte aligned (doubles)
    const uint32_t lo = w[pos >> 2];
    const int sh = (int)(pos & 3) * 8;
    if (sh == 0) return lo;
    return __funnelshift_r(lo, w[(pos >> 2) + 1], sh);
}

template <bool WRITE>
__device__ __forceinline__ int64_t put_literal(const uint8_t *src, int64_t len, uint8_t *out)
{
    if (len <= 0) return 0;
    int64_t o = 0;
    const uint32_t n = (uint32_t)(len - 1);
    if (n < 60) { if (WRITE) out[o] = (uint8_t)(n << 2); ++o; }
    else {
        const int k = n < (1u << 8) ? 1 : n < (1u << 16) ? 2 : n < (1u << 24) ? 3 : 4;
        if (WRITE) out[o] = (uint8_t)((59 + k) << 2);
        ++o;
        for (int i = 0; i < k; ++i) { if (WRITE) out[o] = (uint8_t)(n >> (8 * i)); ++o; }
    }
    if (WRITE) for (int64_t i = 0; i < len; ++i) out[o + i] = src[i];
    return o + len;
}

template <bool WRITE>
__device__ __forceinline__ int64_t put_copy64(int64_t offset, int64_t len, uint8_t *out)
{
    if (len < 12 && offset < 2048 && len >= 4) {
        if (WRITE) { out[0] = (uint8_t)(1 | ((len - 4) << 2) | ((offset >> 8) << 5)); out[1] = (uint8_t)(offset & 0xff); }
        return 2;
    }
    if (WRITE) { out[0] = (uint8_t)(2 | ((len - 1) << 2)); out[1] = (uint8_t)(offset & 0xff); out[2] = (uint8_t)(offset >> 8); }
    return 3;
}

template <bool WRITE>
__device__ __forceinline__ int64_t put_copy(int64_t offset, int64_t len, uint8_t *out)
{
    int64_t o = 0;
    while (len >= 68) { o += put_copy64<WRITE>(offset, 64, out + o); len -= 64; }
    if (len > 64) { o += put_copy64<WRITE>(offset, 60, out + o); len -= 60; }
    return o + put_copy64<WRITE>(offset, len, out + o);
}

template <bool WRITE>
__global__ void __launch_bounds__(WK_WARPS * 32) k_weights_encode(const double *w, const int64_t *list_off, int64_t nlists,
                                                                  const int64_t *out_off, uint8_t *out, int64_t *sizes)
{
    __shared__ int32_t s_table[WK_WARPS][WK_TABLE];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int32_t *table = s_table[wib];
    for (int64_t li = (int64_t)blockIdx.x * WK_WARPS + wib; li < nlists; li += (int64_t)gridDim.x * WK_WARPS) {
        for (int i = lane; i < WK_TABLE; i += 32) table[i] = -1;
        __syncwarp();
        if (lane == 0) {
            const uint8_t *in = reinterpret_cast<const uint8_t *>(w + list_off[li]);
            const int64_t len = 8 * (list_off[li + 1] - list_off[li]);
            uint8_t *dst = WRITE ? out + out_off[li] : nullptr;
            int64_t o = 0;
            uint64_t v = (uint64_t)len;
            do { uint8_t b = (uint8_t)(v & 0x7f); v >>= 7; if (v) b |= 0x80; if (WRITE) dst[o] = b; ++o; } while (v);
            int64_t pos = 0, lit = 0;
            while (pos + 4 <= len) {
                const uint32_t x = ld32u(in, pos);
                const uint32_t h = (x * 0x1e35a7bdu) >> 22;
                const int64_t cand = table[h];
                table[h] = (int32_t)pos;
                if (cand >= 0 && pos - cand <= 65535 && ld32u(in, cand) == x) {
                    int64_t m = 4;
                    while (pos + m < len && in[cand + m] == in[pos + m]) ++m;
                    o += put_literal<WRITE>(in + lit, pos - lit, dst + o);
                    o += put_copy<WRITE>(pos - cand, m, dst + o);
                    pos += m;
                    lit = pos;
                } else
                    ++pos;
            }
            o += put_literal<WRITE>(in + lit, len - lit, dst + o);
            if (!WRITE) sizes[li] = o;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(WK_WARPS * 32) k_weights_decode(const uint8_t *in, const int64_t *in_off, int64_t nlists, double *outd,
                                                                  const int64_t *out_off, int *err)
{
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int64_t li = (int64_t)blockIdx.x * WK_WARPS + wib; li < nlists; li += (int64_t)gridDim.x * WK_WARPS) {
        if (lane != 0) continue;
        const uint8_t *src = in + in_off[li];
        const int64_t inlen = in_off[li + 1] - in_off[li];
        uint8_t *out = reinterpret_cast<uint8_t *>(outd + out_off[li]);
        const int64_t want = 8 * (out_off[li + 1] - out_off[li]);
        int64_t p = 0, o = 0;
        uint64_t total = 0;
        int shift = 0;
        bool ok = true;
        for (;;) {
            if (p >= inlen || shift > 35) { ok = false; break; }
            const uint8_t b = src[p++];
            total |= (uint64_t)(b & 0x7f) << shift;
            shift += 7;
            if (!(b & 0x80)) break;
        }
        if (ok && (int64_t)total != want) ok = false; // the document says how many doubles the blob holds (uncompressedDoubleSize)
        while (ok && p < inlen) {
            const uint8_t tag = src[p++];
            int64_t len, offset = 0;
            if ((tag & 3) == 0) {
                len = (tag >> 2) + 1;
                if (len > 60) {
                    const int k = (int)len - 60;
                    if (p + k > inlen) { ok = false; break; }
                    uint32_t n = 0;
                    for (int i = 0; i < k; ++i) n |= (uint32_t)src[p + i] << (8 * i);
                    p += k;
                    len = (int64_t)n + 1;
                }
                if (p + len > inlen || o + len > want) { ok = false; break; }
                for (int64_t i = 0; i < len; ++i) out[o + i] = src[p + i];
                p += len;
                o += len;
                continue;
            }
            if ((tag & 3) == 1) {
                if (p + 1 > inlen) { ok = false; break; }
                len = ((tag >> 2) & 7) + 4;
                offset = ((int64_t)(tag >> 5) << 8) | src[p];
                p += 1;
            } else if ((tag & 3) == 2) {
                if (p + 2 > inlen) { ok = false; break; }
                len = (tag >> 2) + 1;
                offset = src[p] | ((int64_t)src[p + 1] << 8);
                p += 2;
            } else {
                if (p + 4 > inlen) { ok = false; break; }
                len = (tag >> 2) + 1;
                offset = (int64_t)src[p] | ((int64_t)src[p + 1] << 8) | ((int64_t)src[p + 2] << 16) | ((int64_t)src[p + 3] << 24);
                p += 4;
            }
            if (offset <= 0 || offset > o || o + len > want) { ok = false; break; }
            for (int64_t i = 0; i < len; ++i) out[o + i] = out[o - offset + i]; // overlapping copies repeat the pattern
            o += len;
        }
        if (!ok || o != want) atomicExch(err, 1);
    }
}

int grid_for(mcp_ctx *ctx, int64_t nlists)
{
    const int64_t want = (nlists + WK_WARPS - 1) / WK_WARPS, cap = (int64_t)ctx->prop.multiProcessorCount * 12;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

} // namespace

// sizes -> exclusive scan -> d_out_off[0..nlists] (bytes); *total = bytes needed
int weights_encode_plan(mcp_ctx *ctx, const double *d_w, const int64_t *d_off, int64_t nlists, int64_t *d_out_off, int64_t *total)
{
    MCP_CUDA(ctx, ctx->dSizes.reserve((size_t)(nlists + 1) * sizeof(int64_t)));
    {
        KScope ks(ctx, MCP_K_ENCODE);
        k_weights_encode<false><<<grid_for(ctx, nlists), WK_WARPS * 32, 0, ctx->stream>>>(d_w, d_off, nlists, nullptr, nullptr, ctx->dSizes.as<int64_t>());
        MCP_TRY(check_launch(ctx, "k_weights_encode<size>"));
    }
    MCP_TRY(exclusive_scan_i64(ctx, ctx->dSizes.as<int64_t>(), nlists, d_out_off));
    MCP_CUDA(ctx, cudaMemcpyAsync(ctx->hPinned, d_out_off + nlists, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *total = *reinterpret_cast<int64_t *>(ctx->hPinned);
    return MCP_OK;
}

int weights_encode_emit(mcp_ctx *ctx, const double *d_w, const int64_t *d_off, int64_t nlists, const int64_t *d_out_off, uint8_t *d_out)
{
    KScope ks(ctx, MCP_K_ENCODE);
    k_weights_encode<true><<<grid_for(ctx, nlists), WK_WARPS * 32, 0, ctx->stream>>>(d_w, d_off, nlists, d_out_off, d_out, nullptr);
    return check_launch(ctx, "k_weights_encode<emit>");
}

int weights_decode_device(mcp_ctx *ctx, const uint8_t *d_in, const int64_t *d_in_off, int64_t nlists, double *d_out, const int64_t *d_out_off)
{
    MCP_CUDA(ctx, ctx->dErr.reserve(64));
    MCP_CUDA(ctx, cudaMemsetAsync(ctx->dErr.p, 0, sizeof(int), ctx->stream));
    {
        KScope ks(ctx, MCP_K_DECODE);
        k_weights_decode<<<grid_for(ctx, nlists), WK_WARPS * 32, 0, ctx->stream>>>(d_in, d_in_off, nlists, d_out, d_out_off, ctx->dErr.as<int>());
        MCP_TRY(check_launch(ctx, "k_weights_decode"));
    }
    MCP_CUDA(ctx, cudaMemcpyAsync(ctx->hPinned, ctx->dErr.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*reinterpret_cast<int *>(ctx->hPinned) != 0) return fail(ctx, MCP_E_FORMAT, "weights decode: malformed Snappy stream");
    return MCP_OK;
}

} // namespace mcp
