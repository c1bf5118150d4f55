// codec_device.cuh -- warp-level building blocks of the JavaFastPFOR wire formats.
//
// One warp works on one integer list at a time; lane i owns value i of the current
// 32-int block, which maps 1:1 onto BinaryPacking.BLOCK_SIZE = 32
// (JavaFastPFOR/.../BinaryPacking.java:40).  All arithmetic is uint32 so Java's int
// wrap-around and `>>>` are reproduced bit for bit.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mcp {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int bits32(uint32_t v) { return 32 - __clz((int)v); } // Util.java:101-103

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Word sink: 32-bit words of one encoded list.  `swap` = MCP_FRAME_APP big-endian
// framing (BitManipulationHelper.java:59-64) applied at store time.
struct WordSink {
    uint32_t *w;
    bool swap;
    __device__ __forceinline__ void put(int64_t j, uint32_t x) const { w[j] = swap ? bswap32(x) : x; }
};

struct WordSrc {
    const uint32_t *w;
    bool swap;
    __device__ __forceinline__ uint32_t get(int64_t j) const
    {
        const uint32_t x = __ldg(w + j);
        return swap ? bswap32(x) : x;
    }
};

// ---------------------------------------------------------------- bit packing
// BitPacking layout (BitPacking.java:191-233): value i occupies bits [i*b,(i+1)*b)
// of a little-endian-by-word stream.  Each lane holds value `v` (already reduced to
// b bits unless b == 32); lane j < b returns output word j.
__device__ __forceinline__ uint32_t warp_pack_word(uint32_t v, int b, int lane)
{
    if (b == 32) return v;
    if (b == 0) return 0;
    const int wbit = lane << 5;         // first stream bit of this lane's word
    const int first = wbit / b;         // first value overlapping it
    const int T = 32 / b + 2;           // warp-uniform trip count
    uint32_t w = 0;
    for (int t = 0; t < T; ++t) {
        const int i = first + t;
        const uint32_t x = __shfl_sync(kFull, v, i & 31);
        const int off = i * b - wbit;   // > -b
        if (i < 32 && off < 32) w |= (off >= 0) ? (x << off) : (x >> (-off));
    }
    return w;
}

// Inverse (BitPacking.java:3021-3127): `word` = stream word `lane` (garbage for
// lane >= b is fine); returns value `lane`.
__device__ __forceinline__ uint32_t warp_unpack_val(uint32_t word, int b, int lane)
{
    if (b == 32) return word;
    const int bit = lane * b;
    const int w = bit >> 5, s = bit & 31;
    const uint32_t lo = __shfl_sync(kFull, word, w);
    const uint32_t hi = __shfl_sync(kFull, word, (w + 1) & 31);
    if (b == 0) return 0;
    return __funnelshift_r(lo, hi, s) & ((1u << b) - 1u);
}

// ---------------------------------------------------------------- VariableByte
// VariableByte.java:38-77: 7-bit groups, low group first, bit 7 set on the LAST byte
// of a value; stream zero-padded to 4 bytes and stored as little-endian words.
__device__ __forceinline__ int vb_len(uint32_t v)
{
    // ceil(bits(v) / 7), at least 1: (bits + 6) / 7 with the division as a multiply-shift (exact for 6..38)
    const int l = ((38 - __clz((int)v)) * 37) >> 8;
    return l > 0 ? l : 1;
}

// Encodes `count` values produced by getv(i), i in [0,count).  `stage` is this warp's
// 48-word shared scratch.  Returns the number of words (written to sink at word
// `pos0` onwards when WRITE).
template <bool WRITE, class GetV>
__device__ __forceinline__ int64_t warp_vb_encode(GetV getv, int64_t count, const WordSink &sink, int64_t pos0,
                                                  uint32_t *stage, int lane)
{
    int64_t wpos = pos0;
    int64_t total_bytes = 0;
    int carry = 0;
    uint8_t *sb = reinterpret_cast<uint8_t *>(stage);
    for (int64_t base = 0; base < count; base += 32) {
        const int64_t i = base + lane;
        const bool act = i < count;
        const uint32_t v = act ? getv(i) : 0u;
        const int nb = act ? vb_len(v) : 0;
        const uint32_t incl = warp_incl_scan((uint32_t)nb, lane);
        const int total = (int)__shfl_sync(kFull, incl, 31);
        if (WRITE) {
            int o = carry + (int)incl - nb;
            uint32_t x = v;
            for (int k = 0; k < nb - 1; ++k) { sb[o++] = (uint8_t)(x & 127u); x >>= 7; }
            if (nb) sb[o] = (uint8_t)(x | 128u);
            __syncwarp();
            const int nbytes = carry + total;
            const int nfull = nbytes >> 2;
            for (int j = lane; j < nfull; j += 32) sink.put(wpos + j, stage[j]);
            const int rem = nbytes & 3;
            uint32_t keep = 0;
            if (lane == 0 && rem) keep = stage[nfull] & ((1u << (8 * rem)) - 1u);
            __syncwarp();
            if (lane == 0) stage[0] = keep;
            __syncwarp();
            wpos += nfull;
            carry = rem;
        } else {
            total_bytes += total;
        }
    }
    if (WRITE) {
        if (carry) {
            if (lane == 0) sink.put(wpos, stage[0]); // already zero-padded by the mask above
            ++wpos;
        }
        return wpos - pos0;
    }
    return (total_bytes + 3) >> 2;
}

// Decodes exactly `count` values from src words [p0, avail) (VariableByte.java:115-138
// / :180-203).  With INTEGRATED the values are running sums starting at `init`
// (IntegratedVariableByte.java:114-141, 227-254).  Returns false on a truncated stream.
struct PlainOut {
    uint32_t *out;
    __device__ __forceinline__ void operator()(int64_t i, uint32_t v) const { out[i] = v; }
};

template <bool INTEGRATED, class Src, class Out>
__device__ __forceinline__ bool warp_vb_decode_to(const Src &src, int64_t p0, int64_t avail, int64_t count, const Out &put,
                                                  uint32_t init, int lane)
{
    int64_t decoded = 0;
    uint32_t carry_w = 0x80808080u; // "previous word ends on terminators": a value cannot start before the stream
    uint32_t run = init;
    int64_t p = p0;
    while (decoded < count) {
        if (p >= avail) return false;
        const int64_t j = p + lane;
        const uint32_t w = (j < avail) ? src.get(j) : 0u;
        uint32_t pw = __shfl_up_sync(kFull, w, 1);
        if (lane == 0) pw = carry_w;
        const uint64_t W = ((uint64_t)w << 32) | pw;
        const uint32_t term = (w >> 7) & 0x01010101u;
        const int cnt = __popc(term);
        const uint32_t incl = warp_incl_scan((uint32_t)cnt, lane);
        const int tot = (int)__shfl_sync(kFull, incl, 31);
        uint32_t vals[4];
        int nv = 0;
        uint32_t lane_sum = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (term & (1u << (8 * e))) {
                const int q = 4 + e;
                uint32_t v = (uint32_t)(W >> (8 * q)) & 127u;
#pragma unroll
                for (int k = 1; k <= 4; ++k) {
                    const uint32_t bb = (uint32_t)(W >> (8 * (q - k))) & 255u;
                    if (bb & 128u) break;
                    v = (v << 7) | (bb & 127u);
                }
                vals[nv++] = v;
                lane_sum += v;
            }
        }
        uint32_t basev = 0;
        if (INTEGRATED) {
            const uint32_t sincl = warp_incl_scan(lane_sum, lane);
            basev = run + sincl - lane_sum;
            run += __shfl_sync(kFull, sincl, 31);
        }
        int64_t idx = decoded + (int64_t)incl - cnt;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (r < nv) {
                uint32_t v = vals[r];
                if (INTEGRATED) { basev += v; v = basev; }
                if (idx < count) put(idx, v);
                ++idx;
            }
        }
        decoded += tot;
        carry_w = __shfl_sync(kFull, w, 31);
        p += 32;
    }
    return true;
}

template <bool INTEGRATED>
__device__ __forceinline__ bool warp_vb_decode(const WordSrc &src, int64_t p0, int64_t avail, int64_t count,
                                               uint32_t *out, uint32_t init, int lane)
{
    return warp_vb_decode_to<INTEGRATED>(src, p0, avail, count, PlainOut{out}, init, lane);
}

// In-place inclusive prefix sum over out[0..n) by one warp (Delta.inverseDelta, Delta.java:84-88)
__device__ __forceinline__ void warp_inverse_delta(uint32_t *out, int64_t n, int lane)
{
    uint32_t run = 0;
    for (int64_t base = 0; base < n; base += 32) {
        const int64_t i = base + lane;
        const uint32_t v = (i < n) ? out[i] : 0u;
        const uint32_t incl = warp_incl_scan(v, lane) + run;
        if (i < n) out[i] = incl;
        run = __shfl_sync(kFull, incl, 31);
    }
}

} // namespace mcp
