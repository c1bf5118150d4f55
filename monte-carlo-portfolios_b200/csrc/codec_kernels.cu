// codec_kernels.cu -- batched encode/decode of integer lists in the JavaFastPFOR wire
// formats (kernel (4) of the hot path).  One warp per list, CSR in / CSR out.
//
// Encode is size pass -> exclusive scan -> emit pass; both passes instantiate the SAME
// templated routine (WRITE = false / true) so the sizes can never disagree with what is
// written.  Reference behaviour being reproduced (paths relative to
// /root/reference/JavaFastPFOR/src/main/java/me/lemire/integercompression/):
//   Composition.java:37-62, BinaryPacking.java:43-133, VariableByte.java:38-77,115-138,
//   FastPFOR.java:92-327, FastPFOR128.java:33, differential/Delta.java:24-28,84-88,
//   differential/IntegratedBinaryPacking.java:53-170, differential/IntegratedVariableByte.java,
//   differential/IntegratedComposition.java:41-66, differential/SkippableIntegratedComposition.java:46-71,
//   IntCompressor.java:38-61, differential/IntegratedIntCompressor.java:38-62.
#include "codec_device.cuh"
#include "common.cuh"

namespace mcp {

constexpr int kEncWarps = 8;  // warps per CTA
constexpr int kFpPage = 65536; // FastPFOR.DEFAULT_PAGE_SIZE (FastPFOR.java:46)

struct CodecTraits {
    bool fastpfor;   // FastPFOR / FastPFOR128 first stage
    int fp_block;    // 256 / 128
    bool blocks;     // has a BinaryPacking first stage
    bool integrated; // differential coding inside the codec
    bool sk;         // "skippable" framing: word 0 = n, no Composition marker
    bool chain_tail; // VByte tail continues the delta chain (Skippable) vs restarts at 0
};

__host__ __device__ inline CodecTraits codec_traits(int codec)
{
    CodecTraits t{false, 0, false, false, false, false};
    switch (codec) {
    case MCP_CODEC_BP32_VB: t.blocks = true; break;
    case MCP_CODEC_FASTPFOR256_VB: t.fastpfor = true; t.fp_block = 256; break;
    case MCP_CODEC_FASTPFOR128_VB: t.fastpfor = true; t.fp_block = 128; break;
    case MCP_CODEC_IBP32_IVB: t.blocks = true; t.integrated = true; break;
    case MCP_CODEC_SK_IBP32_IVB: t.blocks = true; t.integrated = true; t.sk = true; t.chain_tail = true; break;
    case MCP_CODEC_VB: break;
    case MCP_CODEC_SK_BP32_VB: t.blocks = true; t.sk = true; break;
    }
    return t;
}

int64_t codec_max_words(int codec, int64_t n)
{
    // worst case: 32 bits/int packed + one header per 32-int block, or 5 VByte bytes per tail int
    // (tail < 256 ints), plus FastPFOR page/bitmap/exception-count words.  Generous on purpose;
    // the reference sizes its scratch 4n+1024 (LoadPortfolios.java:611).
    (void)codec;
    return n + n / 16 + (n / kFpPage + 1) * 40 + 400;
}

struct EncArgs {
    int codec;
    bool delta, app;
    const int32_t *vals;
    const int64_t *list_off; // may be null -> uniform_len
    int64_t uniform_len;
    int64_t nlists;
    const int64_t *out_off;  // bytes (WRITE)
    uint8_t *out;
    int64_t *sizes;          // bytes (!WRITE)
    int *flags;              // [0] error, [1] number of multi-page FastPFOR lists (size pass), [2] shadow ticket (emit)
    uint32_t *shadow;        // emit: per multi-page list, 33 x 65536 words emulating FastPFOR.dataTobePacked
};
constexpr int64_t kShadowWords = 33 * (int64_t)kFpPage;

// x(i): the value sequence the codec sees (after the optional Delta.delta pre-pass)
struct ValSeq {
    const uint32_t *in;
    bool delta;
    __device__ __forceinline__ uint32_t operator()(int64_t i) const
    {
        const uint32_t v = __ldg(in + i);
        return (delta && i > 0) ? v - __ldg(in + i - 1) : v;
    }
};

// ------------------------------------------------------- BinaryPacking family
// Emits the 32-int blocks [0,m) of a list: BinaryPacking.headlessCompress
// (BinaryPacking.java:54-89) or, INTEGRATED, IntegratedBinaryPacking.headlessCompress
// (IntegratedBinaryPacking.java:80-125).  Returns words produced.
template <bool WRITE>
__device__ __forceinline__ int64_t warp_bp32_encode(const ValSeq &x, int64_t m, bool integrated, const WordSink &sink,
                                                    int64_t pos, int lane)
{
    const int64_t pos0 = pos;
    int64_t s = 0;
    auto load_block = [&](int64_t base, uint32_t &raw, uint32_t &d) {
        raw = x(base + lane);
        d = raw;
        if (integrated) {
            uint32_t prev = __shfl_up_sync(kFull, raw, 1);
            if (lane == 0) prev = base > 0 ? x(base - 1) : 0u;
            d = raw - prev;
        }
    };
    for (; s + 128 <= m; s += 128) {
        uint32_t raw[4], d[4];
        int b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            load_block(s + 32 * j, raw[j], d[j]);
            b[j] = bits32(__reduce_or_sync(kFull, d[j]));
        }
        if (WRITE && lane == 0)
            sink.put(pos, ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | (uint32_t)b[3]);
        ++pos;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (WRITE) {
                // IntegratedBitPacking b == 32 stores the RAW values (IntegratedBitPacking.java:1446-1451)
                const uint32_t w = warp_pack_word(b[j] == 32 ? raw[j] : d[j], b[j], lane);
                if (lane < b[j]) sink.put(pos + lane, w);
            }
            pos += b[j];
        }
    }
    for (; s < m; s += 32) {
        uint32_t raw, d;
        load_block(s, raw, d);
        const int b = bits32(__reduce_or_sync(kFull, d));
        if (WRITE) {
            if (lane == 0) sink.put(pos, (uint32_t)b);
            const uint32_t w = warp_pack_word(b == 32 ? raw : d, b, lane);
            if (lane < b) sink.put(pos + 1 + lane, w);
        }
        pos += 1 + b;
    }
    return pos - pos0;
}

// Decodes blocks [0,m): returns words consumed, or -1 on a malformed width.
__device__ __forceinline__ int64_t warp_bp32_decode(const WordSrc &src, int64_t p, int64_t avail, int64_t m,
                                                    bool integrated, uint32_t *out, uint32_t &carry, int lane)
{
    const int64_t p0 = p;
    int64_t s = 0;
    auto one = [&](int b, int64_t base) -> bool {
        if (b > 32 || p + b > avail) return false;
        const uint32_t word = (lane < b) ? src.get(p + lane) : 0u;
        uint32_t v = warp_unpack_val(word, b, lane);
        if (integrated && b != 32) v = warp_incl_scan(v, lane) + carry;
        out[base + lane] = v;
        if (integrated) carry = __shfl_sync(kFull, v, 31);
        p += b;
        return true;
    };
    for (; s + 128 <= m; s += 128) {
        if (p >= avail) return -1;
        const uint32_t h = src.get(p++);
        if (!one((int)(h >> 24), s)) return -1;
        if (!one((int)((h >> 16) & 255u), s + 32)) return -1;
        if (!one((int)((h >> 8) & 255u), s + 64)) return -1;
        if (!one((int)(h & 255u), s + 96)) return -1;
    }
    for (; s < m; s += 32) {
        if (p >= avail) return -1;
        const uint32_t h = src.get(p++);
        if (h > 32u) return -1;
        if (!one((int)h, s)) return -1;
    }
    return p - p0;
}

// ------------------------------------------------------------------- FastPFOR
struct FpBlockPlan { int b, c, maxb; };

// getBestBFromData (FastPFOR.java:107-134) for one block; v[r] = value 32*r+lane.
// `freq` = this warp's 33-int shared histogram.
template <int ROUNDS>
__device__ __forceinline__ FpBlockPlan warp_fp_plan(const uint32_t (&v)[ROUNDS], int *freq, int lane)
{
    constexpr int B = ROUNDS * 32;
    freq[lane] = 0;
    if (lane == 0) freq[32] = 0;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const int nb = bits32(v[r]);
        const unsigned grp = __match_any_sync(kFull, nb);
        if (lane == __ffs(grp) - 1) freq[nb] += __popc(grp);
        __syncwarp();
    }
    // lane l holds f = freqs[l+1]; c_l = number of values wider than l bits = sum_{j>l} freqs[j]
    const int f = freq[lane + 1];
    int c = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_down_sync(kFull, c, d);
        if (lane + d < 32) c += t;
    }
    const unsigned used = __ballot_sync(kFull, f > 0);
    const int maxb = used ? 32 - __clz(used) : 0;
    // candidate b = lane (< maxb, c < B); baseline b = maxb with cost maxb*B.  The Java loop
    // walks b downwards and accepts only strictly smaller costs, so on ties the larger b wins.
    int key = (maxb * B) * 64 + (32 - maxb);
    if (lane < maxb && c < B) {
        int cost = c * 8 + c * (maxb - lane) + lane * B + 8;
        if (maxb - lane == 1) cost -= c;
        const int k2 = cost * 64 + (32 - lane);
        key = min(key, k2);
    }
    key = __reduce_min_sync(kFull, key);
    const int bestb = 32 - (key & 63);
    const int bestc = (bestb == maxb) ? 0 : __shfl_sync(kFull, c, bestb & 31);
    __syncwarp();
    return FpBlockPlan{bestb, bestc, maxb};
}

// One FastPFOR page (encodePage, FastPFOR.java:136-212) of `n` ints (multiple of B) starting
// at x(base).  Returns words produced.  Fresh-instance semantics: exception buffers are
// zero, so the padding of the last packed group of every exception stream is zero.
template <bool WRITE, int ROUNDS>
__device__ int64_t warp_fp_encode_page(const ValSeq &x, int64_t base, int n, uint32_t *outw, int *freq, int *cnt_k,
                                       uint32_t *shadow, int lane)
{
    constexpr int B = ROUNDS * 32;
    const int nblocks = n / B;
    // ---- pass A: sizes
    int packed_words = 0, meta_bytes = 0;
    if (lane == 0) cnt_k[32] = 0;
    cnt_k[lane] = 0; // cnt_k[k], k = 0..32 (index 0 unused)
    __syncwarp();
    for (int j = 0; j < nblocks; ++j) {
        uint32_t v[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) v[r] = x(base + (int64_t)j * B + 32 * r + lane);
        const FpBlockPlan pl = warp_fp_plan<ROUNDS>(v, freq, lane);
        packed_words += ROUNDS * pl.b;
        meta_bytes += 2 + (pl.c > 0 ? 1 + pl.c : 0);
        if (pl.c > 0 && lane == 0) cnt_k[pl.maxb - pl.b] += pl.c;
        __syncwarp();
    }
    const int meta_words = (meta_bytes + 3) >> 2;
    // exception stream k (k = 2..32) takes 1 + ceil(cnt*k/32) words when non-empty; lane l handles k = l+1
    const int myk = lane + 1;
    const int mycnt = (myk >= 2) ? cnt_k[myk] : 0;
    const int mywords = mycnt ? 1 + (int)(((int64_t)mycnt * myk + 31) >> 5) : 0;
    const uint32_t incl = warp_incl_scan((uint32_t)mywords, lane);
    const int exc_words = (int)__shfl_sync(kFull, incl, 31);
    const int total = 1 + packed_words + 1 + meta_words + 1 + exc_words;
    if (!WRITE) return total;

    // ---- pass B: emit
    const int where_meta = 1 + packed_words;
    if (lane == 0) outw[0] = (uint32_t)where_meta;               // FastPFOR.java:182
    for (int j = where_meta + lane; j < total; j += 32) outw[j] = 0u; // meta padding + atomicOr targets
    __syncwarp();
    const unsigned bitmap_bits = __ballot_sync(kFull, mycnt != 0); // bit l <=> stream k=l+1 non-empty == bit (k-1)
    const int exc_base = where_meta + 1 + meta_words + 1;          // first word after the bitmap
    // stream_base[k]: word offset of stream k's count word
    __shared__ int s_stream_base[kEncWarps][33];
    __shared__ int s_stream_pos[kEncWarps][33];
    const int wib = (threadIdx.x >> 5);
    s_stream_base[wib][myk] = exc_base + (int)incl - mywords;
    s_stream_pos[wib][myk] = 0;
    if (lane == 0) {
        outw[where_meta] = (uint32_t)meta_bytes;                   // :186
        outw[where_meta + 1 + meta_words] = bitmap_bits;           // :191-196
    }
    if (mycnt) outw[exc_base + (int)incl - mywords] = (uint32_t)mycnt; // :200
    __syncwarp();
    uint8_t *meta = reinterpret_cast<uint8_t *>(outw + where_meta + 1);
    int mpos = 0, ppos = 1;
    for (int j = 0; j < nblocks; ++j) {
        uint32_t v[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) v[r] = x(base + (int64_t)j * B + 32 * r + lane);
        const FpBlockPlan pl = warp_fp_plan<ROUNDS>(v, freq, lane);
        const int b = pl.b;
        if (lane == 0) {
            meta[mpos] = (uint8_t)b;
            meta[mpos + 1] = (uint8_t)pl.c;
            if (pl.c > 0) meta[mpos + 2] = (uint8_t)pl.maxb;
        }
        if (pl.c > 0) {
            const int k = pl.maxb - b;
            const int sbase = s_stream_base[wib][k] + 1;
            int spos = s_stream_pos[wib][k];
            int e0 = 0;
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
                const uint32_t hi = v[r] >> b; // b <= 31 whenever c > 0
                const unsigned em = __ballot_sync(kFull, hi != 0);
                if (hi != 0) {
                    const int rank = e0 + __popc(em & ((1u << lane) - 1u));
                    meta[mpos + 3 + rank] = (uint8_t)(32 * r + lane); // :166
                    if (shadow) shadow[(int64_t)k * kFpPage + spos + rank] = hi; // dataTobePacked[k][dataPointers[k]++] (:169-170)
                    if (k >= 2) { // stream 1 is implicit (decoder ORs 1<<b, :291-295)
                        const int64_t bit = (int64_t)(spos + rank) * k;
                        const int w = (int)(bit >> 5), sh = (int)(bit & 31);
                        atomicOr(outw + sbase + w, hi << sh);
                        if (sh + k > 32) atomicOr(outw + sbase + w + 1, hi >> (32 - sh));
                    }
                }
                e0 += __popc(em);
            }
            __syncwarp();
            if (lane == 0) s_stream_pos[wib][k] = spos + pl.c;
            mpos += 3 + pl.c;
        } else {
            mpos += 2;
        }
        const uint32_t mask = b >= 32 ? 0xffffffffu : ((1u << b) - 1u);
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) { // BitPacking.fastpack = masked (:175-179)
            const uint32_t w = warp_pack_word(v[r] & mask, b, lane);
            if (lane < b) outw[ppos + lane] = w;
            ppos += b;
        }
        __syncwarp();
    }
    if (shadow && mycnt) {
        // FastPFOR never clears dataTobePacked (only the fill pointers, :143): the last packed group of stream k is
        // padded with whatever earlier pages of this codec instance left at positions [cnt, roundup32(cnt)), and the
        // words that survive the trim (:207-208) keep those bits.  Lane l replays that for stream k = l+1.
        const int nwords = (int)(((int64_t)mycnt * myk + 31) >> 5);
        uint32_t *sw = outw + s_stream_base[wib][myk] + 1;
        const int cend = (mycnt + 31) & ~31;
        for (int i = mycnt; i < cend; ++i) {
            uint32_t v = shadow[(int64_t)myk * kFpPage + i];
            if (myk < 32) v &= (1u << myk) - 1u;
            const int64_t bit = (int64_t)i * myk;
            const int w = (int)(bit >> 5), sh = (int)(bit & 31);
            if (w < nwords) sw[w] |= v << sh;
            if (sh + myk > 32 && w + 1 < nwords) sw[w + 1] |= v >> (32 - sh);
        }
    }
    __syncwarp();
    return total;
}

template <bool WRITE, int ROUNDS>
__device__ int64_t warp_fastpfor_encode(const ValSeq &x, int64_t m, uint32_t *outw, int *freq, int *cnt_k,
                                        uint32_t *shadow, int lane)
{
    // FastPFOR.headlessCompress (:92-105): pages of at most 65536 ints
    int64_t pos = 0;
    for (int64_t s = 0; s < m; s += kFpPage) {
        const int sz = (int)((m - s < kFpPage) ? m - s : kFpPage);
        pos += warp_fp_encode_page<WRITE, ROUNDS>(x, s, sz, WRITE ? outw + pos : nullptr, freq, cnt_k, shadow, lane);
    }
    return pos;
}

// decodePage (FastPFOR.java:233-307).  Returns words consumed or -1.
template <int ROUNDS>
__device__ int64_t warp_fp_decode_page(const uint32_t *inw, int64_t avail, int n, uint32_t *out, int *stream_base,
                                       int *stream_pos, int lane)
{
    constexpr int B = ROUNDS * 32;
    if (avail < 3) return -1;
    const int64_t where_meta = __ldg(inw);
    if (where_meta < 1 || where_meta >= avail) return -1;
    const int64_t bytesize = __ldg(inw + where_meta);
    const int64_t meta_words = (bytesize + 3) >> 2;
    int64_t e = where_meta + 1 + meta_words;
    if (bytesize < 0 || e >= avail) return -1;
    const uint32_t bitmap = __ldg(inw + e);
    ++e;
    // stream table (serial over set bits: at most 31)
    if (lane == 0) {
        int64_t q = e;
        bool ok = true;
        for (int k = 2; k <= 32; ++k) {
            stream_pos[k] = 0;
            if (bitmap & (1u << (k - 1))) {
                if (q >= avail) { ok = false; break; }
                const int64_t size = __ldg(inw + q);
                stream_base[k] = (int)(q + 1);
                q += 1 + ((size * k + 31) >> 5);
            } else
                stream_base[k] = -1;
        }
        stream_base[0] = ok && q <= avail ? (int)q : -1; // page end
        stream_pos[1] = 0;
    }
    __syncwarp();
    const int page_end = stream_base[0];
    if (page_end < 0) return -1;
    const uint8_t *meta = reinterpret_cast<const uint8_t *>(inw + where_meta + 1);
    int64_t p = 1;
    int mpos = 0;
    const int nblocks = n / B;
    for (int j = 0; j < nblocks; ++j) {
        if (mpos + 2 > bytesize) return -1;
        const int b = meta[mpos], c = meta[mpos + 1];
        if (b > 32 || p + (int64_t)ROUNDS * b > where_meta) return -1;
        uint32_t *o = out + (int64_t)j * B;
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const uint32_t word = (lane < b) ? __ldg(inw + p + lane) : 0u;
            o[32 * r + lane] = warp_unpack_val(word, b, lane);
            p += b;
        }
        if (c > 0) {
            if (mpos + 3 + c > bytesize) return -1;
            const int maxb = meta[mpos + 2];
            const int k = maxb - b;
            if (k < 1 || k > 32 || (k >= 2 && stream_base[k] < 0)) return -1;
            __syncwarp();
            const int spos = stream_pos[k];
            for (int ex = lane; ex < c; ex += 32) {
                const int posn = meta[mpos + 3 + ex];
                if (posn >= B) continue;
                uint32_t hi = 1u;
                if (k >= 2) {
                    const uint32_t *sw = inw + stream_base[k];
                    const int64_t bit = (int64_t)(spos + ex) * k;
                    const int64_t w = bit >> 5;
                    const int sh = (int)(bit & 31);
                    const uint32_t lo = __ldg(sw + w);
                    const uint32_t hw = (sh + k > 32) ? __ldg(sw + w + 1) : 0u;
                    hi = __funnelshift_r(lo, hw, sh);
                    if (k < 32) hi &= (1u << k) - 1u;
                }
                o[posn] |= hi << b;
            }
            __syncwarp();
            if (lane == 0) stream_pos[k] = spos + c;
            mpos += 3 + c;
        } else
            mpos += 2;
        __syncwarp();
    }
    return page_end;
}

// ------------------------------------------------------------------- kernels
template <bool WRITE>
__global__ void __launch_bounds__(kEncWarps * 32) k_encode(const EncArgs a)
{
    __shared__ uint32_t s_stage[kEncWarps][48];
    __shared__ int s_freq[kEncWarps][34];
    __shared__ int s_cnt[kEncWarps][34];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * kEncWarps;
    const CodecTraits tr = codec_traits(a.codec);
    for (int64_t li = (int64_t)blockIdx.x * kEncWarps + wib; li < a.nlists; li += nwarps) {
        const int64_t beg = a.list_off ? a.list_off[li] : li * a.uniform_len;
        const int64_t n = a.list_off ? a.list_off[li + 1] - beg : a.uniform_len;
        const ValSeq x{reinterpret_cast<const uint32_t *>(a.vals) + beg, a.delta};
        uint32_t *outw = WRITE ? reinterpret_cast<uint32_t *>(a.out + a.out_off[li]) : nullptr;
        const WordSink sink{outw, a.app};
        int64_t pos = 0;
        if (tr.sk) { // IntCompressor / IntegratedIntCompressor: compressed[0] = input.length
            if (WRITE && lane == 0) sink.put(0, (uint32_t)n);
            pos = 1;
        }
        if (n > 0) {
            int64_t m = 0;
            if (tr.fastpfor) {
                m = n - n % tr.fp_block;
                if (WRITE && lane == 0) sink.put(pos, (uint32_t)m); // header, or Composition's 0 marker
                ++pos;
                if (m > 0) {
                    // FastPFOR writes native words (byte-granular metadata); APP framing swaps afterwards
                    int64_t w;
                    uint32_t *shadow = nullptr;
                    if (m > kFpPage) { // multi-page list: needs the stale-buffer emulation
                        if (WRITE) {
                            int slot = 0;
                            if (lane == 0) slot = atomicAdd(a.flags + 2, 1);
                            slot = __shfl_sync(kFull, slot, 0);
                            shadow = a.shadow + (int64_t)slot * kShadowWords;
                        } else if (lane == 0)
                            atomicAdd(a.flags + 1, 1);
                    }
                    if (tr.fp_block == 256)
                        w = warp_fastpfor_encode<WRITE, 8>(x, m, WRITE ? outw + pos : nullptr, s_freq[wib], s_cnt[wib], shadow, lane);
                    else
                        w = warp_fastpfor_encode<WRITE, 4>(x, m, WRITE ? outw + pos : nullptr, s_freq[wib], s_cnt[wib], shadow, lane);
                    if (WRITE && a.app) {
                        __syncwarp();
                        for (int64_t j = pos + lane; j < pos + w; j += 32) outw[j] = bswap32(outw[j]);
                    }
                    pos += w;
                }
            } else if (tr.blocks) {
                m = n & ~(int64_t)31;
                if (!tr.sk) { // Composition: F1 header, or the 0 marker when F1 wrote nothing (Composition.java:45-48)
                    if (WRITE && lane == 0) sink.put(pos, (uint32_t)m);
                    ++pos;
                }
                pos += warp_bp32_encode<WRITE>(x, m, tr.integrated, sink, pos, lane);
            }
            const int64_t r = n - m;
            if (r > 0) {
                if (tr.integrated) {
                    // IntegratedVariableByte: gaps; the non-skippable composition restarts at 0
                    // (IntegratedVariableByte.java:40), the skippable one chains (:188-189)
                    const bool chain = tr.chain_tail;
                    auto getv = [&](int64_t i) -> uint32_t {
                        const int64_t g = m + i;
                        const uint32_t prev = (i > 0 || (chain && g > 0)) ? x(g - 1) : 0u;
                        return x(g) - prev;
                    };
                    pos += warp_vb_encode<WRITE>(getv, r, sink, pos, s_stage[wib], lane);
                } else {
                    auto getv = [&](int64_t i) -> uint32_t { return x(m + i); };
                    pos += warp_vb_encode<WRITE>(getv, r, sink, pos, s_stage[wib], lane);
                }
            }
        }
        if (WRITE) {
            if (a.app && lane == 0) sink.put(pos, 0u); // LoadPortfolios.java:622: outpos + 1 ints
        } else if (lane == 0) {
            a.sizes[li] = 4 * (pos + (a.app ? 1 : 0));
        }
        __syncwarp();
    }
}

struct DecArgs {
    int codec;
    bool delta, app;
    const uint8_t *in;
    const int64_t *in_off;
    int64_t nlists;
    int32_t *out;
    const int64_t *out_off;
    int *err;
};

__global__ void __launch_bounds__(kEncWarps * 32) k_decode(const DecArgs a)
{
    __shared__ int s_sbase[kEncWarps][34];
    __shared__ int s_spos[kEncWarps][34];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * kEncWarps;
    const CodecTraits tr = codec_traits(a.codec);
    for (int64_t li = (int64_t)blockIdx.x * kEncWarps + wib; li < a.nlists; li += nwarps) {
        const int64_t ib = a.in_off[li], ie = a.in_off[li + 1];
        const int64_t n = a.out_off[li + 1] - a.out_off[li];
        uint32_t *out = reinterpret_cast<uint32_t *>(a.out) + a.out_off[li];
        const int64_t avail = (ie - ib) >> 2;
        const WordSrc src{reinterpret_cast<const uint32_t *>(a.in + ib), a.app};
        bool ok = ((ie - ib) & 3) == 0 && (ib & 3) == 0;
        int64_t p = 0;
        if (ok && tr.sk) {
            ok = avail >= 1 && (int64_t)src.get(0) == n;
            p = 1;
        }
        if (ok && n > 0) {
            int64_t m = 0;
            uint32_t carry = 0;
            if (tr.fastpfor || (tr.blocks && !tr.sk)) {
                ok = p < avail;
                if (ok) {
                    m = src.get(p++);
                    const int64_t blk = tr.fastpfor ? tr.fp_block : 32;
                    ok = (m == n - n % blk);
                }
            } else if (tr.blocks)
                m = n & ~(int64_t)31;
            if (ok && m > 0) {
                if (tr.fastpfor) {
                    // FastPFOR pages were written as native words; APP framing stores them swapped.
                    // Decode needs byte access to the metadata, so un-swap into the output tail is not
                    // possible; instead the host launcher pre-swaps APP-framed FastPFOR input (see
                    // codec_decode_device), so here words are always native.
                    for (int64_t s = 0; ok && s < m; s += kFpPage) {
                        const int sz = (int)((m - s < kFpPage) ? m - s : kFpPage);
                        int64_t used;
                        if (tr.fp_block == 256)
                            used = warp_fp_decode_page<8>(src.w + p, avail - p, sz, out + s, s_sbase[wib], s_spos[wib], lane);
                        else
                            used = warp_fp_decode_page<4>(src.w + p, avail - p, sz, out + s, s_sbase[wib], s_spos[wib], lane);
                        __syncwarp();
                        if (used < 0) ok = false; else p += used;
                    }
                } else {
                    const int64_t used = warp_bp32_decode(src, p, avail, m, tr.integrated, out, carry, lane);
                    if (used < 0) ok = false; else p += used;
                }
            }
            if (ok && n > m) {
                if (tr.integrated) ok = warp_vb_decode<true>(src, p, avail, n - m, out + m, tr.chain_tail ? carry : 0u, lane);
                else ok = warp_vb_decode<false>(src, p, avail, n - m, out + m, 0u, lane);
            }
            if (ok && a.delta) {
                __syncwarp();
                warp_inverse_delta(out, n, lane);
            }
        }
        if (!ok && lane == 0) atomicExch(a.err, 1);
        __syncwarp();
    }
}

// Number of integers a compressed list decodes to, from its headers alone (what Java's
// uncompress() discovers as it goes: Composition.java:54-62).  One warp per list.
__global__ void __launch_bounds__(kEncWarps * 32) k_decode_count(const DecArgs a, int64_t *counts)
{
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * kEncWarps;
    const CodecTraits tr = codec_traits(a.codec);
    for (int64_t li = (int64_t)blockIdx.x * kEncWarps + wib; li < a.nlists; li += nwarps) {
        const int64_t ib = a.in_off[li], ie = a.in_off[li + 1];
        const int64_t avail = (ie - ib) >> 2;
        const WordSrc src{reinterpret_cast<const uint32_t *>(a.in + ib), a.app};
        bool ok = ((ie - ib) & 3) == 0 && (ib & 3) == 0;
        int64_t n = 0;
        if (ok && tr.sk) {
            ok = avail >= 1;
            if (ok) n = src.get(0);
        } else if (ok && avail > 0) {
            int64_t p = 0, m = 0;
            if (tr.fastpfor || tr.blocks) {
                m = src.get(p++);
                const int64_t blk = tr.fastpfor ? tr.fp_block : 32;
                ok = m >= 0 && m % blk == 0;
                if (ok && tr.fastpfor) {
                    for (int64_t s = 0; ok && s < m; s += kFpPage) { // walk the page headers
                        if (p + 2 >= avail) { ok = false; break; }
                        const int64_t wm = src.get(p);
                        if (wm < 1 || p + wm >= avail) { ok = false; break; }
                        const int64_t bytesize = src.get(p + wm);
                        int64_t e = p + wm + 1 + ((bytesize + 3) >> 2);
                        if (bytesize < 0 || e >= avail) { ok = false; break; }
                        const uint32_t bitmap = src.get(e++);
                        for (int k = 2; k <= 32 && ok; ++k)
                            if (bitmap & (1u << (k - 1))) {
                                if (e >= avail) { ok = false; break; }
                                const int64_t size = src.get(e);
                                e += 1 + ((size * k + 31) >> 5);
                            }
                        if (e > avail) ok = false;
                        p = e;
                    }
                } else if (ok) {
                    int64_t s = 0;
                    for (; ok && s + 128 <= m; s += 128) {
                        if (p >= avail) { ok = false; break; }
                        const uint32_t h = src.get(p);
                        p += 1 + (h >> 24) + ((h >> 16) & 255u) + ((h >> 8) & 255u) + (h & 255u);
                    }
                    for (; ok && s < m; s += 32) {
                        if (p >= avail) { ok = false; break; }
                        const uint32_t h = src.get(p);
                        if (h > 32u) ok = false;
                        p += 1 + h;
                    }
                    if (p > avail) ok = false;
                }
            }
            if (ok) { // VByte tail: one value per terminator byte
                int64_t c = 0;
                for (int64_t j = p + lane; j < avail; j += 32) c += __popc((src.get(j) >> 7) & 0x01010101u);
                for (int d = 16; d; d >>= 1) c += __shfl_down_sync(kFull, c, d);
                c = __shfl_sync(kFull, c, 0);
                n = m + c;
            }
        }
        if (lane == 0) {
            counts[li] = ok ? n : 0;
            if (!ok) atomicExch(a.err, 1);
        }
    }
}

// in-place byte swap of 32-bit words (APP-framed FastPFOR input -> native)
__global__ void k_bswap_copy(const uint32_t *in, uint32_t *out, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = bswap32(in[i]);
}

// ---------------------------------------------------------------------- scan
// exclusive scan of int64: per-block partial sums, serial scan of the partials by one
// block, then a second sweep.  Sizes here are "number of lists": bookkeeping, not hot.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;

__global__ void k_scan_partials(const int64_t *in, int64_t n, int64_t *partials)
{
    __shared__ int64_t s[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanThreads * kScanItems;
    int64_t sum = 0;
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
        if (i < n) sum += in[i];
    }
    for (int d = 16; d; d >>= 1) sum += __shfl_down_sync(kFull, sum, d);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += s[w];
        partials[blockIdx.x] = t;
    }
}

__global__ void k_scan_partials_serial(int64_t *partials, int64_t nb, int64_t *total_out)
{
    // one warp, chunked
    const int lane = threadIdx.x;
    int64_t run = 0;
    for (int64_t base = 0; base < nb; base += 32) {
        const int64_t i = base + lane;
        int64_t v = i < nb ? partials[i] : 0;
        int64_t incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t t = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += t;
        }
        if (i < nb) partials[i] = run + incl - v;
        run += __shfl_sync(kFull, incl, 31);
    }
    if (lane == 0) *total_out = run;
}

__global__ void k_scan_apply(const int64_t *in, int64_t n, const int64_t *partials, int64_t *out)
{
    // block-local exclusive scan in the same item order as k_scan_partials: thread t owns items
    // {k*T + t}; to keep a simple order we re-map: thread t handles the contiguous run [t*I, (t+1)*I)
    __shared__ int64_t s[kScanThreads];
    const int64_t base = (int64_t)blockIdx.x * kScanThreads * kScanItems;
    int64_t v[kScanItems];
    int64_t sum = 0;
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + (int64_t)threadIdx.x * kScanItems + k;
        v[k] = i < n ? in[i] : 0;
        sum += v[k];
    }
    s[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < kScanThreads; d <<= 1) {
        const int64_t t = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    int64_t run = partials[blockIdx.x] + s[threadIdx.x] - sum;
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + (int64_t)threadIdx.x * kScanItems + k;
        if (i < n) out[i] = run;
        run += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) out[n] = run;
}

int exclusive_scan_i64(mcp_ctx *ctx, const int64_t *d_in, int64_t n, int64_t *d_out)
{
    if (n == 0) {
        MCP_CUDA(ctx, cudaMemsetAsync(d_out, 0, sizeof(int64_t), ctx->stream));
        return MCP_OK;
    }
    const int64_t per = (int64_t)kScanThreads * kScanItems;
    const int64_t nb = (n + per - 1) / per;
    MCP_CUDA(ctx, ctx->dScanTmp.reserve((size_t)(nb + 1) * sizeof(int64_t)));
    int64_t *part = ctx->dScanTmp.as<int64_t>();
    KScope ks(ctx, MCP_K_SCAN, 3);
    k_scan_partials<<<(unsigned)nb, kScanThreads, 0, ctx->stream>>>(d_in, n, part);
    k_scan_partials_serial<<<1, 32, 0, ctx->stream>>>(part, nb, part + nb);
    k_scan_apply<<<(unsigned)nb, kScanThreads, 0, ctx->stream>>>(d_in, n, part, d_out);
    return check_launch(ctx, "scan");
}

// ------------------------------------------------------------------ launchers
static int grid_for_lists(mcp_ctx *ctx, int64_t nlists)
{
    const int64_t want = (nlists + kEncWarps - 1) / kEncWarps;
    const int64_t cap = (int64_t)ctx->prop.multiProcessorCount * 8; // 8 resident CTAs of 8 warps per SM
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

static int make_enc_args(mcp_ctx *ctx, int codec_id, int framing, const int32_t *d_vals, const int64_t *d_list_off,
                         int64_t uniform_len, int64_t nlists, EncArgs &a)
{
    const int codec = codec_id & MCP_CODEC_MASK;
    if (codec < 0 || codec >= MCP_CODEC_COUNT || (codec_id & ~(MCP_CODEC_MASK | MCP_FLAG_DELTA))) return fail(ctx, MCP_E_ARG, "unknown codec id");
    if (framing != MCP_FRAME_WORDS && framing != MCP_FRAME_APP) return fail(ctx, MCP_E_ARG, "unknown framing");
    if (nlists < 0 || uniform_len < 0) return fail(ctx, MCP_E_ARG, "negative size");
    a = EncArgs{};
    a.codec = codec;
    a.delta = (codec_id & MCP_FLAG_DELTA) != 0;
    a.app = framing == MCP_FRAME_APP;
    a.vals = d_vals;
    a.list_off = d_list_off;
    a.uniform_len = uniform_len;
    a.nlists = nlists;
    return MCP_OK;
}

// size pass + exclusive scan: fills d_out_off[0..nlists] (bytes) and *total_out
int codec_encode_plan(mcp_ctx *ctx, int codec_id, int framing, const int32_t *d_vals, const int64_t *d_list_off,
                      int64_t uniform_len, int64_t nlists, int64_t *d_out_off, int64_t *total_out)
{
    EncArgs a;
    MCP_TRY(make_enc_args(ctx, codec_id, framing, d_vals, d_list_off, uniform_len, nlists, a));
    if (nlists == 0) {
        MCP_CUDA(ctx, cudaMemsetAsync(d_out_off, 0, sizeof(int64_t), ctx->stream));
        *total_out = 0;
        return MCP_OK;
    }
    MCP_CUDA(ctx, ctx->dSizes.reserve((size_t)nlists * sizeof(int64_t)));
    MCP_CUDA(ctx, ctx->dErr.reserve(64));
    MCP_CUDA(ctx, cudaMemsetAsync(ctx->dErr.p, 0, 16, ctx->stream));
    a.sizes = ctx->dSizes.as<int64_t>();
    a.flags = ctx->dErr.as<int>();
    {
        KScope ks(ctx, MCP_K_ENCODE);
        k_encode<false><<<grid_for_lists(ctx, nlists), kEncWarps * 32, 0, ctx->stream>>>(a);
        MCP_TRY(check_launch(ctx, "k_encode<size>"));
    }
    MCP_TRY(exclusive_scan_i64(ctx, a.sizes, nlists, d_out_off));
    MCP_CUDA(ctx, cudaMemcpyAsync(ctx->hPinned, d_out_off + nlists, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    MCP_CUDA(ctx, cudaMemcpyAsync(reinterpret_cast<char *>(ctx->hPinned) + 8, ctx->dErr.as<int>() + 1, sizeof(int),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *total_out = *reinterpret_cast<int64_t *>(ctx->hPinned);
    ctx->planMultiPage = *reinterpret_cast<int *>(reinterpret_cast<char *>(ctx->hPinned) + 8);
    return MCP_OK;
}

int codec_encode_emit(mcp_ctx *ctx, int codec_id, int framing, const int32_t *d_vals, const int64_t *d_list_off,
                      int64_t uniform_len, int64_t nlists, const int64_t *d_out_off, uint8_t *d_out)
{
    EncArgs a;
    MCP_TRY(make_enc_args(ctx, codec_id, framing, d_vals, d_list_off, uniform_len, nlists, a));
    if (nlists == 0) return MCP_OK;
    a.out_off = d_out_off;
    a.out = d_out;
    MCP_CUDA(ctx, ctx->dErr.reserve(64));
    MCP_CUDA(ctx, cudaMemsetAsync(ctx->dErr.p, 0, 16, ctx->stream));
    a.flags = ctx->dErr.as<int>();
    if (codec_traits(a.codec).fastpfor && ctx->planMultiPage > 0) {
        const size_t bytes = (size_t)ctx->planMultiPage * kShadowWords * sizeof(uint32_t);
        MCP_CUDA(ctx, ctx->dShadow.reserve(bytes));
        MCP_CUDA(ctx, cudaMemsetAsync(ctx->dShadow.p, 0, bytes, ctx->stream));
        a.shadow = ctx->dShadow.as<uint32_t>();
    }
    KScope ks(ctx, MCP_K_ENCODE);
    k_encode<true><<<grid_for_lists(ctx, nlists), kEncWarps * 32, 0, ctx->stream>>>(a);
    return check_launch(ctx, "k_encode<emit>");
}

int codec_decode_device(mcp_ctx *ctx, int codec_id, int framing, const uint8_t *d_in, const int64_t *d_in_off,
                        int64_t nlists, int32_t *d_out, const int64_t *d_out_off)
{
    const int codec = codec_id & MCP_CODEC_MASK;
    if (codec < 0 || codec >= MCP_CODEC_COUNT || (codec_id & ~(MCP_CODEC_MASK | MCP_FLAG_DELTA))) return fail(ctx, MCP_E_ARG, "unknown codec id");
    if (framing != MCP_FRAME_WORDS && framing != MCP_FRAME_APP) return fail(ctx, MCP_E_ARG, "unknown framing");
    if (nlists < 0) return fail(ctx, MCP_E_ARG, "negative size");
    if (nlists == 0) return MCP_OK;
    MCP_CUDA(ctx, ctx->dErr.reserve(sizeof(int)));
    MCP_CUDA(ctx, cudaMemsetAsync(ctx->dErr.p, 0, sizeof(int), ctx->stream));
    DecArgs a{};
    a.codec = codec;
    a.delta = (codec_id & MCP_FLAG_DELTA) != 0;
    a.app = framing == MCP_FRAME_APP;
    a.in = d_in;
    a.in_off = d_in_off;
    a.nlists = nlists;
    a.out = d_out;
    a.out_off = d_out_off;
    a.err = ctx->dErr.as<int>();
    const CodecTraits tr = codec_traits(codec);
    if (tr.fastpfor && a.app) {
        // FastPFOR metadata is byte-addressed: bring the big-endian stream to native order once
        int64_t total = 0;
        MCP_CUDA(ctx, cudaMemcpyAsync(ctx->hPinned, d_in_off + nlists, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        total = *reinterpret_cast<int64_t *>(ctx->hPinned);
        if (total & 3) return fail(ctx, MCP_E_FORMAT, "decode: byte length not a multiple of 4");
        MCP_CUDA(ctx, ctx->dSizes.reserve((size_t)total + 16));
        KScope ks(ctx, MCP_K_DECODE);
        k_bswap_copy<<<ctx->prop.multiProcessorCount * 4, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t *>(d_in),
                                                                              ctx->dSizes.as<uint32_t>(), total >> 2);
        MCP_TRY(check_launch(ctx, "k_bswap_copy"));
        a.in = ctx->dSizes.as<uint8_t>();
        a.app = false;
    }
    {
        KScope ks(ctx, MCP_K_DECODE);
        k_decode<<<grid_for_lists(ctx, nlists), kEncWarps * 32, 0, ctx->stream>>>(a);
        MCP_TRY(check_launch(ctx, "k_decode"));
    }
    MCP_CUDA(ctx, cudaMemcpyAsync(ctx->hPinned, ctx->dErr.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*reinterpret_cast<int *>(ctx->hPinned) != 0) return fail(ctx, MCP_E_FORMAT, "decode: malformed compressed stream");
    return MCP_OK;
}

// d_counts[i] = decoded length of list i (device).  APP-framed input is read big-endian.
int codec_decode_counts(mcp_ctx *ctx, int codec_id, int framing, const uint8_t *d_in, const int64_t *d_in_off,
                        int64_t nlists, int64_t *d_counts)
{
    const int codec = codec_id & MCP_CODEC_MASK;
    if (codec < 0 || codec >= MCP_CODEC_COUNT || (codec_id & ~(MCP_CODEC_MASK | MCP_FLAG_DELTA))) return fail(ctx, MCP_E_ARG, "unknown codec id");
    if (framing != MCP_FRAME_WORDS && framing != MCP_FRAME_APP) return fail(ctx, MCP_E_ARG, "unknown framing");
    if (nlists <= 0) return MCP_OK;
    MCP_CUDA(ctx, ctx->dErr.reserve(sizeof(int)));
    MCP_CUDA(ctx, cudaMemsetAsync(ctx->dErr.p, 0, sizeof(int), ctx->stream));
    DecArgs a{};
    a.codec = codec;
    a.app = framing == MCP_FRAME_APP;
    a.in = d_in;
    a.in_off = d_in_off;
    a.nlists = nlists;
    a.err = ctx->dErr.as<int>();
    {
        KScope ks(ctx, MCP_K_DECODE);
        k_decode_count<<<grid_for_lists(ctx, nlists), kEncWarps * 32, 0, ctx->stream>>>(a, d_counts);
        MCP_TRY(check_launch(ctx, "k_decode_count"));
    }
    MCP_CUDA(ctx, cudaMemcpyAsync(ctx->hPinned, ctx->dErr.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*reinterpret_cast<int *>(ctx->hPinned) != 0) return fail(ctx, MCP_E_FORMAT, "decode: malformed compressed stream");
    return MCP_OK;
}

} // namespace mcp
