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
    bool xorbp;      // XorBinaryPacking first stage (128-value blocks, XOR with the predecessor)
    bool optpfd;     // OptPFD first stage (128-value blocks, S16-coded exceptions)
};

__host__ __device__ inline CodecTraits codec_traits(int codec)
{
    CodecTraits t{false, 0, false, false, false, false, false, false};
    switch (codec) {
    case MCP_CODEC_BP32_VB: t.blocks = true; break;
    case MCP_CODEC_FASTPFOR256_VB: t.fastpfor = true; t.fp_block = 256; break;
    case MCP_CODEC_FASTPFOR128_VB: t.fastpfor = true; t.fp_block = 128; break;
    case MCP_CODEC_IBP32_IVB: t.blocks = true; t.integrated = true; break;
    case MCP_CODEC_SK_IBP32_IVB: t.blocks = true; t.integrated = true; t.sk = true; t.chain_tail = true; break;
    case MCP_CODEC_VB: break;
    case MCP_CODEC_SK_BP32_VB: t.blocks = true; t.sk = true; break;
    case MCP_CODEC_SK_FASTPFOR256_VB: t.fastpfor = true; t.fp_block = 256; t.sk = true; break;
    case MCP_CODEC_SK_FASTPFOR128_VB: t.fastpfor = true; t.fp_block = 128; t.sk = true; break;
    // the two below run in the warp-per-list kernels only (`blocks` / `fastpfor` stay false: no batched kernel takes them)
    case MCP_CODEC_XOR128_IVB: t.xorbp = true; t.integrated = true; break;
    case MCP_CODEC_OPTPFD_VB: t.optpfd = true; break;
    }
    return t;
}

int64_t codec_max_words(int codec, int64_t n)
{
    // worst case: 32 bits/int packed + one header per 32-int block, or 5 VByte bytes per tail int
    // (tail < 256 ints), plus FastPFOR page/bitmap/exception-count words.  Generous on purpose;
    // the reference sizes its scratch 4n+1024 (LoadPortfolios.java:611).
    // Per FastPFOR page: offset, byte count, byte padding, bitmap and up to 31 x (count + one rounding word).
    int64_t w = n + n / 16 + (n / kFpPage + 1) * 80 + 400;
    // VariableByte alone spends up to 5 bytes on every int (values >= 2^28: negative ints, wrapped deltas)
    if ((codec & MCP_CODEC_MASK) == MCP_CODEC_VB) w = std::max<int64_t>(w, (5 * n + 3) / 4 + 2);
    // OptPFD: a block may spend up to 4 x 20 + 2 x 127 S16 words + its header where BinaryPacking spends 4 x 32 + 1
    if ((codec & MCP_CODEC_MASK) == MCP_CODEC_OPTPFD_VB) w += 2 * n + 2 * (n / 128);
    return w;
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

// One list through a codec of the BinaryPacking family (Composition / IntCompressor framing around
// BinaryPacking or IntegratedBinaryPacking + a VariableByte tail), by one warp.  Returns the words produced,
// including the loader's trailing zero int when `app`.
template <bool WRITE>
__device__ __forceinline__ int64_t warp_bp_list_encode(const CodecTraits tr, const ValSeq x, int64_t n, const WordSink sink, bool app,
                                                       uint32_t *stage, int lane)
{
    int64_t pos = 0;
    if (tr.sk) { // IntCompressor / IntegratedIntCompressor: compressed[0] = input.length
        if (WRITE && lane == 0) sink.put(0, (uint32_t)n);
        pos = 1;
    }
    if (n > 0) {
        const int64_t m = n & ~(int64_t)31;
        if (!tr.sk) { // Composition: F1 header, or the 0 marker when F1 wrote nothing (Composition.java:45-48)
            if (WRITE && lane == 0) sink.put(pos, (uint32_t)m);
            ++pos;
        }
        pos += warp_bp32_encode<WRITE>(x, m, tr.integrated, sink, pos, lane);
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
                pos += warp_vb_encode<WRITE>(getv, r, sink, pos, stage, lane);
            } else {
                auto getv = [&](int64_t i) -> uint32_t { return x(m + i); };
                pos += warp_vb_encode<WRITE>(getv, r, sink, pos, stage, lane);
            }
        }
    }
    if (app) {
        if (WRITE && lane == 0) sink.put(pos, 0u); // LoadPortfolios.java:622: outpos + 1 ints
        ++pos;
    }
    return pos;
}

// Inverse of warp_bp_list_encode: `avail` words at src decode to exactly n ints at out.  False on a malformed stream.
__device__ __forceinline__ bool warp_bp_list_decode(const CodecTraits tr, const WordSrc src, int64_t avail, int64_t n, uint32_t *out,
                                                    bool delta, int lane)
{
    int64_t p = 0, m = n & ~(int64_t)31;
    uint32_t carry = 0;
    if (tr.sk) {
        if (avail < 1 || (int64_t)src.get(0) != n) return false;
        p = 1;
    }
    if (n == 0) return true;
    if (!tr.sk) {
        if (p >= avail) return false;
        if ((int64_t)src.get(p++) != m) return false;
    }
    if (m > 0) {
        const int64_t used = warp_bp32_decode(src, p, avail, m, tr.integrated, out, carry, lane);
        if (used < 0) return false;
        p += used;
    }
    if (n > m) {
        const bool ok = tr.integrated ? warp_vb_decode<true>(src, p, avail, n - m, out + m, tr.chain_tail ? carry : 0u, lane)
                                      : warp_vb_decode<false>(src, p, avail, n - m, out + m, 0u, lane);
        if (!ok) return false;
    }
    if (delta) {
        __syncwarp();
        warp_inverse_delta(out, n, lane);
    }
    return true;
}

// ------------------------------------------------------------ XorBinaryPacking
// differential/XorBinaryPacking.java:31-71: blocks of 128 values; one header word (four widths), then four sub-blocks of
// 32, every value XOR-ed with its predecessor (the list's first one with 0), packed without masking at the width of the
// sub-block's OR.  Returns words produced.
template <bool WRITE>
__device__ __forceinline__ int64_t warp_xor128_encode(const ValSeq &x, int64_t m, const WordSink &sink, int64_t pos, int lane)
{
    const int64_t pos0 = pos;
    for (int64_t s = 0; s < m; s += 128) {
        uint32_t d[4];
        int b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t base = s + 32 * j;
            const uint32_t raw = x(base + lane);
            uint32_t prev = __shfl_up_sync(kFull, raw, 1);
            if (lane == 0) prev = base > 0 ? x(base - 1) : 0u;
            d[j] = raw ^ prev;
            b[j] = bits32(__reduce_or_sync(kFull, d[j]));
        }
        if (WRITE && lane == 0)
            sink.put(pos, ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | (uint32_t)b[3]);
        ++pos;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (WRITE) {
                const uint32_t w = warp_pack_word(d[j], b[j], lane);
                if (lane < b[j]) sink.put(pos + lane, w);
            }
            pos += b[j];
        }
    }
    return pos - pos0;
}

// XorBinaryPacking.java:73-111.  Returns words consumed, -1 on a malformed width; `carry` = the last value decoded.
__device__ __forceinline__ int64_t warp_xor128_decode(const WordSrc &src, int64_t p, int64_t avail, int64_t m, uint32_t *out,
                                                      uint32_t &carry, int lane)
{
    const int64_t p0 = p;
    for (int64_t s = 0; s < m; s += 128) {
        if (p >= avail) return -1;
        const uint32_t h = src.get(p++);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = (int)((h >> (24 - 8 * j)) & 255u);
            if (b > 32 || p + b > avail) return -1;
            const uint32_t word = (lane < b) ? src.get(p + lane) : 0u;
            uint32_t v = warp_unpack_val(word, b, lane);
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) { // inclusive XOR scan: value i = x_0 ^ ... ^ x_i ^ carry
                const uint32_t t = __shfl_up_sync(kFull, v, dd);
                if (lane >= dd) v ^= t;
            }
            v ^= carry;
            out[s + 32 * j + lane] = v;
            carry = __shfl_sync(kFull, v, 31);
            p += b;
        }
    }
    return p - p0;
}

// ------------------------------------------------------------------ OptPFD
// OptPFD.java:35-200 with S16.java (Simple16 as adapted for the PFD codecs).  Blocks of 128 values: one header word
// b | c << 8 | s << 16 (b = INDEX into kOptBits, c = exceptions, s = S16 words), the S16-coded exceptions (the c high
// parts in block order, then their c positions), four fastpack groups at kOptBits[b] bits.
__device__ __constant__ int kOptBits[17] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16, 20, 32};
__device__ __constant__ int kOptInvBits[33] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 14, 14, 15, 15, 15, 15, 16, 16, 16, 16, 16, 16,
                                               16, 16, 16, 16, 16, 16};
__device__ __constant__ int kS16Num[16] = {28, 21, 21, 21, 14, 9, 8, 7, 6, 6, 5, 5, 4, 3, 2, 1};
__device__ __constant__ unsigned char kS16Bits[16][28] = {
    {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
    {2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
    {1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1},
    {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2},
    {2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2},
    {4, 3, 3, 3, 3, 3, 3, 3, 3}, {3, 4, 4, 4, 4, 3, 3, 3}, {4, 4, 4, 4, 4, 4, 4}, {5, 5, 5, 5, 4, 4}, {4, 4, 5, 5, 5, 5},
    {6, 6, 6, 5, 5}, {5, 5, 6, 6, 6}, {7, 7, 7, 7}, {10, 9, 9}, {14, 14}, {28}};

// The exception sequence of a block at width `w`: entry e < c is the e-th exception's high part (block order), entry c + e
// its position -- as a Java int (S16 compares signed).  Written out in shared memory by opt_materialise.
//
// S16.compressblock (S16.java:80-100): the first of the 16 layouts that holds the next min(num, n) entries; returns how many
// it took (0 = "Too big a number": cannot happen, high parts are < 2^28) and the word.  S16.estimatecompress (:57-72) counts
// the words of a sequence by the same steps.
// By the whole warp: lane j holds entry first + j.  Layout idx fits when every one of its first `num` entries is below
// 2^bits[idx][j] as a Java int (a negative int fits anything, S16.java:86); all 16 layouts are tested at once: `tab` holds, per
// lane j and per bit length d of an entry, the 16-bit set of layouts that take such an entry at place j (all of them beyond a
// layout's last place), so a word is one table read, one AND across the warp and a find-first-set.  Behind the sets: the
// shift of entry j in the word of layout idx (s16_tables builds both, in shared memory).
constexpr int kS16TabBytes = 32 * 32 * 2 + 16 * 32;

__device__ __forceinline__ void s16_tables(uint8_t *tab, int tid, int nthreads)
{
    uint16_t *sets = reinterpret_cast<uint16_t *>(tab);
    for (int i = tid; i < 1024; i += nthreads) {
        const int j = i >> 5, d = i & 31;
        unsigned set = 0;
        for (int idx = 0; idx < 16; ++idx)
            if (j >= kS16Num[idx] || d <= (int)kS16Bits[idx][j < 28 ? j : 0]) set |= 1u << idx;
        sets[i] = (uint16_t)set;
    }
    for (int i = tid; i < 512; i += nthreads) {
        const int idx = i >> 5, j = i & 31;
        int sh = 0;
        for (int q = 0; q < j && q < kS16Num[idx]; ++q) sh += kS16Bits[idx][q];
        tab[2048 + i] = (uint8_t)(sh > 31 ? 31 : sh);
    }
}

template <bool WORD>
__device__ __forceinline__ int warp_s16_block(const int32_t *seq, int first, int n, const uint8_t *tab, int lane, uint32_t *word)
{
    const bool mine = lane < n && lane < 28;
    const int32_t v = mine ? seq[first + lane] : 0;
    const int need = v < 0 ? 0 : 32 - __clz(v);
    unsigned set = mine ? reinterpret_cast<const uint16_t *>(tab)[32 * lane + need] : 0xffffu;
    set = __reduce_and_sync(kFull, set);
    if (set == 0) { if (WORD) *word = 0; return 0; }
    const int idx = __ffs(set) - 1;
    const int num = kS16Num[idx] < n ? kS16Num[idx] : n;
    if (WORD) {
        const uint32_t part = lane < num ? (uint32_t)v << tab[2048 + 32 * idx + lane] : 0u;
        *word = ((uint32_t)idx << 28) | __reduce_or_sync(kFull, part);
    }
    return num;
}

// exception bits of the staged block at width w (all lanes compute the same four words)
__device__ __forceinline__ void opt_mask(const uint32_t (&v)[4], int w, uint32_t (&mask)[4])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) mask[j] = __ballot_sync(kFull, w < 32 && (v[j] >> (w & 31)) != 0u);
}

// the block's exception sequence at width w, written out by the warp: seq[e] = e-th exception's high part (block order),
// seq[c + e] = its position; returns c
__device__ __forceinline__ int opt_materialise(const uint32_t (&v)[4], int w, int32_t *seq, int lane)
{
    uint32_t mk[4];
    opt_mask(v, w, mk);
    const int c = __popc(mk[0]) + __popc(mk[1]) + __popc(mk[2]) + __popc(mk[3]);
    if (c == 0 || c >= 128) return c; // (128 exceptions: the candidate is not priced, OptPFD.java:77)
    int before = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if ((mk[j] >> lane) & 1u) {
            const int e = before + __popc(mk[j] & ((1u << lane) - 1u));
            seq[e] = (int32_t)(v[j] >> (w & 31));
            seq[c + e] = 32 * j + lane;
        }
        before += __popc(mk[j]);
    }
    return c;
}

template <bool WRITE>
__device__ int64_t warp_optpfd_encode(const ValSeq &x, int64_t m, const WordSink &sink, int64_t pos, int32_t *seq /* 256 words */,
                                      const uint8_t *tab /* s16_tables */, int lane)
{
    const int64_t pos0 = pos;
    for (int64_t s = 0; s < m; s += 128) {
        uint32_t v[4];
        uint32_t orv = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[j] = x(s + 32 * j + lane); orv |= v[j]; }
        // ---- getBestBFromData (OptPFD.java:61-98): candidates mini .. 15; `mini` is computed from bit WIDTHS but used as the
        // first INDEX, as the reference does; ties go to the later candidate (<=).  Per candidate the warp writes the exception
        // sequence out (Simple16's greedy layout search reads every entry many times) and prices it together, a word per step.
        const int mb = bits32(__reduce_or_sync(kFull, orv));
        int mini = 0;
        if (mini + 28 < kOptBits[kOptInvBits[mb]]) mini = kOptBits[kOptInvBits[mb]] - 28;
        unsigned key = 0xffffffffu; // (cost << 5) | (31 - candidate): the minimum is the cheapest, latest candidate
        for (int cand = mini; cand < 16; ++cand) {
            const int c = opt_materialise(v, kOptBits[cand], seq, lane);
            __syncwarp();
            if (c < 128) {
                int words = 0;
                for (int at = 0, n = 2 * c; at < n; ++words) {
                    uint32_t word;
                    const int took = warp_s16_block<false>(seq, at, n - at, tab, lane, &word);
                    if (took <= 0) { words = 1 << 20; break; }
                    at += took;
                }
                const int cost = 4 * kOptBits[cand] + words;
                if (cost <= 4 * 32) key = min(key, ((unsigned)cost << 5) | (unsigned)(31 - cand));
            }
            __syncwarp();
        }
        const int b = key == 0xffffffffu ? 16 : 31 - (int)(key & 31u);
        const int w = kOptBits[b];
        const int c = b < 16 ? opt_materialise(v, w, seq, lane) : 0;
        __syncwarp();
        // ---- header, S16 words, packed groups (fastpack masks to w bits)
        int sw = 0;
        for (int at = 0, n = 2 * c; at < n; ++sw) {
            uint32_t word;
            const int took = warp_s16_block<true>(seq, at, n - at, tab, lane, &word);
            if (took <= 0) break;
            if (WRITE && lane == 0) sink.put(pos + 1 + sw, word);
            at += took;
        }
        if (WRITE && lane == 0) sink.put(pos, (uint32_t)b | ((uint32_t)c << 8) | ((uint32_t)sw << 16));
        pos += 1 + sw;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (WRITE) {
                const uint32_t word = warp_pack_word(w >= 32 ? v[j] : (v[j] & ((1u << w) - 1u)), w, lane);
                if (lane < w) sink.put(pos + lane, word);
            }
            pos += w;
        }
        __syncwarp(); // the sequence has been read: the next block may overwrite it
    }
    return pos - pos0;
}

// OptPFD.java:136-160 (decodePage).  `ex` = 2 x 128 + 28 words of shared memory.  Returns words consumed, -1 if malformed.
__device__ int64_t warp_optpfd_decode(const WordSrc &src, int64_t p, int64_t avail, int64_t m, uint32_t *out, uint32_t *ex, int lane)
{
    const int64_t p0 = p;
    for (int64_t s = 0; s < m; s += 128) {
        if (p >= avail) return -1;
        const uint32_t h = src.get(p++);
        const int b = (int)(h & 255u), c = (int)((h >> 8) & 255u), sw = (int)(h >> 16);
        if (b > 16 || c > 128 || p + sw > avail) return -1;
        const int w = kOptBits[b];
        if (lane == 0) { // S16.uncompress (S16.java:135-147): at most 2c values from sw words
            int at = 0, left = 2 * c;
            for (int k = 0; k < sw; ++k) {
                const uint32_t word = src.get(p + k);
                const int idx = (int)(word >> 28);
                const int num = kS16Num[idx] < left ? kS16Num[idx] : left;
                int bits = 0;
                for (int j = 0; j < num; ++j) {
                    const int bw = kS16Bits[idx][j];
                    ex[at + j] = (word >> bits) & (0xffffffffu >> (32 - bw));
                    bits += bw;
                }
                at += num;
                left -= num;
            }
            ex[2 * 128 + 27] = (uint32_t)left; // values the stream failed to deliver
        }
        p += sw;
        if (p + 4 * (int64_t)w > avail) return -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t word = (lane < w) ? src.get(p + lane) : 0u;
            out[s + 32 * j + lane] = warp_unpack_val(word, w, lane);
            p += w;
        }
        __syncwarp();
        if (ex[2 * 128 + 27] != 0u) return -1;
        for (int k = lane; k < c; k += 32) out[s + (ex[k + c] & 127u)] |= w >= 32 ? 0u : ex[k] << w; // Java: << (w & 31); w = 32 has no exceptions
        __syncwarp();
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

// decodePage (FastPFOR.java:233-307).  Returns words consumed or -1.  GLD: the page lies in global memory (read-only
// loads); otherwise wherever the pointer says (a page staged in shared memory).
template <bool GLD>
__device__ __forceinline__ uint32_t page_word(const uint32_t *p) { return GLD ? __ldg(p) : *p; }

template <int ROUNDS, bool GLD = true>
__device__ int64_t warp_fp_decode_page(const uint32_t *inw, int64_t avail, int n, uint32_t *out, int *stream_base,
                                       int *stream_pos, int lane)
{
    constexpr int B = ROUNDS * 32;
    if (avail < 3) return -1;
    const int64_t where_meta = page_word<GLD>(inw);
    if (where_meta < 1 || where_meta >= avail) return -1;
    const int64_t bytesize = page_word<GLD>(inw + where_meta);
    const int64_t meta_words = (bytesize + 3) >> 2;
    int64_t e = where_meta + 1 + meta_words;
    if (bytesize < 0 || e >= avail) return -1;
    const uint32_t bitmap = page_word<GLD>(inw + e);
    ++e;
    // stream table (serial over set bits: at most 31)
    if (lane == 0) {
        int64_t q = e;
        bool ok = true;
        for (int k = 2; k <= 32; ++k) {
            stream_pos[k] = 0;
            if (bitmap & (1u << (k - 1))) {
                if (q >= avail) { ok = false; break; }
                const int64_t size = page_word<GLD>(inw + q);
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
            const uint32_t word = (lane < b) ? page_word<GLD>(inw + p + lane) : 0u;
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
            bool past_end = false;
            for (int ex = lane; ex < c; ex += 32) {
                const int posn = meta[mpos + 3 + ex];
                if (posn >= B) continue;
                uint32_t hi = 1u;
                if (k >= 2) {
                    const uint32_t *sw = inw + stream_base[k];
                    const int64_t bit = (int64_t)(spos + ex) * k;
                    const int64_t w = bit >> 5;
                    const int sh = (int)(bit & 31);
                    // a malformed stream may claim more exceptions than its streams hold: never read past the page
                    if (stream_base[k] + w + (sh + k > 32 ? 1 : 0) >= page_end) { past_end = true; continue; }
                    const uint32_t lo = page_word<GLD>(sw + w);
                    const uint32_t hw = (sh + k > 32) ? page_word<GLD>(sw + w + 1) : 0u;
                    hi = __funnelshift_r(lo, hw, sh);
                    if (k < 32) hi &= (1u << k) - 1u;
                }
                o[posn] |= hi << b;
            }
            __syncwarp();
            if (__any_sync(kFull, past_end)) return -1;
            if (lane == 0) stream_pos[k] = spos + c;
            mpos += 3 + c;
        } else
            mpos += 2;
        __syncwarp();
    }
    return page_end;
}

// ------------------------------------------------------------------- kernels
template <bool WRITE, bool OPT> // OPT: the OptPFD instantiation (16 KB of shared memory the other codecs should not pay for)
__global__ void __launch_bounds__(kEncWarps * 32, 3) k_encode(const EncArgs a)
{
    __shared__ uint32_t s_stage[kEncWarps][48];
    __shared__ int32_t s_blk[kEncWarps][OPT ? 256 : 1]; // OptPFD: the exception sequence of the candidate being priced
    __shared__ __align__(4) uint8_t s_s16[OPT ? kS16TabBytes : 4];
    if (OPT) { s16_tables(s_s16, threadIdx.x, kEncWarps * 32); __syncthreads(); }
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
                if (!tr.sk) { // FastPFOR.compress's count word, or Composition's 0 marker; headlessCompress has neither
                    if (WRITE && lane == 0) sink.put(pos, (uint32_t)m);
                    ++pos;
                }
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
            } else if (tr.xorbp || tr.optpfd) {
                m = n - n % 128; // XorBinaryPacking.java:33 / OptPFD.java:170
                if (WRITE && lane == 0) sink.put(pos, (uint32_t)m); // F1's count word, or Composition's 0 marker
                ++pos;
                if (tr.xorbp) pos += warp_xor128_encode<WRITE>(x, m, sink, pos, lane);
                else if (OPT) pos += warp_optpfd_encode<WRITE>(x, m, sink, pos, s_blk[wib], s_s16, lane);
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
    __shared__ uint32_t s_ex[kEncWarps][2 * 128 + 28]; // OptPFD: the block's S16-decoded exceptions
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
            if (!tr.sk && (tr.fastpfor || tr.blocks || tr.xorbp || tr.optpfd)) {
                ok = p < avail;
                if (ok) {
                    m = src.get(p++);
                    const int64_t blk = tr.fastpfor ? tr.fp_block : (tr.blocks ? 32 : 128);
                    ok = (m == n - n % blk);
                }
            } else if (tr.blocks || tr.fastpfor)
                m = n - n % (tr.fastpfor ? tr.fp_block : 32);
            if (ok && m > 0 && (tr.xorbp || tr.optpfd)) {
                const int64_t used = tr.xorbp ? warp_xor128_decode(src, p, avail, m, out, carry, lane)
                                              : warp_optpfd_decode(src, p, avail, m, out, s_ex[wib], lane);
                if (used < 0) ok = false; else p += used;
            } else if (ok && m > 0) {
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
            if (tr.xorbp || tr.optpfd) {
                m = src.get(p++);
                ok = m >= 0 && m % 128 == 0;
                for (int64_t s = 0; ok && s < m; s += 128) { // walk the block headers
                    if (p >= avail) { ok = false; break; }
                    const uint32_t h = src.get(p);
                    if (tr.xorbp) p += 1 + (h >> 24) + ((h >> 16) & 255u) + ((h >> 8) & 255u) + (h & 255u);
                    else if ((h & 255u) > 16u) ok = false;
                    else p += 1 + (h >> 16) + 4 * kOptBits[h & 255u];
                }
                if (p > avail) ok = false;
            } else if (tr.fastpfor || tr.blocks) {
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

// ============================================================ batched fast path (BinaryPacking family)
// The kernels above give one warp to one list, which for the ~100-int breach lists of BASELINE config C4 leaves the
// SMs latency- and instruction-bound (4 % of the HBM roofline).  The fast path turns the work around:
//   * a CTA owns a BATCH of consecutive lists whose values (<= FE_VCAP ints) it stages in shared memory with
//     coalesced loads;
//   * ONE THREAD packs (or unpacks) ONE 32-int block from shared memory with a 64-bit shift register -- ~5 thread
//     instructions per integer instead of ~5 WARP instructions per integer;
//   * one thread per list writes the list header and the VariableByte tail (< 32 values);
//   * the batch's output is assembled in shared memory and leaves with coalesced stores (byte-swapped on the way
//     for the loader's big-endian framing);
//   * encode is SINGLE PASS: the CSR output offsets come from a decoupled look-back over per-batch totals (batches
//     take dynamic tickets, so a batch only ever waits on batches that are already running) instead of a size pass,
//     a scan and an emit pass -- the input is read once.
// Batches that do not fit the staging areas (very long lists, incompressible data) are handled inside the same kernel
// by the per-warp routines above, so any input is accepted.  Codecs: BP32_VB, SK_BP32_VB (with or without the Delta
// pre-pass) and IBP32_IVB, SK_IBP32_IVB (without).
constexpr int FE_THREADS = 128;
constexpr int FE_VCAP = 4096;   // staged values per batch: one 32-int block per thread
constexpr int FE_MAXL = 128;    // lists per batch
constexpr int FE_MINB = 7;      // resident CTAs per SM the kernels are compiled for (<= 72 registers, <= 31 KB shared memory)
constexpr int FE_OCAP = 2048;   // staged compressed words per batch
constexpr int FE_TCAP = 1024;   // VariableByte tail values per batch (n mod 32 of every list)
constexpr int FE_VPAD = FE_VCAP + FE_VCAP / 32 + 2;

__host__ __device__ inline bool fast_codec_ok(int codec, bool delta)
{
    const CodecTraits t = codec_traits(codec);
    return t.blocks && !t.fastpfor && !(t.integrated && delta);
}

constexpr uint64_t kLbAgg = 1ull << 62, kLbInc = 2ull << 62, kLbMask = (1ull << 62) - 1; // look-back state: flag | words

__device__ __forceinline__ int sw33(int i) { return i + (i >> 5); } // 33-word stride per 32 values: conflict-free block reads

// the look-back state word carries everything a reader needs (flag | words): readers that consume nothing else the
// writer stored can poll it relaxed
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Decoupled look-back by one warp, 32 predecessors per round: the word offset of batch `batch` in the output = the sum
// of the totals of all batches before it.  Every batch publishes its aggregate before it looks back and its inclusive
// prefix after.  (Measured: polling 128 predecessors per round -- four state words per lane -- is SLOWER, 0.51 vs 0.38 ms
// on C4: the wait is for the nearest predecessors' aggregates, not for the length of the walk.)
__device__ __forceinline__ long long lookback_warp(uint64_t *state, int64_t batch, long long T, int lane)
{
    long long base = 0;
    if (batch > 0) {
        for (int64_t j = batch - 1;; j -= 32) {
            const int64_t idx = j - lane;
            uint64_t st = kLbInc; // before the first batch: an inclusive prefix of 0
            if (idx >= 0) while (((st = ld_relaxed_u64(state + idx)) >> 62) == 0) __nanosleep(64);
            const unsigned inc = __ballot_sync(kFull, (st >> 62) == 2);
            const int first = inc ? __ffs(inc) - 1 : 31; // nearest predecessor with an inclusive prefix
            long long v = lane <= first ? (long long)(st & kLbMask) : 0;
#pragma unroll
            for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
            base += v;
            if (inc) break;
        }
        if (lane == 0) st_relaxed_u64(state + batch, kLbInc | (uint64_t)(base + T));
    }
    return base;
}

// exclusive scan of one int per thread over the CTA (FE_THREADS = 4 warps); returns the exclusive prefix, *total = sum
__device__ __forceinline__ int cta_excl_scan(int v, int *s_w /* [5] */, int *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += t;
    }
    __syncthreads(); // s_w may still be read from a previous call
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < FE_THREADS / 32; ++w) {
        const int x = s_w[w];
        if (w < warp) off += x;
        tot += x;
    }
    *total = tot;
    return off + incl - v;
}

struct FastEncArgs {
    int codec;
    bool delta, app;
    const uint32_t *vals;
    const int64_t *list_off; // null -> uniform_len
    int64_t uniform_len;
    int64_t nlists;
    int lists_per_batch;
    int64_t nbatches;
    int64_t *out_off;        // [nlists + 1] bytes (written)
    uint8_t *out;
    int64_t cap;             // bytes
    uint64_t *state;         // [nbatches] look-back state, zeroed; top 2 bits: 1 = aggregate, 2 = inclusive prefix
    unsigned long long *ticket; // zeroed
    int64_t *total;          // bytes (written by the last batch)
    int *err;                // set to 1 when the output would exceed cap
};

// header words in front of block k of a list with nb blocks (BinaryPacking.java:54-89): one per group of four
// blocks while full groups last, then one per block
__device__ __forceinline__ int bp_hdrs_before(int k, int nb)
{
    const int ng = nb >> 2;
    return k < 4 * ng ? (k >> 2) + 1 : ng + (k - 4 * ng) + 1;
}

// value i of a list's VariableByte tail as the codec sees it (g = its index in the staged batch, nb = blocks before it):
// Delta pre-pass / IntegratedVariableByte gaps; the non-skippable integrated composition restarts the chain at 0
// (IntegratedVariableByte.java:40), the skippable one continues it (:188-189)
__device__ __forceinline__ uint32_t tail_value(const uint32_t *s_val, int g, int i, int nb, const CodecTraits tr, bool delta)
{
    const uint32_t x = s_val[sw33(g)];
    if (tr.integrated) return (i > 0 || (tr.chain_tail && nb > 0)) ? x - s_val[sw33(g - 1)] : x;
    return (delta && (i > 0 || nb > 0)) ? x - s_val[sw33(g - 1)] : x;
}

__global__ void __launch_bounds__(FE_THREADS, FE_MINB) k_bp_encode_fast(const FastEncArgs a)
{
    __shared__ int s_tw[FE_MAXL];       // bytes of each list's VariableByte tail
    __shared__ int s_tb[FE_MAXL];       // first slot of each list in the tail table
    __shared__ uint16_t s_tt[FE_TCAP];  // tail table: (list << 8) | byte length, later (list << 8) | byte offset
    __shared__ uint32_t s_val[FE_VPAD];
    __shared__ uint32_t s_out[FE_OCAP];
    __shared__ int s_in[FE_MAXL + 1];   // value offset of each list inside the batch
    __shared__ int s_ib[FE_MAXL + 1];   // first block item of each list
    __shared__ long long s_lo[FE_MAXL + 1]; // word offset of each list's output inside the batch
    __shared__ int s_bsum[FE_THREADS + 1]; // exclusive prefix of the block widths over the items
    __shared__ uint8_t s_w[FE_THREADS];    // width of each block item
    __shared__ uint8_t s_il[FE_THREADS];   // list of each block item
    __shared__ uint32_t s_stage[FE_THREADS / 32][48];
    __shared__ int s_scan[8];
    __shared__ long long s_base;
    __shared__ long long s_vbase;
    __shared__ unsigned long long s_batch;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const CodecTraits tr = codec_traits(a.codec);
    const bool diff = a.delta || tr.integrated;
    if (tid == 0) s_batch = atomicAdd(a.ticket, 1ull);
    __syncthreads();
    const int64_t batch = (int64_t)s_batch;
    const int64_t l0 = batch * a.lists_per_batch;
    const int nl = (int)min((int64_t)a.lists_per_batch, a.nlists - l0);
    // ---- list extents
    {
        const int64_t vb = a.list_off ? a.list_off[l0] : l0 * a.uniform_len;
        if (tid == 0) s_vbase = vb;
        for (int j = tid; j <= nl; j += FE_THREADS) {
            const int64_t o = (a.list_off ? a.list_off[l0 + j] : (l0 + j) * a.uniform_len) - vb;
            s_in[j] = (int)(o > 0x7fffffff ? 0x7fffffff : o);
        }
    }
    __syncthreads();
    const int64_t vbase = s_vbase;
    const int nv = s_in[nl];
    // slots of the lists' VariableByte tails (n mod 32 values each) in the tail table
    int ntv = 0;
    const int my_r = tid < nl ? (s_in[tid + 1] - s_in[tid]) & 31 : 0;
    const int my_tb = cta_excl_scan(my_r, s_scan, &ntv);
    if (tid < nl) s_tb[tid] = my_tb;
    const bool staged = nv <= FE_VCAP && ntv <= FE_TCAP;
    bool fit = false;
    long long T = 0; // words this batch produces
    int my_pos = 0, my_b = 0, my_list = 0, my_k = 0, my_start = 0;
    bool have_item = false;
    if (staged) {
        // ---- stage the values (coalesced), 33-word stride per 32 values
        for (int i = tid; i < nv; i += FE_THREADS) s_val[sw33(i)] = __ldg(a.vals + vbase + i);
        // ---- block items: list j owns items [s_ib[j], s_ib[j+1])
        int ni = 0;
        {
            const int nbj = tid < nl ? (s_in[tid + 1] - s_in[tid]) >> 5 : 0;
            const int ex = cta_excl_scan(nbj, s_scan, &ni);
            if (tid < nl) s_ib[tid] = ex;
            if (tid == 0) s_ib[nl] = ni;
            __syncthreads();
            if (tid < nl)
                for (int k = 0; k < nbj; ++k) s_il[ex + k] = (uint8_t)tid;
        }
        __syncthreads();
        // ---- width of my block (Util.maxbits / maxdiffbits)
        uint32_t v[32];
        have_item = tid < ni;
        if (have_item) {
            my_list = s_il[tid];
            my_k = tid - s_ib[my_list];
            my_start = s_in[my_list] + 32 * my_k;
            uint32_t prev = (diff && my_k > 0) ? s_val[sw33(my_start - 1)] : 0u;
            uint32_t orv = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const uint32_t x = s_val[sw33(my_start + i)];
                v[i] = diff ? x - prev : x;
                if (diff) prev = x;
                orv |= v[i];
            }
            my_b = bits32(orv);
            s_w[tid] = (uint8_t)my_b;
        }
        int tb = 0;
        const int bex = cta_excl_scan(have_item ? my_b : 0, s_scan, &tb);
        s_bsum[tid] = bex;
        if (tid == 0) s_bsum[FE_THREADS] = tb;
        __syncthreads();
        // ---- VariableByte tails (n mod 32 values per list), one THREAD per tail value: (1) every list claims a run of
        // slots in the tail table, (2) every value computes its byte length, (3) one thread per list turns the lengths
        // of its < 32 values into byte offsets.  The table entry is (list << 8) | length, then (list << 8) | offset.
        if (tid < nl)
            for (int i = 0; i < my_r; ++i) s_tt[my_tb + i] = (uint16_t)(tid << 8);
        __syncthreads();
        {
            for (int v = tid; v < ntv; v += FE_THREADS) {
                const int j = s_tt[v] >> 8, i = v - s_tb[j];
                const int nb = (s_in[j + 1] - s_in[j]) >> 5;
                s_tt[v] = (uint16_t)((j << 8) | vb_len(tail_value(s_val, s_in[j] + 32 * nb + i, i, nb, tr, a.delta)));
            }
            __syncthreads();
            if (tid < nl) {
                const int r = (s_in[tid + 1] - s_in[tid]) & 31, v0 = s_tb[tid];
                int run = 0;
                for (int i = 0; i < r; ++i) { const int len = s_tt[v0 + i] & 255; s_tt[v0 + i] = (uint16_t)((tid << 8) | run); run += len; }
                s_tw[tid] = run; // bytes (<= 155)
            }
        }
        __syncthreads();
        // ---- per list: words of the whole list
        int lw = 0;
        if (tid < nl) {
            const int n = s_in[tid + 1] - s_in[tid], nb = n >> 5;
            const int i0 = s_ib[tid];
            const int bw = s_bsum[i0 + nb] - s_bsum[i0]; // i0 + nb <= ni <= FE_THREADS
            const int hdrs = (nb >> 2) + (nb & 3);
            lw = (n > 0 ? 1 + bw + hdrs + ((s_tw[tid] + 3) >> 2) : (tr.sk ? 1 : 0)) + (a.app ? 1 : 0);
        }
        int Tw = 0;
        const int lex = cta_excl_scan(lw, s_scan, &Tw);
        T = Tw;
        if (tid < nl) s_lo[tid] = lex;
        if (tid == 0) s_lo[nl] = Tw;
        __syncthreads();
        fit = T <= FE_OCAP;
        // the batch's total is known: publish it NOW, so that by the time this batch has packed its blocks its
        // predecessors' totals (and usually their prefixes) are already there
        if (tid == 0) st_relaxed_u64(a.state + batch, (batch == 0 ? kLbInc : kLbAgg) | (uint64_t)T);
        if (fit) {
            // ---- pack my block into the staged output (BitPacking.fastpackwithoutmask; IntegratedBitPacking
            // stores the raw values at width 32)
            if (have_item) {
                const int nb = (s_in[my_list + 1] - s_in[my_list]) >> 5;
                my_pos = (int)s_lo[my_list] + 1 + (s_bsum[tid] - s_bsum[s_ib[my_list]]) + bp_hdrs_before(my_k, nb); // word 0 = list header
                uint32_t *o = s_out + my_pos;
                if (my_b == 32) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = tr.integrated ? s_val[sw33(my_start + i)] : v[i];
                } else if (my_b > 0) {
                    uint64_t acc = 0;
                    int fill = 0;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        acc |= (uint64_t)v[i] << fill;
                        fill += my_b;
                        if (fill >= 32) { *o++ = (uint32_t)acc; acc >>= 32; fill -= 32; }
                    }
                }
                const int ng4 = 4 * (nb >> 2);
                if (my_k >= ng4) s_out[my_pos - 1] = (uint32_t)my_b;
                else if ((my_k & 3) == 0)
                    s_out[my_pos - 1] = ((uint32_t)my_b << 24) | ((uint32_t)s_w[tid + 1] << 16) | ((uint32_t)s_w[tid + 2] << 8) | (uint32_t)s_w[tid + 3];
            }
            // ---- per list: header word, zero padding of the VariableByte tail, trailing zero int of the loader's framing
            if (tid < nl) {
                const int n = s_in[tid + 1] - s_in[tid], nb = n >> 5;
                if (n > 0 || tr.sk) s_out[s_lo[tid]] = tr.sk ? (uint32_t)n : (uint32_t)(32 * nb); // IntCompressor: n; Composition: F1 header / 0 marker
                const int tbytes = s_tw[tid], tend = (int)s_lo[tid + 1] - (a.app ? 1 : 0);
                uint8_t *tbp = reinterpret_cast<uint8_t *>(s_out + tend - ((tbytes + 3) >> 2));
                for (int q = tbytes; q & 3; ++q) tbp[q] = 0;   // zero-padded to a word (VariableByte.java:62-66)
                if (a.app) s_out[s_lo[tid + 1] - 1] = 0u;       // LoadPortfolios.java:622: outpos + 1 ints
            }
            // ---- VariableByte tails: every value writes its bytes at its offset
            for (int v = tid; v < ntv; v += FE_THREADS) {
                const int e = s_tt[v], j = e >> 8, off = e & 255, i = v - s_tb[j];
                const int nb = (s_in[j + 1] - s_in[j]) >> 5;
                uint32_t d = tail_value(s_val, s_in[j] + 32 * nb + i, i, nb, tr, a.delta);
                uint8_t *tbp = reinterpret_cast<uint8_t *>(s_out + (int)s_lo[j + 1] - (a.app ? 1 : 0) - ((s_tw[j] + 3) >> 2)) + off;
                while (d >= 128u) { *tbp++ = (uint8_t)(d & 127u); d >>= 7; }
                *tbp = (uint8_t)(d | 128u);
            }
        }
    }
    // ---- batches whose values or tails do not fit the staging areas: sizes by the per-warp routines (one list per warp at a time)
    if (!staged) {
        for (int j = warp; j < nl; j += FE_THREADS / 32) {
            const int64_t beg = a.list_off ? a.list_off[l0 + j] : (l0 + j) * a.uniform_len;
            const int64_t n = (a.list_off ? a.list_off[l0 + j + 1] : (l0 + j + 1) * a.uniform_len) - beg;
            const int64_t w = warp_bp_list_encode<false>(tr, ValSeq{a.vals + beg, a.delta}, n, WordSink{nullptr, a.app}, a.app, s_stage[warp], lane);
            if (lane == 0) s_lo[j] = w;
        }
        __syncthreads();
        if (tid == 0) {
            long long run = 0;
            for (int j = 0; j < nl; ++j) { const long long w = s_lo[j]; s_lo[j] = run; run += w; }
            s_lo[nl] = run;
            st_relaxed_u64(a.state + batch, (batch == 0 ? kLbInc : kLbAgg) | (uint64_t)run);
        }
        __syncthreads();
        T = s_lo[nl];
    }
    __syncthreads();
    // ---- decoupled look-back by warp 0, 32 predecessors per round: word offset of this batch in the output.
    // Batches take tickets in launch order, so every predecessor is already running and publishes its aggregate
    // without waiting on anyone: the spin below is bounded.
    if (warp == 0) {
        const long long base = lookback_warp(a.state, batch, T, lane);
        if (lane == 0) {
            s_base = base;
            if (batch == a.nbatches - 1) {
                *a.total = 4 * (base + T);
                a.out_off[a.nlists] = 4 * (base + T);
            }
            if (4 * (base + T) > a.cap) *a.err = 1;
        }
    }
    __syncthreads();
    const int64_t base = s_base;
    for (int j = tid; j < nl; j += FE_THREADS) a.out_off[l0 + j] = 4 * (base + s_lo[j]);
    if (4 * (base + T) > a.cap) return; // caller's buffer too small: report, write nothing
    uint32_t *gout = reinterpret_cast<uint32_t *>(a.out) + base;
    if (fit) {
        for (int i = tid; i < (int)T; i += FE_THREADS) gout[i] = a.app ? bswap32(s_out[i]) : s_out[i];
    } else {
        for (int j = warp; j < nl; j += FE_THREADS / 32) {
            const int64_t beg = a.list_off ? a.list_off[l0 + j] : (l0 + j) * a.uniform_len;
            const int64_t n = (a.list_off ? a.list_off[l0 + j + 1] : (l0 + j + 1) * a.uniform_len) - beg;
            warp_bp_list_encode<true>(tr, ValSeq{a.vals + beg, a.delta}, n, WordSink{gout + s_lo[j], a.app}, a.app, s_stage[warp], lane);
        }
    }
}


struct FastDecArgs {
    int codec;
    bool delta, app;
    const uint8_t *in;
    const int64_t *in_off;
    int64_t nlists;
    int lists_per_batch;
    uint32_t *out;
    const int64_t *out_off;
    int *err;
};

__global__ void __launch_bounds__(FE_THREADS, FE_MINB) k_bp_decode_fast(const FastDecArgs a)
{
    __shared__ uint32_t s_val[FE_VPAD];   // decoded values of the batch (33-word stride per 32)
    __shared__ uint32_t s_wd[FE_OCAP];    // compressed words of the batch, native order
    __shared__ int s_in[FE_MAXL + 1];     // word offset of each list's stream inside the batch
    __shared__ int s_oo[FE_MAXL + 1];     // value offset of each list inside the batch
    __shared__ int s_ib[FE_MAXL + 1];     // first block item of each list
    __shared__ int s_ipos[FE_THREADS];    // word position (in s_wd) of each block item
    __shared__ uint32_t s_tot[FE_THREADS]; // sum of the block's differences / its last value
    __shared__ uint32_t s_carry[FE_THREADS];
    __shared__ uint8_t s_iw[FE_THREADS];
    __shared__ uint8_t s_il[FE_THREADS];
    __shared__ int s_scan[8];
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const CodecTraits tr = codec_traits(a.codec);
    const bool diff = a.delta || tr.integrated;
    const int64_t l0 = (int64_t)blockIdx.x * a.lists_per_batch;
    const int nl = (int)min((int64_t)a.lists_per_batch, a.nlists - l0);
    const int64_t ibase = a.in_off[l0], obase = a.out_off[l0];
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int j = tid; j <= nl; j += FE_THREADS) {
        const int64_t ib = a.in_off[l0 + j] - ibase, ob = a.out_off[l0 + j] - obase;
        if ((a.in_off[l0 + j] & 3) != 0) s_bad = 1;
        s_in[j] = (int)((ib >> 2) > 0x7fffffff ? 0x7fffffff : (ib >> 2));
        s_oo[j] = (int)(ob > 0x7fffffff ? 0x7fffffff : ob);
    }
    __syncthreads();
    if (s_bad) { if (tid == 0) atomicExch(a.err, 1); return; }
    const uint32_t *gin = reinterpret_cast<const uint32_t *>(a.in + ibase);
    const int nw = s_in[nl], nv = s_oo[nl];
    if (nw > FE_OCAP || nv > FE_VCAP) {
        // batch does not fit the staging areas: per-warp routines straight from / to global memory
        for (int j = warp; j < nl; j += FE_THREADS / 32) {
            const int64_t ib = a.in_off[l0 + j], ie = a.in_off[l0 + j + 1];
            const int64_t n = a.out_off[l0 + j + 1] - a.out_off[l0 + j];
            const WordSrc src{reinterpret_cast<const uint32_t *>(a.in + ib), a.app};
            const bool ok = ((ie - ib) & 3) == 0 && warp_bp_list_decode(tr, src, (ie - ib) >> 2, n, a.out + a.out_off[l0 + j], a.delta, lane);
            if (!ok && lane == 0) atomicExch(a.err, 1);
            __syncwarp();
        }
        return;
    }
    for (int i = tid; i < nw; i += FE_THREADS) { const uint32_t x = __ldg(gin + i); s_wd[i] = a.app ? bswap32(x) : x; }
    int ni = 0;
    {
        const int nbj = tid < nl ? (s_oo[tid + 1] - s_oo[tid]) >> 5 : 0;
        const int ex = cta_excl_scan(nbj, s_scan, &ni);
        if (tid < nl) s_ib[tid] = ex;
        if (tid == 0) s_ib[nl] = ni;
    }
    __syncthreads();
    // ---- one thread per list: walk the headers (the position of a header depends on the widths before it), note
    // every block's position and width, decode the VariableByte tail (differences are summed below)
    if (tid < nl) {
        const int n = s_oo[tid + 1] - s_oo[tid], nb = n >> 5, r = n & 31;
        const int avail = s_in[tid + 1] - s_in[tid];
        const uint32_t *w = s_wd + s_in[tid];
        int p = 0, item = s_ib[tid];
        bool ok = true;
        if (tr.sk) { ok = avail >= 1 && w[0] == (uint32_t)n; p = 1; }
        if (ok && n > 0) {
            if (!tr.sk) { ok = p < avail && w[p] == (uint32_t)(32 * nb); ++p; }
            int k = 0;
            for (; ok && k + 4 <= nb; k += 4) { // BinaryPacking.java:102-123
                if (p >= avail) { ok = false; break; }
                const uint32_t h = w[p++];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int b = (int)((h >> (24 - 8 * q)) & 255u);
                    if (b > 32 || p + b > avail) { ok = false; break; }
                    s_il[item] = (uint8_t)tid; s_ipos[item] = s_in[tid] + p; s_iw[item] = (uint8_t)b;
                    ++item; p += b;
                }
            }
            for (; ok && k < nb; ++k) {         // :124-131
                if (p >= avail) { ok = false; break; }
                const uint32_t h = w[p++];
                if (h > 32u || p + (int)h > avail) { ok = false; break; }
                s_il[item] = (uint8_t)tid; s_ipos[item] = s_in[tid] + p; s_iw[item] = (uint8_t)h;
                ++item; p += (int)h;
            }
            // VariableByte.headlessUncompress (VariableByte.java:180-203)
            if (ok && r > 0) {
                const int o0 = s_oo[tid] + 32 * nb;
                int got = 0, s = 0, shift = 0;
                uint32_t v = 0;
                while (got < r) {
                    if (p >= avail) { ok = false; break; }
                    const uint32_t c = (w[p] >> s) & 0xffu;
                    s += 8; p += s >> 5; s &= 31;
                    v += (c & 127u) << (shift & 31);
                    if (c & 128u) { s_val[sw33(o0 + got)] = v; ++got; v = 0; shift = 0; }
                    else shift += 7;
                }
            }
        }
        if (!ok) s_bad = 1;
        for (; item < s_ib[tid + 1]; ++item) { s_il[item] = (uint8_t)tid; s_ipos[item] = 0; s_iw[item] = 0; } // malformed: keep the tables defined
    }
    __syncthreads();
    if (s_bad) { if (tid == 0) atomicExch(a.err, 1); return; }
    // ---- one thread per block: unpack 32 values (BitPacking.fastunpack), running sum inside the block when the
    // stream holds differences
    uint32_t v[32];
    const bool have_item = tid < ni;
    int my_start = 0, my_b = 0;
    if (have_item) {
        const int j = s_il[tid], k = tid - s_ib[j];
        my_start = s_oo[j] + 32 * k;
        my_b = s_iw[tid];
        const uint32_t *ptr = s_wd + s_ipos[tid];
        if (my_b == 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = ptr[i];
        } else {
            const uint32_t mask = (1u << my_b) - 1u;
            uint64_t acc = 0;
            int fill = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (fill < my_b) { acc |= (uint64_t)(*ptr++) << fill; fill += 32; }
                v[i] = (uint32_t)acc & mask;
                acc >>= my_b;
                fill -= my_b;
            }
        }
        const bool raw = tr.integrated && my_b == 32; // IntegratedBitPacking stores raw values at width 32
        if (diff && !raw) {
#pragma unroll
            for (int i = 1; i < 32; ++i) v[i] += v[i - 1];
        }
        s_tot[tid] = v[31];
    }
    __syncthreads();
    // ---- one thread per list: carry of every block (sum of everything before it), and the tail's running sum
    if (diff && tid < nl) {
        const int n = s_oo[tid + 1] - s_oo[tid], nb = n >> 5, r = n & 31;
        uint32_t carry = 0;
        for (int it = s_ib[tid]; it < s_ib[tid] + nb; ++it) {
            const bool raw = tr.integrated && s_iw[it] == 32;
            s_carry[it] = raw ? 0u : carry;
            carry = raw ? s_tot[it] : carry + s_tot[it];
        }
        if (tr.integrated && !tr.chain_tail) carry = 0; // IntegratedVariableByte.java:40 restarts; the skippable form chains
        const int o0 = s_oo[tid] + 32 * nb;
        for (int i = 0; i < r; ++i) { carry += s_val[sw33(o0 + i)]; s_val[sw33(o0 + i)] = carry; }
    }
    __syncthreads();
    if (have_item) {
        const uint32_t c = diff ? s_carry[tid] : 0u;
#pragma unroll
        for (int i = 0; i < 32; ++i) s_val[sw33(my_start + i)] = v[i] + c;
    }
    __syncthreads();
    uint32_t *gout = a.out + obase;
    for (int i = tid; i < nv; i += FE_THREADS) gout[i] = s_val[sw33(i)];
}

// ============================================================ batched fast path (FastPFOR / FastPFOR128)
// Same batch machinery for the codec the north star names: Composition(FastPFOR, VariableByte) on lists of a few
// hundred to a few thousand ints (one page each; the risk path's breach lists are ~1 000).  Per batch (<= 16 lists,
// <= 4 096 values staged in shared memory):
//   * one thread per 32-value group: bit widths into a per-block histogram (FastPFOR.getBestBFromData's freqs);
//   * one thread per block: the cost search (b, cexcept, maxb);
//   * one thread per list: running packed-word / metadata-byte offsets of its blocks and the fill of the 32 exception
//     streams (the serial part: <= 16 blocks);
//   * one thread per group again: fastpack of its 32 values at the block's width, exception positions into the byte
//     metadata and exception high bits into their stream with shared-memory atomicOr;
//   * VariableByte tail (n mod block, up to 255 values): one thread per value, offsets by a warp scan per list;
//   * single pass over the input; output offsets by the decoupled look-back of k_bp_encode_fast.
// A batch that does not fit (a list beyond the staging area, incompressible data) raises flag 2 and the host reruns
// the call through the warp-per-list kernels.
constexpr int FP_MAXL = 16;   // lists per batch
constexpr int FP_MAXBLK = 32; // FastPFOR blocks per batch (4 096 / 128)

template <int BG> // 32-value groups per block: 8 (FastPFOR) or 4 (FastPFOR128)
__global__ void __launch_bounds__(FE_THREADS, 6) k_fp_encode_fast(const FastEncArgs a)
{
    constexpr int B = BG * 32;
    __shared__ uint32_t s_val[FE_VPAD];
    __shared__ uint32_t s_out[FE_OCAP];
    __shared__ int s_in[FP_MAXL + 1];       // value offset of each list inside the batch
    __shared__ int s_fb[FP_MAXL + 1];       // first block of each list
    __shared__ long long s_lo[FP_MAXL + 1]; // word offset of each list's output inside the batch
    __shared__ int s_hist[FP_MAXBLK][33];   // bit-width histogram of each block
    __shared__ uint8_t s_bb[FP_MAXBLK], s_bc[FP_MAXBLK], s_bm[FP_MAXBLK], s_bl[FP_MAXBLK]; // b, cexcept, maxb, list of each block
    __shared__ uint16_t s_bpw[FP_MAXBLK], s_bmo[FP_MAXBLK], s_bsp[FP_MAXBLK]; // packed-word offset, metadata byte offset, stream position
    __shared__ uint16_t s_cntk[FP_MAXL][33];  // exceptions of each width difference k, per list
    __shared__ uint16_t s_sbase[FP_MAXL][33]; // word offset of stream k (its count word) behind the bitmap
    __shared__ int s_pw[FP_MAXL], s_mb[FP_MAXL], s_ew[FP_MAXL]; // packed words, metadata bytes, exception words per list
    __shared__ int s_tw[FP_MAXL], s_tb[FP_MAXL];                // tail bytes, first tail-table slot per list
    __shared__ uint8_t s_ge[FE_THREADS];      // exceptions in each group
    __shared__ uint16_t s_tt[FE_TCAP];        // tail table: (list << 11) | byte length, later | byte offset
    __shared__ int s_scan[8];
    __shared__ long long s_base, s_vbase;
    __shared__ unsigned long long s_batch;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_batch = atomicAdd(a.ticket, 1ull);
    __syncthreads();
    const int64_t batch = (int64_t)s_batch;
    const int64_t l0 = batch * a.lists_per_batch;
    const int nl = (int)min((int64_t)a.lists_per_batch, a.nlists - l0);
    {
        const int64_t vb = a.list_off ? a.list_off[l0] : l0 * a.uniform_len;
        if (tid == 0) s_vbase = vb;
        for (int j = tid; j <= nl; j += FE_THREADS) {
            const int64_t o = (a.list_off ? a.list_off[l0 + j] : (l0 + j) * a.uniform_len) - vb;
            s_in[j] = (int)(o > 0x7fffffff ? 0x7fffffff : o);
        }
    }
    __syncthreads();
    const int64_t vbase = s_vbase;
    const int nv = s_in[nl];
    // blocks and tail slots of every list
    int nblk = 0, ntv = 0;
    const int my_n = tid < nl ? s_in[tid + 1] - s_in[tid] : 0;
    const int my_nb = my_n / B, my_r = my_n - my_nb * B;
    const int my_fb = cta_excl_scan(my_nb, s_scan, &nblk);
    const int my_tb = cta_excl_scan(my_r, s_scan, &ntv);
    if (tid < nl) { s_fb[tid] = my_fb; s_tb[tid] = my_tb; }
    if (tid == 0) s_fb[nl] = nblk;
    bool fit = nv <= FE_VCAP && ntv <= FE_TCAP && nl <= FP_MAXL && nblk <= FP_MAXBLK;
    long long T = 0;
    if (fit) {
        for (int i = tid; i < nv; i += FE_THREADS) s_val[sw33(i)] = __ldg(a.vals + vbase + i);
        for (int i = tid; i < FP_MAXBLK * 33; i += FE_THREADS) (&s_hist[0][0])[i] = 0;
        if (tid < nl) {
            for (int q = 0; q < my_nb; ++q) s_bl[my_fb + q] = (uint8_t)tid;
            for (int i = 0; i < my_r; ++i) s_tt[my_tb + i] = (uint16_t)(tid << 11);
        }
        __syncthreads();
        // ---- group items: thread it <-> block it / BG, group it % BG
        const int bi = tid / BG, g = tid % BG;
        const bool have = bi < nblk;
        uint32_t v[32];
        int my_start = 0;
        if (have) {
            const int j = s_bl[bi], q = bi - s_fb[j];
            my_start = s_in[j] + q * B + g * 32;
            uint32_t prev = (a.delta && my_start > s_in[j]) ? s_val[sw33(my_start - 1)] : 0u;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const uint32_t x = s_val[sw33(my_start + i)];
                v[i] = a.delta ? x - prev : x;
                if (a.delta) prev = x;
                atomicAdd(&s_hist[bi][bits32(v[i])], 1);
            }
        }
        // ---- tails: byte length of every value
        for (int t = tid; t < ntv; t += FE_THREADS) {
            const int j = s_tt[t] >> 11, i = t - s_tb[j];
            const int gidx = s_in[j] + ((s_in[j + 1] - s_in[j]) / B) * B + i;
            const uint32_t x = s_val[sw33(gidx)];
            const uint32_t d = (a.delta && gidx > s_in[j]) ? x - s_val[sw33(gidx - 1)] : x;
            s_tt[t] = (uint16_t)((j << 11) | vb_len(d));
        }
        __syncthreads();
        // ---- one thread per block: getBestBFromData (FastPFOR.java:107-134)
        if (tid < nblk) {
            const int *f = s_hist[tid];
            int maxb = 32;
            while (maxb > 0 && f[maxb] == 0) --maxb;
            int bestb = maxb, bestc = 0, bestcost = maxb * B, c = 0;
            for (int cand = maxb - 1; cand >= 0; --cand) {
                c += f[cand + 1];
                if (c == B) break;
                int cost = c * 8 + c * (maxb - cand) + cand * B + 8;
                if (maxb - cand == 1) cost -= c;
                if (cost < bestcost) { bestcost = cost; bestb = cand; bestc = c; }
            }
            s_bb[tid] = (uint8_t)bestb; s_bc[tid] = (uint8_t)bestc; s_bm[tid] = (uint8_t)maxb;
        }
        // ---- tails: byte offsets, a warp per list (inclusive scans over chunks of 32 values)
        for (int j = warp; j < nl; j += FE_THREADS / 32) {
            const int n = s_in[j + 1] - s_in[j], r = n - (n / B) * B, t0 = s_tb[j];
            int run = 0;
            for (int base = 0; base < r; base += 32) {
                const int i = base + lane;
                const int len = i < r ? (s_tt[t0 + i] & 2047) : 0;
                const int incl = (int)warp_incl_scan((uint32_t)len, lane);
                if (i < r) s_tt[t0 + i] = (uint16_t)((j << 11) | (run + incl - len));
                run += __shfl_sync(kFull, incl, 31);
            }
            if (lane == 0) s_tw[j] = run;
        }
        __syncthreads();
        // ---- one thread per list: offsets of its blocks, fill of its exception streams (encodePage's bookkeeping)
        int lw = 0;
        if (tid < nl) {
            uint16_t cnt[33];
#pragma unroll
            for (int k = 0; k < 33; ++k) cnt[k] = 0;
            int pw = 0, mb = 0;
            for (int q = 0; q < my_nb; ++q) {
                const int b_ = my_fb + q, b = s_bb[b_], c = s_bc[b_];
                s_bpw[b_] = (uint16_t)pw; pw += BG * b;
                s_bmo[b_] = (uint16_t)mb; mb += 2 + (c ? 1 + c : 0);
                if (c) { const int k = s_bm[b_] - b; s_bsp[b_] = cnt[k]; cnt[k] += (uint16_t)c; }
            }
            int ew = 0;
            for (int k = 2; k <= 32; ++k) {
                s_sbase[tid][k] = (uint16_t)ew;
                s_cntk[tid][k] = cnt[k];
                if (cnt[k]) ew += 1 + ((cnt[k] * k + 31) >> 5);
            }
            s_pw[tid] = pw; s_mb[tid] = mb; s_ew[tid] = ew;
            const int page = my_nb ? 1 + pw + 1 + ((mb + 3) >> 2) + 1 + ew : 0;
            lw = (my_n > 0 ? 1 + page + ((s_tw[tid] + 3) >> 2) : 0) + (a.app ? 1 : 0);
        }
        int Tw = 0;
        const int lex = cta_excl_scan(lw, s_scan, &Tw);
        T = Tw;
        if (tid < nl) s_lo[tid] = lex;
        if (tid == 0) s_lo[nl] = Tw;
        __syncthreads();
        fit = Tw <= FE_OCAP;
        if (fit) {
            if (tid == 0) st_relaxed_u64(a.state + batch, (batch == 0 ? kLbInc : kLbAgg) | (uint64_t)T);
            for (int i = tid; i < Tw; i += FE_THREADS) s_out[i] = 0u;
            // exceptions of my group, and of the groups before it in the block
            int my_e = 0;
            if (have) {
                const int b = s_bb[bi];
#pragma unroll
                for (int i = 0; i < 32; ++i) my_e += (b < 32 && (v[i] >> b) != 0) ? 1 : 0;
                s_ge[tid] = (uint8_t)my_e;
            }
            __syncthreads();
            if (have) {
                const int j = s_bl[bi], b = s_bb[bi], c = s_bc[bi], maxb = s_bm[bi];
                uint32_t *page = s_out + s_lo[j] + 1;
                // BitPacking.fastpack = masked (FastPFOR.java:175-179)
                uint32_t *o = page + 1 + s_bpw[bi] + g * b;
                const uint32_t mask = b >= 32 ? 0xffffffffu : ((1u << b) - 1u);
                if (b == 32) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = v[i];
                } else if (b > 0) {
                    uint64_t acc = 0;
                    int fill = 0;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        acc |= (uint64_t)(v[i] & mask) << fill;
                        fill += b;
                        if (fill >= 32) { *o++ = (uint32_t)acc; acc >>= 32; fill -= 32; }
                    }
                }
                if (c > 0 && my_e > 0) {
                    int rank = 0;
                    for (int gg = 0; gg < g; ++gg) rank += s_ge[bi * BG + gg];
                    const int k = maxb - b, where_meta = 1 + s_pw[j];
                    uint8_t *meta = reinterpret_cast<uint8_t *>(page + where_meta + 1) + s_bmo[bi] + 3;
                    uint32_t *sw = page + where_meta + 1 + ((s_mb[j] + 3) >> 2) + 1 + s_sbase[j][k < 2 ? 2 : k] + 1;
                    const int spos = s_bsp[bi];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const uint32_t hi = v[i] >> b; // b <= 31 whenever c > 0
                        if (hi != 0) {
                            meta[rank] = (uint8_t)(g * 32 + i); // :166
                            if (k >= 2) { // stream 1 is implicit (the decoder ORs 1 << b, :291-295)
                                const int bit = (spos + rank) * k, w = bit >> 5, sh = bit & 31;
                                atomicOr(sw + w, hi << sh);
                                if (sh + k > 32) atomicOr(sw + w + 1, hi >> (32 - sh));
                            }
                            ++rank;
                        }
                    }
                }
                if (g == 0) { // block metadata bytes (:159-165)
                    uint8_t *meta = reinterpret_cast<uint8_t *>(page + 1 + s_pw[j] + 1) + s_bmo[bi];
                    meta[0] = (uint8_t)b; meta[1] = (uint8_t)c;
                    if (c > 0) meta[2] = (uint8_t)maxb;
                }
            }
            // ---- one thread per list: header words of the list and of the page
            if (tid < nl && my_n > 0) {
                uint32_t *o = s_out + s_lo[tid];
                o[0] = (uint32_t)(my_nb * B); // Composition: FastPFOR's header, or the 0 marker
                if (my_nb) {
                    uint32_t *page = o + 1;
                    const int where_meta = 1 + s_pw[tid], mw = (s_mb[tid] + 3) >> 2;
                    page[0] = (uint32_t)where_meta;       // :182
                    page[where_meta] = (uint32_t)s_mb[tid]; // :186
                    uint32_t bitmap = 0;
                    for (int k = 2; k <= 32; ++k)
                        if (s_cntk[tid][k]) {
                            bitmap |= 1u << (k - 1);
                            page[where_meta + 1 + mw + 1 + s_sbase[tid][k]] = s_cntk[tid][k]; // :200
                        }
                    page[where_meta + 1 + mw] = bitmap; // :191-196
                }
            }
            // ---- VariableByte tail bytes
            for (int t = tid; t < ntv; t += FE_THREADS) {
                const int e = s_tt[t], j = e >> 11, off = e & 2047, i = t - s_tb[j];
                const int gidx = s_in[j] + ((s_in[j + 1] - s_in[j]) / B) * B + i;
                const uint32_t x = s_val[sw33(gidx)];
                uint32_t d = (a.delta && gidx > s_in[j]) ? x - s_val[sw33(gidx - 1)] : x;
                uint8_t *tbp = reinterpret_cast<uint8_t *>(s_out + (int)s_lo[j + 1] - (a.app ? 1 : 0) - ((s_tw[j] + 3) >> 2)) + off;
                while (d >= 128u) { *tbp++ = (uint8_t)(d & 127u); d >>= 7; }
                *tbp = (uint8_t)(d | 128u);
            }
        }
    }
    if (!fit) { // the warp-per-list kernels will redo the call: keep the look-back chain moving
        if (tid == 0) {
            *a.err = 2;
            st_relaxed_u64(a.state + batch, (batch == 0 ? kLbInc : kLbAgg) | 0ull);
        }
        T = 0;
    }
    __syncthreads();
    if (warp == 0) {
        const long long base = lookback_warp(a.state, batch, T, lane);
        if (lane == 0) {
            s_base = base;
            if (batch == a.nbatches - 1) {
                *a.total = 4 * (base + T);
                a.out_off[a.nlists] = 4 * (base + T);
            }
            if (4 * (base + T) > a.cap) atomicMax(a.err, 1);
        }
    }
    __syncthreads();
    if (!fit) return;
    const int64_t base = s_base;
    for (int j = tid; j < nl; j += FE_THREADS) a.out_off[l0 + j] = 4 * (base + s_lo[j]);
    if (4 * (base + T) > a.cap) return;
    uint32_t *gout = reinterpret_cast<uint32_t *>(a.out) + base;
    for (int i = tid; i < (int)T; i += FE_THREADS) gout[i] = a.app ? bswap32(s_out[i]) : s_out[i];
}

// ------------------------------------------------------------------------------------------------ FastPFOR, decode
// Inverse of k_fp_encode_fast for the same shapes (one page per list).  One thread per list parses the page header, the
// exception-stream table and the per-block byte metadata (the serial part: positions depend on the widths before
// them); one thread per 32-value group unpacks; the exceptions of a block are patched in by its group threads, strided;
// the VariableByte tail (up to block-1 values) goes through the warp-cooperative terminator-popcount routine; the
// Delta prefix is a three-level scan (inside the group, over a list's groups, over its tail).
struct SmemSrc {
    const uint32_t *w;
    __device__ __forceinline__ uint32_t get(int64_t j) const { return w[j]; }
};
struct SwzOut { // value i of a tail -> its 33-stride slot in the staged batch
    uint32_t *s_val;
    int base;
    __device__ __forceinline__ void operator()(int64_t i, uint32_t v) const { s_val[sw33(base + (int)i)] = v; }
};

template <int BG>
__global__ void __launch_bounds__(FE_THREADS, 6) k_fp_decode_fast(const FastDecArgs a)
{
    constexpr int B = BG * 32;
    __shared__ uint32_t s_val[FE_VPAD];
    __shared__ uint32_t s_wd[FE_OCAP];
    __shared__ int s_in[FP_MAXL + 1], s_oo[FP_MAXL + 1], s_fb[FP_MAXL + 1];
    __shared__ int s_page[FP_MAXL], s_meta[FP_MAXL], s_tp[FP_MAXL]; // word of the page / of its metadata bytes / of the VByte tail, in s_wd
    __shared__ uint16_t s_sbase[FP_MAXL][33];                       // first packed word of stream k, relative to the page
    __shared__ uint8_t s_bb[FP_MAXBLK], s_bc[FP_MAXBLK], s_bk[FP_MAXBLK], s_bl[FP_MAXBLK];
    __shared__ uint16_t s_bpw[FP_MAXBLK], s_bmo[FP_MAXBLK], s_bsp[FP_MAXBLK];
    __shared__ uint32_t s_gt[FE_THREADS]; // sum of each group's values (Delta)
    __shared__ uint32_t s_gc[FE_THREADS]; // carry into each group
    __shared__ uint32_t s_tc[FP_MAXL];    // carry into each tail
    __shared__ int s_scan[8];
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t l0 = (int64_t)blockIdx.x * a.lists_per_batch;
    const int nl = (int)min((int64_t)a.lists_per_batch, a.nlists - l0);
    const int64_t ibase = a.in_off[l0], obase = a.out_off[l0];
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int j = tid; j <= nl; j += FE_THREADS) {
        const int64_t ib = a.in_off[l0 + j] - ibase, ob = a.out_off[l0 + j] - obase;
        if ((a.in_off[l0 + j] & 3) != 0) s_bad = 1;
        s_in[j] = (int)((ib >> 2) > 0x7fffffff ? 0x7fffffff : (ib >> 2));
        s_oo[j] = (int)(ob > 0x7fffffff ? 0x7fffffff : ob);
    }
    __syncthreads();
    if (s_bad) { if (tid == 0) atomicExch(a.err, 1); return; }
    const int nw = s_in[nl], nv = s_oo[nl];
    int nblk = 0;
    const int my_n = tid < nl ? s_oo[tid + 1] - s_oo[tid] : 0;
    const int my_nb = my_n / B;
    const int my_fb = cta_excl_scan(my_nb, s_scan, &nblk);
    if (tid < nl) s_fb[tid] = my_fb;
    if (tid == 0) s_fb[nl] = nblk;
    if (nw > FE_OCAP || nv > FE_VCAP || nl > FP_MAXL || nblk > FP_MAXBLK) { // not this kernel's shape: the host reruns the call
        if (tid == 0) atomicMax(a.err, 2);
        return;
    }
    const uint32_t *gin = reinterpret_cast<const uint32_t *>(a.in + ibase);
    for (int i = tid; i < nw; i += FE_THREADS) { const uint32_t x = __ldg(gin + i); s_wd[i] = a.app ? bswap32(x) : x; }
    __syncthreads();
    // ---- one thread per list: page header, stream table, block metadata (FastPFOR.decodePage, FastPFOR.java:233-307)
    if (tid < nl) {
        const int avail = s_in[tid + 1] - s_in[tid] - (a.app ? 1 : 0); // the loader's trailing zero int is not part of the stream
        const uint32_t *w = s_wd + s_in[tid];
        bool ok = true;
        int p = 0;
        if (my_n > 0) {
            ok = avail >= 1 && w[0] == (uint32_t)(my_nb * B);
            p = 1;
            if (ok && my_nb > 0) {
                const uint32_t *page = w + 1;
                const int pav = avail - 1;
                const int where_meta = pav >= 3 ? (int)page[0] : -1;
                ok = where_meta >= 1 && where_meta < pav;
                int bytesize = 0, mw = 0, q = 0;
                if (ok) {
                    bytesize = (int)page[where_meta];
                    mw = (bytesize + 3) >> 2;
                    q = where_meta + 1 + mw;
                    ok = bytesize >= 0 && q < pav;
                }
                uint16_t ssize[33];
                if (ok) {
                    const uint32_t bitmap = page[q++];
                    for (int k = 2; k <= 32; ++k) {
                        ssize[k] = 0;
                        s_sbase[tid][k] = 0;
                        if (ok && (bitmap & (1u << (k - 1)))) {
                            if (q >= pav) { ok = false; break; }
                            const uint32_t size = page[q];
                            if (size > 65535u) { ok = false; break; }
                            ssize[k] = (uint16_t)size;
                            s_sbase[tid][k] = (uint16_t)(q + 1);
                            q += 1 + (int)((size * (uint32_t)k + 31u) >> 5);
                        }
                    }
                    ok = ok && q <= pav;
                }
                if (ok) {
                    const uint8_t *meta = reinterpret_cast<const uint8_t *>(page + where_meta + 1);
                    uint16_t cnt[33];
#pragma unroll
                    for (int k = 0; k < 33; ++k) cnt[k] = 0;
                    int mpos = 0, pw = 1;
                    for (int blk = 0; blk < my_nb; ++blk) {
                        const int b_ = my_fb + blk;
                        if (mpos + 2 > bytesize) { ok = false; break; }
                        const int b = meta[mpos], c = meta[mpos + 1];
                        if (b > 32 || pw + BG * b > where_meta) { ok = false; break; }
                        int k = 0;
                        if (c > 0) {
                            if (mpos + 3 + c > bytesize) { ok = false; break; }
                            k = meta[mpos + 2] - b;
                            if (k < 1 || k > 32 || b + k > 32) { ok = false; break; }
                            if (k >= 2 && cnt[k] + c > ssize[k]) { ok = false; break; }
                        }
                        s_bl[b_] = (uint8_t)tid; s_bb[b_] = (uint8_t)b; s_bc[b_] = (uint8_t)c; s_bk[b_] = (uint8_t)k;
                        s_bpw[b_] = (uint16_t)pw; s_bmo[b_] = (uint16_t)mpos; s_bsp[b_] = cnt[k];
                        pw += BG * b;
                        if (c > 0) { cnt[k] += (uint16_t)c; mpos += 3 + c; } else mpos += 2;
                    }
                    s_page[tid] = s_in[tid] + 1;
                    s_meta[tid] = s_in[tid] + 1 + where_meta + 1;
                    p = 1 + q;
                }
            }
        }
        s_tp[tid] = p;
        if (!ok) s_bad = 1;
    }
    __syncthreads();
    if (s_bad) { if (tid == 0) atomicExch(a.err, 1); return; }
    // ---- VariableByte tails: a warp per list
    for (int j = warp; j < nl; j += FE_THREADS / 32) {
        const int n = s_oo[j + 1] - s_oo[j], r = n - (n / B) * B;
        if (r == 0) continue;
        const int avail = s_in[j + 1] - s_in[j] - (a.app ? 1 : 0);
        const SmemSrc src{s_wd + s_in[j]};
        const SwzOut put{s_val, s_oo[j] + (n / B) * B};
        if (!warp_vb_decode_to<false>(src, s_tp[j], avail, r, put, 0u, lane) && lane == 0) s_bad = 1;
    }
    // ---- one thread per 32-value group: fastunpack
    const int bi = tid / BG, g = tid % BG;
    const bool have = bi < nblk;
    int my_start = 0;
    if (have) {
        const int j = s_bl[bi], b = s_bb[bi];
        my_start = s_oo[j] + (bi - s_fb[j]) * B + g * 32;
        const uint32_t *ptr = s_wd + s_page[j] + s_bpw[bi] + g * b;
        if (b == 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) s_val[sw33(my_start + i)] = ptr[i];
        } else {
            const uint32_t mask = (1u << b) - 1u;
            uint64_t acc = 0;
            int fill = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (fill < b) { acc |= (uint64_t)(*ptr++) << fill; fill += 32; }
                s_val[sw33(my_start + i)] = (uint32_t)acc & mask;
                acc >>= b;
                fill -= b;
            }
        }
    }
    __syncthreads();
    if (s_bad) { if (tid == 0) atomicExch(a.err, 1); return; }
    // ---- exceptions: the BG group threads of a block share its list, strided (FastPFOR.java:283-300)
    if (have && s_bc[bi] > 0) {
        const int j = s_bl[bi], b = s_bb[bi], c = s_bc[bi], k = s_bk[bi];
        const uint8_t *pos = reinterpret_cast<const uint8_t *>(s_wd + s_meta[j]) + s_bmo[bi] + 3;
        const uint32_t *sw = s_wd + s_page[j] + s_sbase[j][k < 2 ? 2 : k];
        const int blk_start = s_oo[j] + (bi - s_fb[j]) * B, spos = s_bsp[bi];
        for (int e = g; e < c; e += BG) {
            const int ps = pos[e];
            if (ps >= B) continue;
            uint32_t hi = 1u;
            if (k >= 2) {
                const int bit = (spos + e) * k, w = bit >> 5, sh = bit & 31;
                const uint32_t lo = sw[w], hw = (sh + k > 32) ? sw[w + 1] : 0u;
                hi = __funnelshift_r(lo, hw, sh);
                if (k < 32) hi &= (1u << k) - 1u;
            }
            s_val[sw33(blk_start + ps)] |= hi << b;
        }
    }
    __syncthreads();
    if (a.delta) { // Delta.inverseDelta over each list: inside the groups, over the groups, over the tail
        if (have) {
            uint32_t run = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) { run += s_val[sw33(my_start + i)]; s_val[sw33(my_start + i)] = run; }
            s_gt[tid] = run;
        }
        __syncthreads();
        if (tid < nl) {
            uint32_t carry = 0;
            for (int it = my_fb * BG; it < (my_fb + my_nb) * BG; ++it) { s_gc[it] = carry; carry += s_gt[it]; }
            s_tc[tid] = carry;
        }
        __syncthreads();
        if (have) {
            const uint32_t c = s_gc[tid];
#pragma unroll
            for (int i = 0; i < 32; ++i) s_val[sw33(my_start + i)] += c;
        }
        for (int j = warp; j < nl; j += FE_THREADS / 32) {
            const int n = s_oo[j + 1] - s_oo[j], m = (n / B) * B, r = n - m;
            uint32_t run = s_tc[j];
            for (int base = 0; base < r; base += 32) {
                const int i = base + lane, o = sw33(s_oo[j] + m + i);
                const uint32_t x = i < r ? s_val[o] : 0u;
                const uint32_t incl = warp_incl_scan(x, lane) + run;
                if (i < r) s_val[o] = incl;
                run = __shfl_sync(kFull, incl, 31);
            }
        }
        __syncthreads();
    }
    uint32_t *gout = a.out + obase;
    for (int i = tid; i < nv; i += FE_THREADS) gout[i] = s_val[sw33(i)];
}

// ============================================================ short-list path: one THREAD per list
// Lists of ~100 ints (BASELINE config C4: 1 % of 10 000 trials) spend the batched kernels above on bookkeeping: block
// items, tail tables, four CTA-wide scans and a dozen CTA barriers per batch.  Here a WARP owns 32 consecutive lists and
// a THREAD owns one list from its first value to its last word, so every piece of list state (position, width headers,
// the VariableByte byte accumulator, the running value of the Delta / integrated chain) is a register and nothing ever
// waits on another warp:
//   * values arrive in STAGES of 32 per list: the warp copies 128 contiguous bytes of list j with 4-byte cp.async
//     (LDGSTS) into a TRANSPOSED tile (value i of list j at word 33 i + j): coalesced in global memory, conflict-free
//     in shared memory both ways.  Block stages first (stage k = block k of every list that has one), then the
//     VariableByte tails of all lists together, so a warp runs ONE tail loop however the list lengths fall.  Two
//     tiles: stage s + 1 -- and the first stage of the warp's NEXT batch -- is in flight while stage s is packed;
//   * thread j packs its block (or VariableByte-encodes its tail values) into its column of a transposed output
//     area (word w of list j at 33 w + j), again conflict-free;
//   * the 32 list sizes are scanned with shuffles, the batch takes its place in the output by the same decoupled
//     look-back as above (one state word per warp batch; the next batch's first stage is requested BEFORE the
//     look-back, which hides most of the wait for the predecessors), and the lists leave with coalesced stores.
// Decode mirrors it: the NEXT batch's compressed streams stream into a second transposed area while this batch's are
// unpacked.  A batch with a list that does not fit its SL_OC-word column takes the warp-per-list routines inside the
// same kernel; only a FastPFOR list of a whole block or more raises flag 2 (the host then takes the batched kernels).
constexpr int SL_WARPS = 2;      // warps per CTA in the encoder (independent of each other)
constexpr int SL_OC = 56;        // words of one list's compressed stream that fit its column
constexpr int SL_CTAS = 7;       // resident encoder CTAs per SM (2 x (2 x 4.2 + 7.4) KB of shared memory each)
constexpr int SL_DCTAS = 11;     // resident decoder CTAs per SM (one warp, 4.2 + 2 x 7.4 KB each)
constexpr double SL_MAXAVG = 144.0; // lists averaging more values go to the thread-per-block kernels

__host__ __device__ inline bool sl_codec_ok(int codec, bool delta)
{
    const CodecTraits t = codec_traits(codec);
    return !(t.integrated && delta) && !t.xorbp && !t.optpfd; // (XorBinaryPacking / OptPFD: warp-per-list kernels only)
}

__device__ __forceinline__ void cp_async4(uint32_t *smem_dst, const uint32_t *gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 7-bit groups of d spread over the low bytes of a 64-bit word, terminator bit on byte len-1 (VariableByte.java:44-68)
__device__ __forceinline__ uint64_t vb_spread(uint32_t d, int len)
{
    const uint32_t lo = (d & 0x7fu) | ((d << 1) & 0x7f00u) | ((d << 2) & 0x7f0000u) | ((d << 3) & 0x7f000000u);
    return (((uint64_t)(d >> 28) << 32) | lo) | (0x80ull << (8 * (len - 1)));
}

// A list the thread-per-list kernels cannot hold, by one warp: the BinaryPacking family through warp_bp_list_encode,
// a list that is all VariableByte (FastPFOR below one block: Composition's 0 marker first; VariableByte alone) here.
template <bool WRITE>
__device__ __forceinline__ int64_t warp_short_list_encode(const CodecTraits tr, const ValSeq x, int64_t n, const WordSink sink, bool app,
                                                          uint32_t *stage, int lane)
{
    if (tr.blocks) return warp_bp_list_encode<WRITE>(tr, x, n, sink, app, stage, lane);
    int64_t pos = 0;
    if (tr.sk) { if (WRITE && lane == 0) sink.put(0, (uint32_t)n); pos = 1; } // IntCompressor: compressed[0] = n
    if (n > 0) {
        if ((tr.fastpfor || tr.xorbp || tr.optpfd) && !tr.sk) { if (WRITE && lane == 0) sink.put(0, 0u); pos = 1; } // Composition's 0 marker
        auto getv = [&](int64_t i) -> uint32_t { return x(i); };
        pos += warp_vb_encode<WRITE>(getv, n, sink, pos, stage, lane);
    }
    if (app) { if (WRITE && lane == 0) sink.put(pos, 0u); ++pos; }
    return pos;
}

__device__ __forceinline__ bool warp_short_list_decode(const CodecTraits tr, const WordSrc src, int64_t avail, int64_t n, uint32_t *out,
                                                       bool delta, int lane)
{
    if (tr.blocks) return warp_bp_list_decode(tr, src, avail, n, out, delta, lane);
    int64_t p = 0;
    if (tr.sk) { if (avail < 1 || (int64_t)src.get(0) != n) return false; p = 1; }
    if (n == 0) return true;
    if ((tr.fastpfor || tr.xorbp || tr.optpfd) && !tr.sk) { if (avail < 1 || src.get(0) != 0u) return false; p = 1; }
    if (!warp_vb_decode<false>(src, p, avail, n, out, 0u, lane)) return false;
    if (delta) { __syncwarp(); warp_inverse_delta(out, n, lane); }
    return true;
}

// Any list of a flat-kernel call by one warp, first-stage blocks included (a list of at least one FastPFOR / XorBinaryPacking /
// OptPFD block among lists that are otherwise all VariableByte: the 128-int codecs on ~100-value lists meet one every few
// hundred lists).  Same steps as k_encode.  `x` is the sequence the flat kernels code (gaps for XorBinaryPacking's integrated
// tail); fq: 68 ints of scratch; blk / tab: OptPFD's 256 words and its Simple16 tables.  *unsupported is raised for a multi-page FastPFOR list (the
// stale-buffer emulation lives in k_encode).
template <bool WRITE, bool OPT>
__device__ __forceinline__ int64_t warp_any_list_encode(const CodecTraits tr, const ValSeq x, int64_t n, const WordSink sink, bool app,
                                                        uint32_t *stage, int *fq, int32_t *blk, const uint8_t *tab, int lane, bool *unsupported)
{
    const int64_t blkn = tr.fastpfor ? tr.fp_block : 128;
    if (!(tr.fastpfor || tr.xorbp || tr.optpfd) || n < blkn) return warp_short_list_encode<WRITE>(tr, x, n, sink, app, stage, lane);
    const int64_t m = n - n % blkn;
    if (tr.fastpfor && m > kFpPage) { *unsupported = true; return 0; }
    if (WRITE && lane == 0) sink.put(0, (uint32_t)(tr.sk ? n : m)); // IntCompressor's n, or the first stage's count word
    int64_t pos = 1;
    if (tr.fastpfor) {
        const int64_t w = tr.fp_block == 256 ? warp_fastpfor_encode<WRITE, 8>(x, m, WRITE ? sink.w + pos : nullptr, fq, fq + 34, nullptr, lane)
                                             : warp_fastpfor_encode<WRITE, 4>(x, m, WRITE ? sink.w + pos : nullptr, fq, fq + 34, nullptr, lane);
        if (WRITE && app) { // pages are written as native words (byte-granular metadata): swapped afterwards
            __syncwarp();
            for (int64_t j = pos + lane; j < pos + w; j += 32) sink.w[j] = bswap32(sink.w[j]);
        }
        pos += w;
    } else if (tr.xorbp) {
        pos += warp_xor128_encode<WRITE>(ValSeq{x.in, false}, m, sink, pos, lane); // the blocks XOR the values themselves
    } else if (OPT) {
        pos += warp_optpfd_encode<WRITE>(x, m, sink, pos, blk, tab, lane);
    }
    if (n > m) {
        if (tr.xorbp) { // IntegratedVariableByte restarts at 0 behind the blocks (IntegratedVariableByte.java:40)
            const uint32_t *v = x.in;
            auto getv = [&](int64_t i) -> uint32_t { return i > 0 ? __ldg(v + m + i) - __ldg(v + m + i - 1) : __ldg(v + m); };
            pos += warp_vb_encode<WRITE>(getv, n - m, sink, pos, stage, lane);
        } else {
            auto getv = [&](int64_t i) -> uint32_t { return x(m + i); };
            pos += warp_vb_encode<WRITE>(getv, n - m, sink, pos, stage, lane);
        }
    }
    if (app) { if (WRITE && lane == 0) sink.put(pos, 0u); ++pos; }
    return pos;
}

// scr: cap words of shared memory (a big-endian-framed FastPFOR page is un-swapped into it); fq: 68 ints; ex: OptPFD's 284 words.
// `delta`: what the flat decoder undoes (true for XorBinaryPacking: its all-tail lists are gap-coded).
template <bool OPT>
__device__ __forceinline__ bool warp_any_list_decode(const CodecTraits tr, const WordSrc src, int64_t avail, int64_t n, uint32_t *out,
                                                     bool delta, uint32_t *scr, int cap, int *fq, uint32_t *ex, int lane, bool *unsupported)
{
    const int64_t blkn = tr.fastpfor ? tr.fp_block : 128;
    if (!(tr.fastpfor || tr.xorbp || tr.optpfd) || n < blkn) return warp_short_list_decode(tr, src, avail, n, out, delta, lane);
    const int64_t m = n - n % blkn;
    if (tr.fastpfor && m > kFpPage) { *unsupported = true; return true; }
    if (avail < 1 || (int64_t)src.get(0) != (tr.sk ? n : m)) return false;
    int64_t p = 1, used = -1;
    uint32_t carry = 0;
    if (tr.fastpfor) {
        if (src.swap) {
            const int64_t take = avail - p;
            if (take > cap) { *unsupported = true; return true; }
            for (int64_t j = lane; j < take; j += 32) scr[j] = src.get(p + j);
            __syncwarp();
            used = tr.fp_block == 256 ? warp_fp_decode_page<8, false>(scr, take, (int)m, out, fq, fq + 34, lane)
                                      : warp_fp_decode_page<4, false>(scr, take, (int)m, out, fq, fq + 34, lane);
        } else
            used = tr.fp_block == 256 ? warp_fp_decode_page<8, true>(src.w + p, avail - p, (int)m, out, fq, fq + 34, lane)
                                      : warp_fp_decode_page<4, true>(src.w + p, avail - p, (int)m, out, fq, fq + 34, lane);
    } else if (tr.xorbp) {
        used = warp_xor128_decode(src, p, avail, m, out, carry, lane);
    } else if (OPT) {
        used = warp_optpfd_decode(src, p, avail, m, out, ex, lane);
    }
    __syncwarp();
    if (used < 0) return false;
    p += used;
    if (n > m) {
        const bool ok = tr.xorbp ? warp_vb_decode<true>(src, p, avail, n - m, out + m, 0u, lane) // restarts at 0 (see the encoder)
                                 : warp_vb_decode<false>(src, p, avail, n - m, out + m, 0u, lane);
        if (!ok) return false;
    }
    if (delta && !tr.xorbp) { __syncwarp(); warp_inverse_delta(out, n, lane); }
    return true;
}

// what a lane knows about its list of the batch the warp is working on (or about to)
struct SlList {
    int64_t beg, n64; // first value (index into vals), length
    int off, n;       // first value relative to the batch's first value, length (0 when the batch takes the slow path)
    int nb, r;        // 32-int blocks of the first stage, values past them
};

__global__ void __launch_bounds__(SL_WARPS * 32, SL_CTAS) k_sl_encode(const FastEncArgs a)
{
    __shared__ uint32_t s_inw[SL_WARPS][2][32 * 33];
    __shared__ uint32_t s_outw[SL_WARPS][SL_OC * 33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *my_out = s_outw[warp] + lane; // word w of my list: my_out[33 * w]
    const CodecTraits tr = codec_traits(a.codec);
    const bool diff = a.delta || tr.integrated;

    // ---- a batch's extents, and the request for its first stage
    auto take_batch = [&](int64_t &batch, int &nl, SlList &L, int64_t &vbase, bool &slow, int &maxnb, int &nst) {
        unsigned long long tk = 0;
        if (lane == 0) tk = atomicAdd(a.ticket, 1ull);
        batch = (int64_t)__shfl_sync(kFull, tk, 0);
        nl = 0; slow = false; maxnb = 0; nst = 0; vbase = 0;
        L = SlList{0, 0, 0, 0, 0, 0};
        if (batch >= a.nbatches) return;
        const int64_t l0 = batch * 32;
        nl = (int)min((int64_t)32, a.nlists - l0);
        if (lane < nl) {
            L.beg = a.list_off ? a.list_off[l0 + lane] : (l0 + lane) * a.uniform_len;
            L.n64 = (a.list_off ? a.list_off[l0 + lane + 1] : (l0 + lane + 1) * a.uniform_len) - L.beg;
        }
        vbase = __shfl_sync(kFull, L.beg, 0);
        const bool big = lane < nl && (L.beg - vbase > 0x7fffff || L.n64 > 0x7fffff || (tr.fastpfor && L.n64 >= tr.fp_block));
        slow = __any_sync(kFull, big);
        if (!slow) {
            L.off = lane < nl ? (int)(L.beg - vbase) : 0; L.n = (int)L.n64;
            L.nb = tr.blocks ? L.n >> 5 : 0; L.r = L.n - 32 * L.nb;
            maxnb = __reduce_max_sync(kFull, L.nb);
            nst = maxnb + ((__reduce_max_sync(kFull, L.r) + 31) >> 5);
        }
    };
    // stage s of a batch: block s of every list that has one, then the values past the blocks, 32 per list and stage
    auto request_stage = [&](int s, const SlList &L, int maxnb, int64_t vbase, uint32_t *tile) {
        int src, cnt;
        if (s < maxnb) { src = L.off + 32 * s; cnt = s < L.nb ? 32 : 0; }
        else { const int t = s - maxnb; src = L.off + 32 * L.nb + 32 * t; cnt = min(32, max(0, L.r - 32 * t)); }
        const uint32_t *gv = a.vals + vbase + lane;
        const int sc = (src << 6) | cnt; // one shuffle per list (src < 2^25 on this path)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int scj = __shfl_sync(kFull, sc, j);
            if (lane < (scj & 63)) cp_async4(tile + lane * 33 + j, gv + (scj >> 6));
        }
        cp_async_commit();
    };

    int64_t batch, vbase;
    int nl, maxnb, nst;
    bool slow;
    SlList L;
    take_batch(batch, nl, L, vbase, slow, maxnb, nst);
    if (nst > 0) request_stage(0, L, maxnb, vbase, s_inw[warp][0]);
    while (batch < a.nbatches) {
        const int64_t l0 = batch * 32;
        bool ovf = slow;
        int pos = 0;
        if (!ovf) {
            const int n = L.n, nb = L.nb, ng4 = nb & ~3;
            // IntCompressor: compressed[0] = n; Composition: the first stage's header, 0 when it wrote nothing
            if (tr.sk) { my_out[0] = (uint32_t)n; pos = 1; }
            else if (n > 0 && (tr.fastpfor || tr.blocks)) { my_out[0] = (uint32_t)(32 * nb); pos = 1; }
            uint32_t prev = 0, hdrw = 0;
            int hpos = 0, vfill = 0;
            uint64_t vacc = 0;
            for (int s = 0; s < nst; ++s) {
                if (s + 1 < nst) { request_stage(s + 1, L, maxnb, vbase, s_inw[warp][(s + 1) & 1]); cp_async_wait<1>(); }
                else cp_async_wait<0>();
                __syncwarp();
                const uint32_t *col = s_inw[warp][s & 1] + lane; // value i of my list in this stage: col[33 * i]
                bool t_ovf = false;
                if (s < nb) {
                    // ---- one 32-int block: Util.maxbits / maxdiffbits, BitPacking.fastpackwithoutmask
                    uint32_t v[32];
                    uint32_t orv = 0;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const uint32_t x = col[33 * i];
                        v[i] = diff ? x - prev : x;
                        if (diff) prev = x;
                        orv |= v[i];
                    }
                    const int b = bits32(orv);
                    if (pos + 1 + b > SL_OC) t_ovf = true;
                    else {
                        if (s < ng4) { // BinaryPacking.java:59-78: one header word per four blocks
                            if ((s & 3) == 0) { hpos = pos++; hdrw = 0; }
                            hdrw |= (uint32_t)b << (24 - 8 * (s & 3));
                            if ((s & 3) == 3) my_out[33 * hpos] = hdrw;
                        } else
                            my_out[33 * pos++] = (uint32_t)b; // :79-86
                        uint32_t *o = my_out + 33 * pos;
                        if (b == 32) { // IntegratedBitPacking stores the raw values at width 32
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[33 * i] = tr.integrated ? col[33 * i] : v[i];
                        } else if (b > 0) {
                            uint64_t acc = 0;
                            int fill = 0;
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                acc |= (uint64_t)v[i] << fill;
                                fill += b;
                                if (fill >= 32) { *o = (uint32_t)acc; o += 33; acc >>= 32; fill -= 32; }
                            }
                        }
                        pos += b;
                    }
                } else if (s >= maxnb) {
                    // ---- VariableByte on the values past the last block (VariableByte.java:38-77); the non-skippable
                    // integrated composition restarts the chain at 0 (IntegratedVariableByte.java:40)
                    const int c = min(32, L.r - 32 * (s - maxnb));
                    if (tr.integrated && !tr.chain_tail && s == maxnb) prev = 0;
                    for (int i = 0; i < c; ++i) {
                        const uint32_t x = col[33 * i];
                        const uint32_t d = diff ? x - prev : x;
                        if (diff) prev = x;
                        const int len = vb_len(d);
                        vacc |= vb_spread(d, len) << (8 * vfill);
                        vfill += len;
                        if (vfill >= 4) {
                            if (pos < SL_OC) my_out[33 * pos] = (uint32_t)vacc; else t_ovf = true;
                            ++pos; vacc >>= 32; vfill -= 4;
                            if (vfill >= 4) { // only after a 5-byte value
                                if (pos < SL_OC) my_out[33 * pos] = (uint32_t)vacc; else t_ovf = true;
                                ++pos; vacc >>= 32; vfill -= 4;
                            }
                        }
                    }
                }
                ovf = __any_sync(kFull, t_ovf);
                __syncwarp(); // this stage's tile has been read: it may be requested again
                if (ovf) break;
            }
            cp_async_wait<0>();
            bool t_ovf = false;
            if (vfill > 0) { // zero-padded to a word (VariableByte.java:69-70)
                if (pos < SL_OC) my_out[33 * pos] = (uint32_t)vacc; else t_ovf = true;
                ++pos;
            }
            if (a.app) { // LoadPortfolios.java:622: outpos + 1 ints
                if (pos < SL_OC) my_out[33 * pos] = 0u; else t_ovf = true;
                ++pos;
            }
            const bool late = __any_sync(kFull, t_ovf);
            ovf = ovf || late;
        }
        long long lw = lane < nl ? pos : 0; // (lanes past the last list carry an empty list's framing words)
        bool emit_slow = false;
        if (ovf) {
            lw = 0;
            if (__any_sync(kFull, lane < nl && tr.fastpfor && L.n64 >= tr.fp_block)) {
                if (lane == 0) atomicMax(a.err, 2); // a FastPFOR page: the host takes the batched kernels
            } else { // sizes by the warp-per-list routines, one list at a time
                emit_slow = true;
                for (int j = 0; j < nl; ++j) {
                    const int64_t bj = __shfl_sync(kFull, L.beg, j), nj = __shfl_sync(kFull, L.n64, j);
                    __syncwarp();
                    const int64_t w = warp_short_list_encode<false>(tr, ValSeq{a.vals + bj, a.delta}, nj, WordSink{nullptr, a.app}, a.app,
                                                                    s_inw[warp][0], lane);
                    if (lane == j) lw = w;
                }
                __syncwarp();
            }
        }
        // ---- place of every list inside the batch; publish the batch's size
        long long incl = lw;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long t = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += t;
        }
        const long long T = __shfl_sync(kFull, incl, 31), lex = incl - lw;
        if (lane == 0) st_relaxed_u64(a.state + batch, (batch == 0 ? kLbInc : kLbAgg) | (uint64_t)T);
        // ---- the next batch: its ticket, its extents and its first stage go out before this batch's look-back
        const int64_t c_batch = batch, c_beg = L.beg, c_n64 = L.n64;
        const int c_nl = nl;
        take_batch(batch, nl, L, vbase, slow, maxnb, nst);
        if (nst > 0) request_stage(0, L, maxnb, vbase, s_inw[warp][0]);
        // ---- decoupled look-back, 32 predecessors per round
        const long long base = lookback_warp(a.state, c_batch, T, lane);
        if (lane < c_nl) a.out_off[l0 + lane] = 4 * (base + lex);
        if (lane == 0 && c_batch == a.nbatches - 1) {
            *a.total = 4 * (base + T);
            a.out_off[a.nlists] = 4 * (base + T);
        }
        if (4 * (base + T) > a.cap) { // caller's buffer too small: report, write nothing
            if (lane == 0) atomicMax(a.err, 1);
            __syncwarp();
            continue;
        }
        uint32_t *gout = reinterpret_cast<uint32_t *>(a.out) + base;
        if (!ovf) {
            const int ilw = (int)lw, ilex = (int)lex;
            const uint32_t *so = s_outw[warp];
            const uint32_t sel = a.app ? 0x0123u : 0x3210u;
            __syncwarp();
#pragma unroll 8
            const int ow = (ilex << 7) | ilw; // one shuffle per list (ilex <= 32 SL_OC, ilw <= SL_OC)
            for (int j = 0; j < 32; ++j) { // at most two words per lane and list (SL_OC <= 64)
                const int owj = __shfl_sync(kFull, ow, j);
                const int wj = owj & 127, oj = owj >> 7;
                if (lane < wj) gout[oj + lane] = __byte_perm(so[33 * lane + j], 0, sel);
                if (lane + 32 < wj) gout[oj + lane + 32] = __byte_perm(so[33 * (lane + 32) + j], 0, sel);
            }
        } else if (emit_slow) {
            // (the next batch's first stage may be landing in tile 0: the slow path stages in tile 1)
            for (int j = 0; j < c_nl; ++j) {
                const int64_t bj = __shfl_sync(kFull, c_beg, j), nj = __shfl_sync(kFull, c_n64, j);
                const long long oj = __shfl_sync(kFull, lex, j);
                __syncwarp();
                warp_short_list_encode<true>(tr, ValSeq{a.vals + bj, a.delta}, nj, WordSink{gout + oj, a.app}, a.app, s_inw[warp][1], lane);
            }
        }
        __syncwarp(); // the columns have been read: the next batch may write them
    }
    cp_async_wait<0>();
}

__global__ void __launch_bounds__(32, SL_DCTAS) k_sl_decode(const FastDecArgs a, int64_t nbatches)
{
    __shared__ uint32_t s_val[32 * 33];
    __shared__ uint32_t s_wdw[2][SL_OC * 33];
    const int lane = threadIdx.x;
    const CodecTraits tr = codec_traits(a.codec);
    const bool diff = a.delta || tr.integrated;
    const uint32_t sel = a.app ? 0x0123u : 0x3210u; // big-endian framing: words are swapped as they are read

    struct Ext { int64_t ib, ob, aw, n64; bool bad, big; };
    // a batch's extents, and the request for its streams (transposed: word w of list j at 33 w + j)
    auto prefetch = [&](int64_t batch, Ext &x, uint32_t *wd) {
        x = Ext{0, 0, 0, 0, false, false};
        if (batch < nbatches) {
            const int64_t l0 = batch * 32;
            const int nl = (int)min((int64_t)32, a.nlists - l0);
            int64_t ie = 0;
            if (lane < nl) {
                x.ib = a.in_off[l0 + lane]; ie = a.in_off[l0 + lane + 1];
                x.ob = a.out_off[l0 + lane]; x.n64 = a.out_off[l0 + lane + 1] - x.ob;
            }
            x.bad = __any_sync(kFull, ((x.ib | ie) & 3) != 0 || ie < x.ib || x.n64 < 0);
            x.aw = (ie - x.ib) >> 2;
            const int64_t obase = __shfl_sync(kFull, x.ob, 0);
            const int64_t ibase0 = __shfl_sync(kFull, x.ib, 0);
            x.big = __any_sync(kFull, x.aw > SL_OC || x.n64 > 32 * SL_OC || x.ob - obase > 32 * 32 * SL_OC || x.ib - ibase0 > 4 * 32 * SL_OC ||
                                          (tr.fastpfor && x.n64 >= tr.fp_block));
            if (!x.bad && !x.big) {
                const int64_t ibase = __shfl_sync(kFull, x.ib, 0);
                const int ra = ((int)((x.ib - ibase) >> 2) << 7) | (int)x.aw; // one shuffle per list: (first word << 7) | words
                const uint32_t *g0 = reinterpret_cast<const uint32_t *>(a.in + ibase) + lane;
#pragma unroll 8
                for (int j = 0; j < 32; ++j) {
                    const int raj = __shfl_sync(kFull, ra, j);
                    const int awj = raj & 127;
                    const uint32_t *g = g0 + (raj >> 7);
                    if (lane < awj) cp_async4(wd + 33 * lane + j, g);
                    if (lane + 32 < awj) cp_async4(wd + 33 * (lane + 32) + j, g + 32);
                }
            }
        }
        cp_async_commit();
    };

    const int64_t stride = gridDim.x;
    int64_t batch = blockIdx.x;
    Ext cur, nxt;
    prefetch(batch, cur, s_wdw[0]);
    for (int it = 0; batch < nbatches; batch += stride, ++it, cur = nxt) {
        __syncwarp(); // the previous batch's area has been read: the next batch's streams may land in it
        prefetch(batch + stride, nxt, s_wdw[(it + 1) & 1]);
        cp_async_wait<1>();
        __syncwarp();
        const int64_t l0 = batch * 32;
        const int nl = (int)min((int64_t)32, a.nlists - l0);
        if (cur.bad) { if (lane == 0) atomicMax(a.err, 1); continue; }
        if (cur.big) {
            if (__any_sync(kFull, tr.fastpfor && cur.n64 >= tr.fp_block)) { if (lane == 0) atomicMax(a.err, 2); continue; } // the host takes the batched kernels
            for (int j = 0; j < nl; ++j) {
                const int64_t ibj = __shfl_sync(kFull, cur.ib, j), awj = __shfl_sync(kFull, cur.aw, j);
                const int64_t obj = __shfl_sync(kFull, cur.ob, j), nj = __shfl_sync(kFull, cur.n64, j);
                const WordSrc src{reinterpret_cast<const uint32_t *>(a.in + ibj), a.app};
                const bool ok = warp_short_list_decode(tr, src, awj, nj, a.out + obj, a.delta, lane);
                if (!ok && lane == 0) atomicMax(a.err, 1);
                __syncwarp();
            }
            continue;
        }
        const int64_t obase = __shfl_sync(kFull, cur.ob, 0);
        const int avail = (int)cur.aw, n = (int)cur.n64, oo = (int)(cur.ob - obase);
        const int on = (oo << 12) | n; // n <= 32 SL_OC = 1792, oo <= 32 x 1792
        const uint32_t *my = s_wdw[it & 1] + lane; // word p of my list's stream: my[33 * p]
        auto rd = [&](int p) -> uint32_t { return __byte_perm(my[33 * p], 0, sel); };
        const int nb = tr.blocks ? n >> 5 : 0, ng4 = nb & ~3;
        int p = 0;
        bool ok = true;
        if (tr.sk) { ok = lane >= nl || (avail >= 1 && rd(0) == (uint32_t)n); p = 1; }
        else if (n > 0 && (tr.fastpfor || tr.blocks)) { ok = avail >= 1 && rd(0) == (uint32_t)(32 * nb); p = 1; }
        uint32_t carry = 0, hdr = 0, vv = 0;
        int vs = 0, vshift = 0;
        const int maxn = __reduce_max_sync(kFull, n);
        uint32_t *gout = a.out + obase;
        for (int k = 0; 32 * k < maxn; ++k) {
            const int c = min(32, n - 32 * k);
            uint32_t *col = s_val + lane; // value i of my list in this round: col[33 * i]
            if (ok && k < nb) {
                // ---- one block: width from the group header (BinaryPacking.java:102-123) or its own (:124-131)
                int b = 33;
                if (k < ng4) {
                    if ((k & 3) == 0) { if (p < avail) hdr = rd(p++); else ok = false; }
                    b = (int)((hdr >> (24 - 8 * (k & 3))) & 255u);
                } else if (p < avail)
                    b = (int)min(rd(p++), 33u);
                if (b > 32 || p + b > avail) ok = false;
                if (ok) {
                    uint32_t v[32];
                    const uint32_t *ptr = my + 33 * p;
                    if (b == 32) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __byte_perm(ptr[33 * i], 0, sel);
                    } else {
                        const uint32_t mask = (1u << b) - 1u;
                        uint64_t acc = 0;
                        int fill = 0;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (fill < b) { acc |= (uint64_t)__byte_perm(*ptr, 0, sel) << fill; ptr += 33; fill += 32; }
                            v[i] = (uint32_t)acc & mask;
                            acc >>= b;
                            fill -= b;
                        }
                    }
                    const bool raw = tr.integrated && b == 32; // IntegratedBitPacking stores raw values at width 32
                    if (diff && !raw) {
                        v[0] += carry;
#pragma unroll
                        for (int i = 1; i < 32; ++i) v[i] += v[i - 1];
                    }
                    carry = v[31];
#pragma unroll
                    for (int i = 0; i < 32; ++i) col[33 * i] = v[i];
                    p += b;
                }
            } else if (ok && c > 0) {
                // ---- VariableByte.headlessUncompress (VariableByte.java:180-203); IntegratedVariableByte.java:40 restarts
                if (tr.integrated && !tr.chain_tail && k == nb) carry = 0;
                int got = 0;
                uint32_t w = p < avail ? rd(p) : 0u;
                auto one_byte = [&](uint32_t by) {
                    vv += (by & 127u) << (vshift & 31);
                    if (by & 128u) {
                        if (diff) { carry += vv; vv = carry; }
                        col[33 * got] = vv;
                        ++got; vv = 0; vshift = 0;
                    } else
                        vshift += 7;
                };
                while (got < c) {
                    if (p >= avail) { ok = false; break; }
                    if (vs == 0 && got + 4 <= c) {
                        // a whole word at a time while all of its (at most four) values belong to this round: no
                        // per-byte position bookkeeping
                        one_byte(w & 0xffu);
                        one_byte((w >> 8) & 0xffu);
                        one_byte((w >> 16) & 0xffu);
                        one_byte(w >> 24);
                        ++p; w = p < avail ? rd(p) : 0u;
                        continue;
                    }
                    const uint32_t by = (w >> vs) & 0xffu;
                    vs += 8;
                    if (vs == 32) { vs = 0; ++p; w = p < avail ? rd(p) : 0u; }
                    one_byte(by);
                }
            }
            __syncwarp();
            // ---- the round's values leave: 128 contiguous bytes per list
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int onj = __shfl_sync(kFull, on, j); // (oo << 12) | n: one shuffle per list
                const int idx = 32 * k + lane;
                if (idx < (onj & 4095)) gout[(onj >> 12) + idx] = s_val[33 * lane + j];
            }
            __syncwarp();
        }
        if (__any_sync(kFull, !ok) && lane == 0) atomicMax(a.err, 1);
    }
    cp_async_wait<0>();
}

// ================================================================================================
// Flat (lane-parallel) VariableByte kernels: lists that are ALL VariableByte -- Composition(FastPFOR | FastPFOR128,
// VariableByte) on lists shorter than one block (Composition writes the 0 marker and hands the whole list to
// VariableByte, FastPFOR.java:311-313, Composition.java:45-48: BASELINE config C4 under the north-star codec), the
// skippable forms of the same, and VariableByte alone.
//
// The thread-per-list kernels above walk a list serially (~45 instructions per value, a warp as slow as its longest
// list).  Here a CTA owns a TILE of consecutive whole lists and works on the tile's values as ONE flat array, one tile per CTA,
// with no state shared between CTAs:
//   encode  k_vb4_size (words per tile) -> k_vb4_scan (one CTA: where every tile goes) -> k_vb4_tile<.., EMIT> (the bytes);
//   decode  k_vb4_decode (the output offsets are an input: one launch).
// Each kernel's comment says how.  A tile that does not fit its window / staging areas is done by the warp-per-list routines
// inside the same kernel; a FastPFOR list of a whole block or more raises flag 2 and the host takes the batched kernels.
__device__ __forceinline__ void cp_async16(uint32_t *smem_dst, const uint32_t *gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}

// ---- flat VariableByte encoder: size pass, tile scan, emit pass.
// History, by measurement (profiles/r2/codec_flat_log.md).  Every single-pass version -- thread-per-list batches, flat tiles
// of 2 048 and 4 096 values, 32-, 256- and 768-wide look-backs, dynamic tickets and static assignment -- ended at ~0.5 ms for
// the 10^8 ints of C4: with tiles this small (a tile's time on an SM is about one L2 round trip) the decoupled look-back is
// the clock of the kernel: the frontier of known output positions advances by one poll window per round trip, and a grid in
// lock step waits for its slowest member at every tile (47 % of the stall samples at the barrier behind the look-back).
// So the output position of a tile comes from a SIZE pass instead: k_vb4_size (below; this kernel with EMIT = false computes the
// same number from its own prefixes -- it was the first size pass, 82 us instead of 75) writes the tile's word count, one CTA scans the
// ~26 000 counts, and the EMIT pass knows where it writes before it starts -- no state
// words, no polling, no residency requirement: one tile per CTA, the hardware scheduler does the balancing, and the input is
// read twice (HBM is not the bound here).
// The tile kernel: THREADS x VPT consecutive values per tile (a fixed number of whole lists); values arrive by coalesced 16-byte
// cp.async into a chunk-swizzled staging tile (conflict-free for the copy and for the LDS.128 of a thread's row), are turned
// into gaps once and stay in registers.  A thread whose values are all below 2^14 (one or two bytes: every thread of C4)
// keeps a "two bytes" bit mask instead of lengths; positions of the window outside the tile are phantom zeros whose bytes go
// to a scratch area, so no thread carries validity tests.  List bookkeeping is PULLED by one thread per list (byte prefix at
// a list's first value = its thread's prefix + a popcount), value threads never loop over list starts.
template <int THREADS, int VPT>
struct V4 {
    static constexpr int WARPS = THREADS / 32;
    static constexpr int SPAN = THREADS * VPT;           // window positions per tile
    static constexpr int CH = VPT / 4;                   // 16-byte chunks per thread
    static constexpr int MAXL = THREADS - 1;             // lists per tile: one thread per list, one more for the end offset
    static constexpr int OCAP = SPAN / 2 + 3 * THREADS;  // staged output words: two bytes per value + framing of MAXL lists
    static constexpr int SCR = VPT / 4 + 2;              // words of scratch for phantom bytes (leading and trailing)
    // word of chunk c of thread t in the staging tile: chunks permuted so that a quarter-warp's 16-byte accesses (row stride
    // 64 or 128 bytes) fall into eight different bank groups
    __device__ static __forceinline__ int slot(int t, int c) { return t * VPT + 4 * (VPT == 16 ? (c ^ ((t >> 1) & 3)) : (c ^ (t & 7))); }
};

__device__ __forceinline__ long long warp_sum_i64(long long v)
{
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
    return v;
}

// sum of the nibbles of x below nibble k (k <= 16)
__device__ __forceinline__ int nib_prefix64(uint64_t x, int k)
{
    if (k < 16) x &= (1ull << (4 * k)) - 1ull;
    x = (x & 0x0f0f0f0f0f0f0f0full) + ((x >> 4) & 0x0f0f0f0f0f0f0f0full);
    return (int)((x * 0x0101010101010101ull) >> 56);
}

struct TileEncArgs {
    FastEncArgs f;
    int64_t *tile_words;  // [nbatches] size pass: words of the tile; after the scan: first word of the tile in the output
    int hdr_mode;         // word in front of a list's bytes: 0 none, 1 IntCompressor's n, 2 Composition's 0 marker (non-empty lists)
    int page;             // FastPFOR block size: a list this long is not all VariableByte (0: no limit)
    int *defer_cnt;       // tiles that hold such a list are left to k_vb4_pages (null: flag 2, the host takes other kernels)
    int *defer_tiles;
};

template <int THREADS, int VPT, bool DELTA, bool EMIT>
__global__ void __launch_bounds__(THREADS) k_vb4_tile(const TileEncArgs ta)
{
    using C = V4<THREADS, VPT>;
    const FastEncArgs &a = ta.f;
    __shared__ __align__(16) uint32_t s_vals[C::SPAN];
    __shared__ __align__(16) uint32_t s_out[EMIT ? C::OCAP + 4 + 2 * C::SCR : 64 * C::WARPS];
    __shared__ int s_rel[THREADS + 1];              // list starts as window positions ([nl] = the tile's end)
    __shared__ uint32_t s_flag[C::SPAN / 32 + 1];   // bit p: a non-empty list starts at window position p (and the phantom one at the end)
    __shared__ uint32_t s_tp[THREADS];              // per thread: (bytes | list starts << 16) of the warp's lower lanes
    __shared__ uint32_t s_m2[THREADS];              // per thread: bit k = value k takes two bytes
    __shared__ uint64_t s_len[THREADS][VPT / 16];   // per thread off the short path: byte length of value k in nibble k
    __shared__ uint8_t s_gen[THREADS];              // per thread: 1 = off the short path
    __shared__ int s_ob[THREADS + 2];               // [r + 1], r-th non-empty list: byte address of its payload in s_out minus its flat byte prefix
    __shared__ uint32_t s_w1[C::WARPS], s_w2[C::WARPS];
    __shared__ long long s_red[C::WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long tile = blockIdx.x;
    const int64_t l0 = tile * a.lists_per_batch;
    const int nl = (int)min((int64_t)a.lists_per_batch, a.nlists - l0);
    long long gbase = 0;
    if (EMIT) { // (uniform) the caller's buffer is too small: the scan kernel has reported it, nothing is written
        if (*a.total > a.cap) return;
        gbase = ta.tile_words[tile];
    }
    // ---- the tile's extent, this thread's list
    int64_t base, end, my_off = 0, my_end = 0;
    if (a.list_off) {
        const int64_t *lo = a.list_off + l0;
        base = lo[0]; end = lo[nl];
        if (tid <= nl) { my_off = lo[tid]; my_end = lo[tid < nl ? tid + 1 : tid]; }
    } else {
        base = l0 * a.uniform_len; end = base + nl * a.uniform_len;
        my_off = base + min(tid, nl) * a.uniform_len; my_end = base + min(tid + 1, nl) * a.uniform_len;
    }
    const int64_t nv = end - base;
    const int a0 = (int)((reinterpret_cast<uintptr_t>(a.vals + base) >> 2) & 3); // the window starts on a 16-byte boundary
    bool bad = nv < 0 || a0 + nv > C::SPAN;
    const int span = bad ? 0 : a0 + (int)nv;
    if (tid < C::SPAN / 32 + 1) s_flag[tid] = 0;
    {
        const uint32_t *wp = a.vals + base - a0 + 4 * tid;
#pragma unroll
        for (int i = 0; i < C::CH; ++i) { // chunk q = i THREADS + tid of the window: coalesced
            const int q = i * THREADS + tid;
            if (4 * q < span) cp_async16(s_vals + C::slot(q / C::CH, q % C::CH), wp + 4 * i * THREADS);
        }
    }
    cp_async_commit();
    const int64_t my_n = my_end - my_off;
    if (tid <= nl) bad = bad || my_off < base || my_off > end || my_n < 0 || (ta.page > 0 && my_n >= ta.page);
    __syncthreads(); // (the start bits are zero)
    if (tid <= nl && !bad) {
        const int rel = a0 + (int)(my_off - base);
        s_rel[tid] = rel;
        // a value belongs to the LAST list that starts at its position: only non-empty lists set a bit; the window positions
        // past the tile's end form a phantom list of their own
        if (tid == nl || my_n > 0) atomicOr(&s_flag[rel >> 5], 1u << (rel & 31));
    }
    cp_async_wait<0>();
    int slow = __syncthreads_or(bad); // values, list starts and start bits are in place
    if (!slow) {
        // ---- pass A: the sequence the codec sees (Delta.delta: gaps, a list's first value as it is), byte lengths
        const int p0 = VPT * tid;
        const uint32_t fw = VPT == 32 ? s_flag[tid] : (s_flag[tid >> 1] >> (16 * (tid & 1))) & 0xffffu;
        const uint32_t full = VPT == 32 ? 0xffffffffu : 0xffffu;
        const int lo = a0 - p0, hi = min(VPT, span - p0);
        const uint32_t vm = hi <= 0 ? 0u : ((hi >= 32 ? 0xffffffffu : (1u << hi) - 1u) & (lo <= 0 ? full : ~((1u << lo) - 1u)));
        uint32_t d[VPT];
        uint32_t m2 = 0, wor = 0;
        uint64_t nib[VPT / 16];
        bool gen = false;
        int nbytes = 0;
        if (vm) {
            uint32_t prev = tid > 0 ? s_vals[C::slot(tid - 1, C::CH - 1) + 3] : 0u;
#pragma unroll
            for (int c = 0; c < C::CH; ++c) {
                const uint4 q = *reinterpret_cast<const uint4 *>(s_vals + C::slot(tid, c));
                const uint32_t x[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int pk = 4 * c + k;
                    d[pk] = (DELTA && !((fw >> pk) & 1u)) ? x[k] - prev : x[k];
                    prev = x[k];
                }
            }
            // phantom positions (before the tile's first value: at most three; past its last): zeros, one byte each
            if (lo > 0) {
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (k < lo) d[k] = 0u;
            }
            if (hi < VPT) {
#pragma unroll
                for (int k = 0; k < VPT; ++k)
                    if (k >= hi) d[k] = 0u;
            }
#pragma unroll
            for (int k = 0; k < VPT; ++k) wor |= d[k];
            gen = wor >= 16384u;
            if (!gen) {
#pragma unroll
                for (int k = 0; k < VPT; ++k) m2 |= (d[k] >= 128u ? 1u : 0u) << k;
                nbytes = VPT + __popc(m2);
            } else {
#pragma unroll
                for (int h = 0; h < VPT / 16; ++h) {
                    uint64_t nb64 = 0;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int len = vb_len(d[16 * h + k]);
                        nb64 |= (uint64_t)len << (4 * k);
                        nbytes += len;
                    }
                    nib[h] = nb64;
                    s_len[tid][h] = nb64;
                }
            }
        }
        s_m2[tid] = m2;
        s_gen[tid] = gen ? 1 : 0;
        if (!EMIT && VPT == 32) {
            // the size pass needs no prefixes: when every value of the tile takes one or two bytes (the "two bytes" masks
            // are then a bit vector over the window), a list's bytes are its length plus a popcount over its range
            if (!__syncthreads_or(gen)) {
                int words = 0;
                if (tid < nl) {
                    const int n = (int)my_n;
                    int bytes = n;
                    if (n > 0) {
                        const int plo = s_rel[tid], phi = plo + n, wlo = plo >> 5, whi = (phi - 1) >> 5;
                        for (int w = wlo; w <= whi; ++w) {
                            uint32_t m = s_m2[w];
                            if (w == wlo) m &= 0xffffffffu << (plo & 31);
                            if (w == whi && (phi & 31)) m &= (1u << (phi & 31)) - 1u;
                            bytes += __popc(m);
                        }
                    }
                    const int hdr = ta.hdr_mode == 1 ? 1 : (ta.hdr_mode == 2 && n > 0 ? 1 : 0);
                    words = hdr + ((bytes + 3) >> 2) + (a.app ? 1 : 0);
                }
                words = __reduce_add_sync(kFull, words);
                if (lane == 0) s_w2[warp] = (uint32_t)words;
                __syncthreads();
                if (tid == 0) {
                    long long T = 0;
#pragma unroll
                    for (int w = 0; w < C::WARPS; ++w) T += s_w2[w];
                    ta.tile_words[tile] = T;
                }
                return;
            }
        }
        const uint32_t my1 = (uint32_t)nbytes | ((uint32_t)__popc(fw) << 16);
        uint32_t incl1 = my1;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, incl1, dd);
            if (lane >= dd) incl1 += t;
        }
        s_tp[tid] = incl1 - my1;
        if (lane == 31) s_w1[warp] = incl1;
        __syncthreads();
        uint32_t wo1 = 0, tot1 = 0;
#pragma unroll
        for (int w = 0; w < C::WARPS; ++w) {
            const uint32_t x = s_w1[w];
            if (w < warp) wo1 += x;
            tot1 += x;
        }
        const uint32_t ex1 = wo1 + incl1 - my1;
        // ---- one thread per list: its bytes (flat byte prefix at the next list's first value minus the one at its own),
        // words and framing: IntCompressor writes n first, Composition the 0 marker when FastPFOR took nothing, the
        // loader's framing one trailing zero word (LoadPortfolios.java:622)
        auto byte_prefix = [&](int p) -> int {
            if (p >= C::SPAN) return (int)(tot1 & 0xffffu);
            const int t = p / VPT, k = p % VPT;
            int b;
            if (s_gen[t]) {
                b = 0;
#pragma unroll
                for (int h = 0; h < VPT / 16; ++h)
                    if (k > 16 * h) b += nib_prefix64(s_len[t][h], min(16, k - 16 * h));
            } else
                b = k + __popc(s_m2[t] & ((1u << k) - 1u));
            b += (int)(s_tp[t] & 0xffffu);
#pragma unroll
            for (int w = 0; w < C::WARPS - 1; ++w)
                if (w < (t >> 5)) b += (int)(s_w1[w] & 0xffffu);
            return b;
        };
        int n = 0, bytes = 0, hdr = 0, words = 0, bplo = 0;
        if (tid <= nl) {
            const int plo = s_rel[tid];
            bplo = byte_prefix(plo);
            if (tid < nl) {
                n = (int)my_n;
                bytes = n > 0 ? byte_prefix(plo + n) - bplo : 0;
                hdr = ta.hdr_mode == 1 ? 1 : (ta.hdr_mode == 2 && n > 0 ? 1 : 0);
                words = hdr + ((bytes + 3) >> 2) + (a.app ? 1 : 0);
            }
        }
        const uint32_t my2 = (uint32_t)words | ((n > 0 ? 1u : 0u) << 16);
        uint32_t incl2 = my2;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, incl2, dd);
            if (lane >= dd) incl2 += t;
        }
        if (lane == 31) s_w2[warp] = incl2;
        __syncthreads();
        uint32_t wo2 = 0, tot2 = 0;
#pragma unroll
        for (int w = 0; w < C::WARPS; ++w) {
            const uint32_t x = s_w2[w];
            if (w < warp) wo2 += x;
            tot2 += x;
        }
        const int T = (int)(tot2 & 0xffffu); // words of the tile
        if (T > C::OCAP) slow = 1; // (uniform) the tile does not fit the staging area
        else if (!EMIT) {
            if (tid == 0) ta.tile_words[tile] = T;
            return;
        } else {
            // the tile is staged at the word offset its first word has inside a 16-byte line of the output, so that it leaves
            // with aligned 16-byte loads and stores
            uint32_t *gout = reinterpret_cast<uint32_t *>(a.out) + gbase;
            const int sh = (int)((reinterpret_cast<uintptr_t>(gout) >> 2) & 3);
            const uint32_t ex2 = wo2 + incl2 - my2;
            const int Wj = (int)(ex2 & 0xffffu); // first word of list `tid` inside the tile
            uint8_t *sb = reinterpret_cast<uint8_t *>(s_out);
            if (tid < nl) {
                uint32_t *lw = s_out + sh + Wj;
                if (hdr) lw[0] = ta.hdr_mode == 1 ? (uint32_t)n : 0u;
                if (bytes > 0) lw[hdr + ((bytes - 1) >> 2)] = 0u; // the zero padding of the last payload word
                if (a.app) lw[words - 1] = 0u;
                if (n > 0) s_ob[(ex2 >> 16) + 1] = 4 * (sh + Wj + hdr) - bplo;
                a.out_off[l0 + tid] = 4 * (gbase + Wj);
            } else if (tid == nl)
                s_ob[(ex2 >> 16) + 1] = 4 * (C::OCAP + 4 + C::SCR) - bplo; // the phantom list past the tile's end: its bytes go to scratch
            if (tid == 0) s_ob[0] = 4 * (C::OCAP + 4); // ... and so do the phantom bytes before the tile's first value
            __syncthreads(); // framing words and payload addresses are in place
            // ---- pass B: bytes (VariableByte.java:44-68: 7-bit groups, low group first, bit 7 on the LAST byte)
            if (vm) {
                const int *obp = s_ob + (ex1 >> 16); // the list this thread's first position belongs to ([0]: leading phantoms)
                int adj = *obp;
                int addr = adj + (int)(ex1 & 0xffffu);
                if (!gen) {
#pragma unroll
                    for (int k = 0; k < VPT; ++k) {
                        if ((fw >> k) & 1u) { ++obp; const int t = *obp; addr += t - adj; adj = t; }
                        const bool two = (m2 >> k) & 1u;
                        uint32_t b0 = d[k] & 127u;
                        if (!two) b0 = d[k] | 128u;
                        sb[addr] = (uint8_t)b0;
                        if (two) sb[addr + 1] = (uint8_t)((d[k] + 0x4000u) >> 7);
                        addr += 1;
                        if (two) addr += 1;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < VPT; ++k) {
                        if ((fw >> k) & 1u) { ++obp; const int t = *obp; addr += t - adj; adj = t; }
                        const int len = (int)((nib[k / 16] >> (4 * (k % 16))) & 15u);
                        uint32_t v = d[k];
                        uint8_t *dst = sb + addr;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (j < len - 1) { dst[j] = (uint8_t)(v & 127u); v >>= 7; }
                        dst[len - 1] = (uint8_t)(v | 128u);
                        addr += len;
                    }
                }
            }
            __syncthreads(); // the bytes are in place
            const uint32_t sel = a.app ? 0x0123u : 0x3210u;
            uint32_t *gal = gout - sh; // 16-byte aligned; staged word i goes to gal[i], i in [sh, sh + T)
            const int wend = sh + T;
            for (int v = tid; 4 * v < wend; v += THREADS) {
                const uint4 q = *reinterpret_cast<const uint4 *>(s_out + 4 * v);
                const uint4 o = make_uint4(__byte_perm(q.x, 0, sel), __byte_perm(q.y, 0, sel), __byte_perm(q.z, 0, sel), __byte_perm(q.w, 0, sel));
                if (4 * v >= sh && 4 * v + 4 <= wend) *reinterpret_cast<uint4 *>(gal + 4 * v) = o;
                else {
                    if (4 * v + 0 >= sh && 4 * v + 0 < wend) gal[4 * v + 0] = o.x;
                    if (4 * v + 1 >= sh && 4 * v + 1 < wend) gal[4 * v + 1] = o.y;
                    if (4 * v + 2 >= sh && 4 * v + 2 < wend) gal[4 * v + 2] = o.z;
                    if (4 * v + 3 >= sh && 4 * v + 3 < wend) gal[4 * v + 3] = o.w;
                }
            }
            return;
        }
    }
    // ---- a tile that does not fit: sizes (and streams) by the warp-per-list routines
    const CodecTraits tr = codec_traits(a.codec);
    auto list_off = [&](int64_t l) -> int64_t { return a.list_off ? a.list_off[l] : l * a.uniform_len; };
    bool page = false;
    if (tid < nl) page = ta.page > 0 && my_n >= ta.page;
    const int has_page = __syncthreads_or(page);
    uint32_t *stage = s_out + 64 * warp;
    int64_t *s_w64 = reinterpret_cast<int64_t *>(s_vals); // (the staging tile is scratch here: the lists' sizes in words)
    if (has_page) {
        if (ta.defer_cnt) return; // k_vb4_pages writes this tile
        if (tid == 0) atomicMax(a.err, 2); // a FastPFOR page: the host takes the batched kernels
    } else {
        for (int j = warp; j < nl; j += C::WARPS) {
            const int64_t beg = list_off(l0 + j), n = list_off(l0 + j + 1) - beg;
            const int64_t w = warp_short_list_encode<false>(tr, ValSeq{a.vals + beg, a.delta}, n, WordSink{nullptr, a.app}, a.app, stage, lane);
            if (lane == 0) s_w64[j] = w;
            __syncwarp();
        }
    }
    __syncthreads();
    const long long wj = (!has_page && tid < nl) ? (long long)s_w64[tid] : 0;
    long long incl = wj;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const long long t = __shfl_up_sync(kFull, incl, dd);
        if (lane >= dd) incl += t;
    }
    if (lane == 31) s_red[warp] = incl;
    __syncthreads();
    long long wo = 0, T = 0;
#pragma unroll
    for (int w = 0; w < C::WARPS; ++w) {
        const long long x = s_red[w];
        if (w < warp) wo += x;
        T += x;
    }
    if (!EMIT) {
        if (tid == 0) ta.tile_words[tile] = T;
        return;
    }
    const long long wbase = wo + incl - wj;
    if (tid < nl) a.out_off[l0 + tid] = 4 * (gbase + wbase);
    __syncthreads(); // (s_w64 is overwritten next)
    if (tid < nl) s_w64[tid] = wbase;
    __syncthreads();
    if (!has_page) {
        uint32_t *gout = reinterpret_cast<uint32_t *>(a.out) + gbase;
        for (int j = warp; j < nl; j += C::WARPS) {
            const int64_t beg = list_off(l0 + j), n = list_off(l0 + j + 1) - beg;
            warp_short_list_encode<true>(tr, ValSeq{a.vals + beg, a.delta}, n, WordSink{gout + s_w64[j], a.app}, a.app, stage, lane);
            __syncwarp();
        }
    }
}

constexpr int VB4_THREADS = 128, VB4_VPT = 32;
using VB4 = V4<VB4_THREADS, VB4_VPT>;

// ---- size pass of the flat VariableByte encoder, on its own: what the emit pass needs from it is one number per tile, and
// that number needs no prefixes, no staging and no per-thread state -- a list's bytes are its length, plus one for every gap of
// 128 or more (a popcount over the tile's "two bytes" bit vector), plus the rare third to fifth bytes (added to the list's
// count one by one).  A thread forms 8 consecutive gaps from two 16-byte loads (a warp reads 1 KB at a time) and stores one byte
// of the bit vector; 4 rounds cover the tile.  HBM-bound: 65 us for the 400 MB of C4 (the tile kernel used as size pass: 82).
template <bool DELTA>
__global__ void __launch_bounds__(VB4_THREADS) k_vb4_size(const TileEncArgs ta)
{
    using C = V4<VB4_THREADS, 32>;
    const FastEncArgs &a = ta.f;
    __shared__ __align__(16) uint32_t s_bits[C::SPAN / 32];  // bit p: the value at window position p takes two bytes or more
    __shared__ uint32_t s_flag[C::SPAN / 32 + 1];             // bit p: a non-empty list starts at window position p
    __shared__ int s_rel[VB4_THREADS + 1];                    // list starts as window positions ([nl] = the tile's end)
    __shared__ int s_extra[VB4_THREADS];                      // per list: bytes beyond the second of its values
    __shared__ uint32_t s_w[C::WARPS];
    __shared__ long long s_red[C::WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long tile = blockIdx.x;
    const int64_t l0 = tile * a.lists_per_batch;
    const int nl = (int)min((int64_t)a.lists_per_batch, a.nlists - l0);
    int64_t base, end, my_off = 0, my_end = 0;
    if (a.list_off) {
        const int64_t *lo = a.list_off + l0;
        base = lo[0]; end = lo[nl];
        if (tid <= nl) { my_off = lo[tid]; my_end = lo[tid < nl ? tid + 1 : tid]; }
    } else {
        base = l0 * a.uniform_len; end = base + nl * a.uniform_len;
        my_off = base + min(tid, nl) * a.uniform_len; my_end = base + min(tid + 1, nl) * a.uniform_len;
    }
    const int64_t nv = end - base;
    const int a0 = (int)((reinterpret_cast<uintptr_t>(a.vals + base) >> 2) & 3); // the window starts on a 16-byte boundary
    bool bad = nv < 0 || a0 + nv > C::SPAN;
    const int span = bad ? 0 : a0 + (int)nv;
    const int64_t my_n = my_end - my_off;
    if (tid <= nl) bad = bad || my_off < base || my_off > end || my_n < 0 || (ta.page > 0 && my_n >= ta.page);
    s_flag[tid] = 0;
    if (tid == 0) s_flag[C::SPAN / 32] = 0;
    s_extra[tid] = 0;
    // the loads do not depend on the list books: request them first
    const uint32_t *wp = a.vals + base - a0;
    uint4 u[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = 8 * (i * VB4_THREADS + tid);
        u[i][0] = p < span ? __ldg(reinterpret_cast<const uint4 *>(wp + p)) : make_uint4(0, 0, 0, 0);
        u[i][1] = p + 4 < span ? __ldg(reinterpret_cast<const uint4 *>(wp + p + 4)) : make_uint4(0, 0, 0, 0);
    }
    __syncthreads(); // (the start bits are zero)
    if (tid <= nl && !bad) {
        const int rel = a0 + (int)(my_off - base);
        s_rel[tid] = rel;
        if (tid < nl && my_n > 0) atomicOr(&s_flag[rel >> 5], 1u << (rel & 31));
    }
    const int slow = __syncthreads_or(bad);
    if (!slow) {
        uint8_t *sbits8 = reinterpret_cast<uint8_t *>(s_bits);
        const uint8_t *sflag8 = reinterpret_cast<const uint8_t *>(s_flag);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int bidx = i * VB4_THREADS + tid, p = 8 * bidx; // this thread's byte of the bit vector, its first position
            const uint32_t x[8] = {u[i][0].x, u[i][0].y, u[i][0].z, u[i][0].w, u[i][1].x, u[i][1].y, u[i][1].z, u[i][1].w};
            uint32_t prev = __shfl_up_sync(kFull, x[7], 1);
            if (lane == 0) prev = (p > 0 && p < span) ? __ldg(wp + p - 1) : 0u;
            uint32_t bits = 0;
            if (p < span) {
                const uint32_t st = sflag8[bidx];
                uint32_t d[8], wor = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    d[k] = (DELTA && !((st >> k) & 1u)) ? x[k] - prev : x[k];
                    prev = x[k];
                }
                if (p < a0 || p + 8 > span) { // positions outside the tile count as nothing
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (p + k < a0 || p + k >= span) d[k] = 0u;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    wor |= d[k];
                    bits |= (d[k] >= 128u ? 1u : 0u) << k;
                }
                if (wor >= 16384u) { // rare: third to fifth bytes, added to their list's count (the list: by bisection)
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (d[k] >= 16384u) {
                            int lo = 0, hi = nl; // s_rel[lo] <= p + k < s_rel[hi]
                            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_rel[mid] <= p + k) lo = mid; else hi = mid; }
                            atomicAdd(&s_extra[lo], vb_len(d[k]) - 2);
                        }
                    }
                }
            }
            sbits8[bidx] = (uint8_t)bits;
        }
        __syncthreads();
        int words = 0;
        if (tid < nl) {
            const int n = (int)my_n;
            int bytes = n + s_extra[tid];
            if (n > 0) {
                const int plo = s_rel[tid], phi = plo + n, wlo = plo >> 5, whi = (phi - 1) >> 5;
                for (int w = wlo; w <= whi; ++w) {
                    uint32_t m = s_bits[w];
                    if (w == wlo) m &= 0xffffffffu << (plo & 31);
                    if (w == whi && (phi & 31)) m &= (1u << (phi & 31)) - 1u;
                    bytes += __popc(m);
                }
            }
            const int hdr = ta.hdr_mode == 1 ? 1 : (ta.hdr_mode == 2 && n > 0 ? 1 : 0);
            words = hdr + ((bytes + 3) >> 2) + (a.app ? 1 : 0);
        }
        words = __reduce_add_sync(kFull, words);
        if (lane == 0) s_w[warp] = (uint32_t)words;
        __syncthreads();
        if (tid == 0) {
            long long T = 0;
#pragma unroll
            for (int w = 0; w < C::WARPS; ++w) T += s_w[w];
            ta.tile_words[tile] = T;
        }
        return;
    }
    // ---- a tile that does not fit the window: sizes by the warp-per-list routines
    const CodecTraits tr = codec_traits(a.codec);
    auto list_off = [&](int64_t l) -> int64_t { return a.list_off ? a.list_off[l] : l * a.uniform_len; };
    bool page = false;
    if (tid < nl) page = ta.page > 0 && my_n >= ta.page;
    const int has_page = __syncthreads_or(page);
    __shared__ uint32_t s_stage[C::WARPS][64];
    long long mine = 0;
    if (has_page) {
        if (ta.defer_cnt) { // k_vb4_pages sizes this tile (before the scan)
            if (tid == 0) ta.defer_tiles[atomicAdd(ta.defer_cnt, 1)] = (int)tile;
            return;
        }
        if (tid == 0) atomicMax(a.err, 2); // a FastPFOR page: the host takes the batched kernels
    } else {
        for (int j = warp; j < nl; j += C::WARPS) {
            const int64_t beg = list_off(l0 + j), n = list_off(l0 + j + 1) - beg;
            mine += warp_short_list_encode<false>(tr, ValSeq{a.vals + beg, a.delta}, n, WordSink{nullptr, a.app}, a.app, s_stage[warp], lane);
            __syncwarp();
        }
    }
    if (lane == 0) s_red[warp] = mine;
    __syncthreads();
    if (tid == 0) {
        long long T = 0;
#pragma unroll
        for (int w = 0; w < C::WARPS; ++w) T += s_red[w];
        ta.tile_words[tile] = T;
    }
}

// The tiles the two kernels above left aside (a list of at least one FastPFOR block among lists that are otherwise all
// VariableByte -- FastPFOR128 on ~100-value lists meets one every few hundred lists): sized (EMIT = false, before the scan)
// and written (EMIT = true, after it) by the warp-per-list routines, a CTA per tile of the worklist.  Kept out of the
// kernels above on purpose: inlined there, the FastPFOR page encoder costs them 8-24 registers and 4 KB of shared memory.
template <bool EMIT, bool OPT> // OPT: OptPFD (its pricing scratch)
__global__ void __launch_bounds__(VB4_THREADS) k_vb4_pages(const TileEncArgs ta)
{
    constexpr int WARPS = VB4_THREADS / 32;
    const FastEncArgs &a = ta.f;
    __shared__ uint32_t s_stage[WARPS][64];
    __shared__ int s_fq[WARPS][68];
    __shared__ int32_t s_blk[WARPS][OPT ? 256 : 1];
    __shared__ __align__(4) uint8_t s_s16[OPT ? kS16TabBytes : 4];
    if (OPT) { s16_tables(s_s16, threadIdx.x, VB4_THREADS); __syncthreads(); }
    __shared__ long long s_w64[VB4_THREADS];
    __shared__ long long s_red[WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const CodecTraits tr = codec_traits(a.codec);
    auto list_off = [&](int64_t l) -> int64_t { return a.list_off ? a.list_off[l] : l * a.uniform_len; };
    const int ndefer = *ta.defer_cnt;
    if (EMIT && *a.total > a.cap) return;
    for (int d = blockIdx.x; d < ndefer; d += gridDim.x) {
        const long long tile = ta.defer_tiles[d];
        const int64_t l0 = tile * a.lists_per_batch;
        const int nl = (int)min((int64_t)a.lists_per_batch, a.nlists - l0);
        bool multi = false;
        for (int j = warp; j < nl; j += WARPS) {
            const int64_t beg = list_off(l0 + j), n = list_off(l0 + j + 1) - beg;
            const int64_t w = warp_any_list_encode<false, OPT>(tr, ValSeq{a.vals + beg, a.delta}, n, WordSink{nullptr, a.app}, a.app, s_stage[warp],
                                                               s_fq[warp], s_blk[warp], s_s16, lane, &multi);
            if (lane == 0) s_w64[j] = w;
            __syncwarp();
        }
        if (__syncthreads_or(multi)) { // a multi-page list: the host takes the warp-per-list kernels (stale-buffer emulation)
            if (tid == 0) { atomicMax(a.err, 2); if (!EMIT) ta.tile_words[tile] = 0; }
            continue;
        }
        const long long wj = tid < nl ? s_w64[tid] : 0;
        long long incl = wj;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const long long t = __shfl_up_sync(kFull, incl, dd);
            if (lane >= dd) incl += t;
        }
        if (lane == 31) s_red[warp] = incl;
        __syncthreads();
        long long wo = 0, T = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const long long x = s_red[w];
            if (w < warp) wo += x;
            T += x;
        }
        if (!EMIT) {
            if (tid == 0) ta.tile_words[tile] = T;
        } else {
            const long long gbase = ta.tile_words[tile], wbase = wo + incl - wj;
            if (tid < nl) a.out_off[l0 + tid] = 4 * (gbase + wbase);
            __syncthreads();
            if (tid < nl) s_w64[tid] = wbase;
            __syncthreads();
            uint32_t *gout = reinterpret_cast<uint32_t *>(a.out) + gbase;
            for (int j = warp; j < nl; j += WARPS) {
                const int64_t beg = list_off(l0 + j), n = list_off(l0 + j + 1) - beg;
                warp_any_list_encode<true, OPT>(tr, ValSeq{a.vals + beg, a.delta}, n, WordSink{gout + s_w64[j], a.app}, a.app, s_stage[warp],
                                                s_fq[warp], s_blk[warp], s_s16, lane, &multi);
                __syncwarp();
            }
        }
        __syncthreads();
    }
}

// exclusive scan of the tiles' word counts in place (one CTA: a few ten thousand tiles).  Every thread owns a contiguous run
// of counts and loads it with 16-byte loads that are all in flight at once (the kernel is two memory round trips and a block
// scan: ~5 us for 26 000 tiles; a warp-per-segment version that walked its segment 32 counts at a time took 18).  The total in
// bytes goes to *total and to out_off[nlists], flag 1 when the caller's buffer is too small for it.
constexpr int SCN_THREADS = 1024, SCN_VEC = 12; // longlong2 loads per thread and pass (24 counts)
__global__ void __launch_bounds__(SCN_THREADS, 1) k_vb4_scan(int64_t *tile_words, int64_t nb, int64_t *total, int64_t *out_off_end, int64_t cap, int *err)
{
    __shared__ long long s_w[32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    const int64_t chunk = (int64_t)SCN_THREADS * 2 * SCN_VEC; // counts per pass of the CTA
    for (int64_t c0 = 0; c0 < nb; c0 += chunk) {
        const int64_t i0 = c0 + (int64_t)tid * 2 * SCN_VEC; // (even: 16-byte aligned, the array starts 32 bytes into a 256-byte aligned block)
        longlong2 x[SCN_VEC];
#pragma unroll
        for (int k = 0; k < SCN_VEC; ++k) {
            const int64_t i = i0 + 2 * k;
            if (i + 1 < nb) x[k] = *reinterpret_cast<const longlong2 *>(tile_words + i);
            else x[k] = make_longlong2(i < nb ? tile_words[i] : 0, 0);
        }
        long long sum = 0;
#pragma unroll
        for (int k = 0; k < SCN_VEC; ++k) sum += x[k].x + x[k].y;
        long long incl = sum;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const long long t = __shfl_up_sync(kFull, incl, dd);
            if (lane >= dd) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        long long run = s_carry, tot = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            const long long y = s_w[w];
            if (w < warp) run += y;
            tot += y;
        }
        run += incl - sum;
#pragma unroll
        for (int k = 0; k < SCN_VEC; ++k) {
            const int64_t i = i0 + 2 * k;
            const long long e0 = run, e1 = run + x[k].x;
            run = e1 + x[k].y;
            if (i + 1 < nb) *reinterpret_cast<longlong2 *>(tile_words + i) = make_longlong2(e0, e1);
            else if (i < nb) tile_words[i] = e0;
        }
        __syncthreads();
        if (tid == 0) s_carry += tot;
        __syncthreads();
    }
    if (tid == 0) {
        const long long tot = s_carry;
        *total = 4 * tot;
        *out_off_end = 4 * tot;
        if (4 * tot > cap) atomicMax(err, 1);
    }
}

// ---- flat VariableByte decoder.  One tile per CTA (a fixed number of whole lists: at most D4_WCAP compressed words, D4_VSPAN
// values); no cross-CTA state -- the output offsets are an input.
//   words   arrive by coalesced 16-byte cp.async; a thread owns 4g consecutive words (g = 1, 3 or 5, odd: its LDS.128 are
//           conflict-free), sets IntCompressor count words to zero (their bytes are not VariableByte), swaps the loader's byte
//           order, counts terminator bits and writes the clean words back;
//   values  one block scan gives every thread the index of its first value; the byte walk (VariableByte.java:115-138)
//           restarts at every list's first payload word (the Java decoder starts each list with v = 0, shift = 0) and scatters
//           into a chunk-swizzled tile aligned like the output;
//   Delta   a thread owns 32 consecutive values: running sums that restart at list starts, one segmented scan across threads
//           (Delta.inverseDelta per list), rows written back; the tile leaves with coalesced 16-byte stores.
// Every list must hold exactly its n values (terminator counts by list) or the call fails with MCP_E_FORMAT.  A tile that does
// not fit is done by the warp-per-list routines inside the same kernel; a FastPFOR list of a whole block or more raises flag 2.
constexpr int D4_THREADS = 128;
constexpr int D4_WARPS = D4_THREADS / 32;
constexpr int D4_VPT = 32;
constexpr int D4_VSPAN = D4_THREADS * D4_VPT;   // 4 096 window positions of the output
constexpr int D4_WCAP = 20 * D4_THREADS;        // 2 560 compressed words
constexpr int D4_MAXL = D4_THREADS - 1;

// word of window position i of the output in the staging tile (rows of 32 values, 16-byte chunks permuted by the row index:
// conflict-free for a thread reading its row and for the coalesced copy-out)
__device__ __forceinline__ int d4_slot(int i) { return i ^ ((i >> 3) & 0x1c); }

template <bool DELTA>
__global__ void __launch_bounds__(D4_THREADS) k_vb4_decode(const FastDecArgs a, int hdr_mode, int page, int *defer_cnt, int *defer_tiles)
{
    __shared__ __align__(16) uint32_t s_words[D4_WCAP];
    __shared__ __align__(1024) uint32_t s_val[D4_VSPAN]; // (1 KB aligned: the swizzle can be applied to the shared address itself)
    __shared__ uint32_t s_wstart[D4_WCAP / 32]; // bit w: a list's payload starts at window word w (the decoder restarts there)
    __shared__ uint32_t s_whdr[D4_WCAP / 32];   // bit w: window word w is an IntCompressor count word (not payload)
    __shared__ uint32_t s_vstart[D4_VSPAN / 32 + 1]; // bit p: a non-empty list starts at window position p of the output
    __shared__ int s_wrel[D4_THREADS + 1];
    __shared__ int s_cp[D4_THREADS + 1];        // values ending before each thread's first word
    __shared__ uint32_t s_w[D4_WARPS], s_sv[D4_WARPS], s_sf[D4_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sel = a.app ? 0x0123u : 0x3210u; // big-endian framing: words are swapped as they are read
    const int64_t l0 = (int64_t)blockIdx.x * a.lists_per_batch;
    const int nl = (int)min((int64_t)a.lists_per_batch, a.nlists - l0);
    const int64_t *gin = a.in_off + l0, *goo = a.out_off + l0;
    const int64_t ib = gin[0], ie = gin[nl], ob = goo[0], oe = goo[nl];
    int64_t my_i0 = 0, my_i1 = 0, my_o0 = 0, my_o1 = 0;
    if (tid <= nl) { my_i0 = gin[tid]; my_o0 = goo[tid]; my_i1 = gin[min(tid + 1, nl)]; my_o1 = goo[min(tid + 1, nl)]; }
    const int64_t nwords = (ie - ib) >> 2, nvals = oe - ob;
    const int aw = (int)((reinterpret_cast<uintptr_t>(a.in + ib) >> 2) & 3);   // 16-byte aligned window over the words
    const int ao = (int)((reinterpret_cast<uintptr_t>(a.out + ob) >> 2) & 3);  // staging tile aligned like the output
    int st = 0; // 1 = malformed, 2 = FastPFOR page, 4 = does not fit the tile
    if (((ib | ie) & 3) != 0 || nwords < 0 || nvals < 0) st = 1;
    else if (aw + nwords > D4_WCAP || ao + nvals > D4_VSPAN) st = 4;
    if (tid < nl) {
        const int64_t n = my_o1 - my_o0;
        if (((my_i0 | my_i1) & 3) != 0 || my_i1 < my_i0 || n < 0 || my_i0 < ib || my_i1 > ie || my_o0 < ob || my_o1 > oe) st |= 1;
        else if (page > 0 && n >= page) st |= 2;
    }
    const int wspan = st ? 0 : aw + (int)nwords, vspan = st ? 0 : ao + (int)nvals;
    if (tid < D4_WCAP / 32) { s_wstart[tid] = 0; s_whdr[tid] = 0; }
    s_vstart[tid] = 0;
    if (tid == 0) s_vstart[D4_VSPAN / 32] = 0;
    {
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(a.in + ib) - aw + 4 * tid;
#pragma unroll
        for (int i = 0; i < D4_WCAP / 4 / D4_THREADS; ++i)
            if (4 * (i * D4_THREADS + tid) < wspan) cp_async16(s_words + 4 * (i * D4_THREADS + tid), wp + 4 * i * D4_THREADS);
    }
    cp_async_commit();
    cp_async_wait<0>();
    int tile_st = 0; // (__syncthreads_or returns a truth value, not the OR of the arguments)
    if (__syncthreads_or(st)) tile_st = (__syncthreads_or(st & 1) ? 1 : 0) | (__syncthreads_or(st & 2) ? 2 : 0) | (__syncthreads_or(st & 4) ? 4 : 0);
    if (tile_st & 1) { if (tid == 0) atomicMax(a.err, 3); return; }
    if (tile_st & 2) {
        if (tid == 0) {
            if (defer_cnt) defer_tiles[atomicAdd(defer_cnt, 1)] = (int)blockIdx.x; // k_vb4_decode_pages takes this tile
            else atomicMax(a.err, 2);
        }
        return;
    }
    if (tile_st & 4) {
        const CodecTraits tr = codec_traits(a.codec);
        for (int j = warp; j < nl; j += D4_WARPS) {
            const WordSrc src{reinterpret_cast<const uint32_t *>(a.in + gin[j]), a.app};
            const bool ok = warp_short_list_decode(tr, src, (gin[j + 1] - gin[j]) >> 2, goo[j + 1] - goo[j], a.out + goo[j], a.delta, lane);
            if (!ok && lane == 0) atomicMax(a.err, 4);
            __syncwarp();
        }
        return;
    }
    // ---- per-list books: header checks, restart / count-word / first-value bits (the words are staged, the bitmaps zero)
    bool ok = true;
    int my_n = 0;
    if (tid <= nl) s_wrel[tid] = aw + (int)((my_i0 - ib) >> 2);
    if (tid < nl) {
        const int w0 = aw + (int)((my_i0 - ib) >> 2), w1 = aw + (int)((my_i1 - ib) >> 2);
        my_n = (int)(my_o1 - my_o0);
        const int hdr = hdr_mode == 1 ? 1 : (hdr_mode == 2 && my_n > 0 ? 1 : 0);
        if (w1 - w0 < hdr) ok = false;
        else {
            if (hdr) {
                const uint32_t h = __byte_perm(s_words[w0], 0, sel);
                if (h != (hdr_mode == 1 ? (uint32_t)my_n : 0u)) ok = false; // IntCompressor's n / Composition's marker (FastPFOR took nothing)
                if (hdr_mode == 1) atomicOr(&s_whdr[w0 >> 5], 1u << (w0 & 31));
            }
            if (w1 > w0 + hdr) atomicOr(&s_wstart[(w0 + hdr) >> 5], 1u << ((w0 + hdr) & 31));
            else if (my_n > 0) ok = false;
            if (my_n > 0) { const int p = ao + (int)(my_o0 - ob); atomicOr(&s_vstart[p >> 5], 1u << (p & 31)); }
        }
    }
    __syncthreads(); // the bitmaps are in place
    // ---- this thread's 4g words: clean, count the values that end in them
    const int g = ((wspan + 4 * D4_THREADS - 1) / (4 * D4_THREADS)) | 1;
    const int q0 = 4 * g * tid;
    int cnt = 0;
    for (int i = 0; i < g; ++i) {
        const int q = q0 + 4 * i;
        if (q < wspan) {
            uint4 u = *reinterpret_cast<const uint4 *>(s_words + q);
            const uint32_t hm = (s_whdr[q >> 5] >> (q & 31)) & 15u;
            uint32_t w[4] = {u.x, u.y, u.z, u.w};
            if (q >= aw && q + 4 <= wspan && hm == 0) { // (the usual case: four payload or zero words)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    w[k] = __byte_perm(w[k], 0, sel);
                    cnt += __popc(w[k] & 0x80808080u);
                }
                if (a.app) *reinterpret_cast<uint4 *>(s_words + q) = make_uint4(w[0], w[1], w[2], w[3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool in = q + k >= aw && q + k < wspan && !((hm >> k) & 1u);
                    w[k] = in ? __byte_perm(w[k], 0, sel) : 0u;
                    cnt += __popc(w[k] & 0x80808080u);
                }
                *reinterpret_cast<uint4 *>(s_words + q) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
    uint32_t incl = (uint32_t)cnt;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, dd);
        if (lane >= dd) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads(); // (also: the clean words are in place)
    uint32_t wo = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < D4_WARPS; ++w) {
        const uint32_t x = s_w[w];
        if (w < warp) wo += x;
        tot += x;
    }
    const int V0 = (int)(wo + incl) - cnt;
    s_cp[tid] = V0;
    if (tid == 0) s_cp[D4_THREADS] = (int)tot;
    __syncthreads();
    // ---- every list must hold exactly its n values: values ending before its first word = values before it
    if (tid <= nl) {
        const int wpos = tid < nl ? s_wrel[tid] : wspan;
        const int t = min(wpos / (4 * g), D4_THREADS);
        int c = s_cp[t];
        if (t < D4_THREADS)
            for (int p = 4 * g * t; p < wpos; ++p) c += __popc(s_words[p] & 0x80808080u);
        if (c != (int)(my_o0 - ob)) ok = false;
    }
    if ((int)tot != (int)nvals) ok = false;
    if (__syncthreads_or(!ok)) { if (tid == 0) atomicMax(a.err, 5); return; }
    // ---- the byte walk (VariableByte.java:115-138): v += (c & 127) << shift; a set bit 7 ends the value.  The shift is kept
    // as a multiplier (1 << shift: one IMAD per byte instead of a shift and an add); past the fifth byte of a value -- which
    // only the zero bytes after a list's last value reach -- it is 0, and those bytes add nothing either way.
    {
        uint32_t acc = 0, mul = 1;
        if (q0 > aw && q0 < wspan && !((s_wstart[q0 >> 5] >> (q0 & 31)) & 1u)) { // bytes of a value that began in the previous word
            const uint32_t pv = s_words[q0 - 1];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t by = (pv >> (8 * e)) & 255u;
                if (by & 128u) { acc = 0; mul = 1; }
                else { acc += (by & 127u) * mul; mul <<= 7; }
            }
            // (if the previous word is itself a list's first payload word, only its own bytes count: exactly what was walked)
        }
        // shared byte address of value slot idx before the swizzle; d4_slot on the address: a ^ ((a >> 3) & 0x70)
        uint32_t va = (uint32_t)__cvta_generic_to_shared(s_val) + 4u * (uint32_t)(ao + V0);
        for (int i = 0; i < g; ++i) {
            const int q = q0 + 4 * i;
            if (q < wspan) {
                const uint4 u = *reinterpret_cast<const uint4 *>(s_words + q);
                const uint32_t fl = (s_wstart[q >> 5] >> (q & 31)) & 15u;
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if ((fl >> k) & 1u) { acc = 0; mul = 1; }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t by = (w[k] >> (8 * e)) & 255u;
                        acc += (by & 127u) * mul;
                        if (by & 128u) {
                            asm volatile("st.shared.u32 [%0], %1;" ::"r"(va ^ ((va >> 3) & 0x70u)), "r"(acc) : "memory");
                            va += 4; acc = 0; mul = 1;
                        } else
                            mul <<= 7;
                    }
                }
            }
        }
    }
    __syncthreads(); // the values are staged
    if (DELTA) {
        // ---- Delta.inverseDelta per list: running sums that restart at list starts
        uint32_t *row = s_val + D4_VPT * tid;
        const uint32_t fw = s_vstart[tid];
        uint32_t y[D4_VPT];
#pragma unroll
        for (int c = 0; c < D4_VPT / 4; ++c) {
            const uint4 u = *reinterpret_cast<const uint4 *>(row + 4 * (c ^ (tid & 7)));
            y[4 * c] = u.x; y[4 * c + 1] = u.y; y[4 * c + 2] = u.z; y[4 * c + 3] = u.w;
        }
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < D4_VPT; ++k) {
            if ((fw >> k) & 1u) run = 0;
            run += y[k];
            y[k] = run;
        }
        // segmented scan over the threads: (sum since the last start, a start was seen)
        uint32_t v = run, f = fw != 0;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const uint32_t pv = __shfl_up_sync(kFull, v, dd), pf = __shfl_up_sync(kFull, f, dd);
            if (lane >= dd) { if (!f) v += pv; f |= pf; }
        }
        if (lane == 31) { s_sv[warp] = v; s_sf[warp] = f; }
        uint32_t ev = __shfl_up_sync(kFull, v, 1), ef = __shfl_up_sync(kFull, f, 1);
        if (lane == 0) { ev = 0; ef = 0; }
        __syncthreads();
        uint32_t cw = 0;
#pragma unroll
        for (int w = 0; w < D4_WARPS - 1; ++w)
            if (w < warp) cw = s_sf[w] ? s_sv[w] : cw + s_sv[w];
        const uint32_t carry = ef ? ev : cw + ev;
        const int kf = fw ? __ffs(fw) - 1 : D4_VPT; // the carry reaches the values before this thread's first list start
#pragma unroll
        for (int k = 0; k < D4_VPT; ++k)
            if (k < kf) y[k] += carry;
#pragma unroll
        for (int c = 0; c < D4_VPT / 4; ++c)
            *reinterpret_cast<uint4 *>(row + 4 * (c ^ (tid & 7))) = make_uint4(y[4 * c], y[4 * c + 1], y[4 * c + 2], y[4 * c + 3]);
        __syncthreads();
    }
    // ---- the tile leaves: coalesced 16-byte stores (the staging tile is aligned like the output)
    uint32_t *gal = a.out + ob - ao;
#pragma unroll 2
    for (int q = tid; 4 * q < vspan; q += D4_THREADS) {
        const uint4 u = *reinterpret_cast<const uint4 *>(s_val + d4_slot(4 * q));
        if (q > 0 && 4 * q + 4 <= vspan) *reinterpret_cast<uint4 *>(gal + 4 * q) = u;
        else {
            if (4 * q + 0 >= ao && 4 * q + 0 < vspan) gal[4 * q + 0] = u.x;
            if (4 * q + 1 >= ao && 4 * q + 1 < vspan) gal[4 * q + 1] = u.y;
            if (4 * q + 2 >= ao && 4 * q + 2 < vspan) gal[4 * q + 2] = u.z;
            if (4 * q + 3 >= ao && 4 * q + 3 < vspan) gal[4 * q + 3] = u.w;
        }
    }
}


// ---------------------------------------------------------------------- scan
// exclusive scan of int64: per-block partial sums, serial scan of the partials by one
// block, then a second sweep.  Sizes here are "number of lists": bookkeeping, not hot.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;

// The tiles k_vb4_decode left aside (a list with a FastPFOR page): warp-per-list routines, a CTA per tile of the worklist.
constexpr int D4P_SCR = 1280; // words a warp may un-swap a big-endian-framed page into (a list beyond it: flag 2)
template <bool OPT>
__global__ void __launch_bounds__(D4_THREADS) k_vb4_decode_pages(const FastDecArgs a, const int *defer_cnt, const int *defer_tiles)
{
    __shared__ uint32_t s_scr[D4_WARPS][D4P_SCR];
    __shared__ int s_fq[D4_WARPS][68];
    __shared__ uint32_t s_ex[D4_WARPS][OPT ? 2 * 128 + 28 : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const CodecTraits tr = codec_traits(a.codec);
    const int ndefer = *defer_cnt;
    for (int d = blockIdx.x; d < ndefer; d += gridDim.x) {
        const int64_t l0 = (int64_t)defer_tiles[d] * a.lists_per_batch;
        const int nl = (int)min((int64_t)a.lists_per_batch, a.nlists - l0);
        const int64_t *gin = a.in_off + l0, *goo = a.out_off + l0;
        for (int j = warp; j < nl; j += D4_WARPS) {
            const WordSrc src{reinterpret_cast<const uint32_t *>(a.in + gin[j]), a.app};
            bool multi = false;
            const bool ok = warp_any_list_decode<OPT>(tr, src, (gin[j + 1] - gin[j]) >> 2, goo[j + 1] - goo[j], a.out + goo[j], a.delta, s_scr[warp],
                                                      D4P_SCR, s_fq[warp], s_ex[warp], lane, &multi);
            if (lane == 0 && multi) atomicMax(a.err, 2); // the host takes the batched / warp-per-list kernels
            else if (!ok && lane == 0) atomicMax(a.err, 4);
            __syncwarp();
        }
    }
}

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

// ---- fast-path launchers -------------------------------------------------------------------------------------
// lists per batch: fill ~92 % of the value staging area on average (a batch that overflows is still handled, by the
// per-warp routines); 0 = the fast path is not worth it (long lists: the per-warp kernels are the better fit)
static int fast_lists_per_batch(int codec, int64_t total_vals, int64_t nlists, int64_t uniform_len = 0)
{
    if (nlists <= 0) return 0;
    const CodecTraits tr = codec_traits(codec);
    const double avg = (double)total_vals / (double)nlists;
    if (avg > (tr.fastpfor ? 4000.0 : 1024.0)) return 0; // (a FastPFOR batch may be a single list of up to FE_VCAP values)
    if (uniform_len > 0 && tr.fastpfor) {
        // lists of one known length (the risk step's breach lists): the staging areas can be filled exactly -- 1 001-int lists
        // go four to a batch (96 of 128 threads hold a 32-value group) instead of three
        const int64_t tail = uniform_len % tr.fp_block;
        int64_t g = FE_VCAP / uniform_len;
        if (tail > 0 && FE_TCAP / tail < g) g = FE_TCAP / tail;
        const int64_t nb = uniform_len / tr.fp_block;
        if (nb > 0 && g * nb > FP_MAXBLK) g = FP_MAXBLK / nb;
        return (int)(g < 1 ? 1 : (g > FP_MAXL ? FP_MAXL : g));
    }
    int64_t g = (int64_t)(0.92 * FE_VCAP / (avg > 1.0 ? avg : 1.0));
    int64_t gmax = FE_MAXL;
    if (tr.fastpfor) { // VariableByte tails of up to block-1 values per list share the tail table
        const double tail = avg < tr.fp_block - 1 ? avg : tr.fp_block - 1;
        const int64_t gt = (int64_t)(0.85 * FE_TCAP / (tail > 1.0 ? tail : 1.0));
        if (gt < g) g = gt;
        gmax = FP_MAXL;
    }
    if (g < 1) g = 1;
    if (g > gmax) g = gmax;
    return (int)g;
}

// lists per tile of the flat VariableByte kernels: `target` values per tile on average (the tile's window holds
// VF_SPAN values / VF_WSPAN words; a tile that overflows is done by the warp-per-list routines inside the kernel)
static int vbf_lists_per_tile(int64_t total_vals, int64_t nlists, double target, int maxl)
{
    const double avg = nlists > 0 ? (double)total_vals / (double)nlists : 1.0;
    int64_t g = (int64_t)(target / (avg > 1.0 ? avg : 1.0));
    return (int)(g < 1 ? 1 : (g > maxl ? maxl : g));
}

// the batched (thread-per-block) kernels: BinaryPacking family and FastPFOR / FastPFOR128
static bool batched_applicable(int codec, bool delta, int64_t total_vals, int64_t nlists)
{
    const CodecTraits tr = codec_traits(codec);
    const bool ok = fast_codec_ok(codec, delta) || (tr.fastpfor && !tr.sk); // (the batched FastPFOR kernels write Composition's framing)
    return ok && fast_lists_per_batch(codec, total_vals, nlists) > 0;
}

// the short-list (thread-per-list) kernels: every codec, lists averaging <= SL_MAXAVG values
static bool shortlist_applicable(mcp_ctx *ctx, int codec, bool delta, int64_t total_vals, int64_t nlists, bool encode)
{
    if (ctx && !ctx->shortListCodec) return false;
    // measured on C4 (profiles/r1/codec_short_lists.txt): the encoder beats the batched kernels for every codec (0.39 vs
    // 0.41 ms for the BinaryPacking family, 0.51 vs 1.59 ms when the whole list is a VariableByte tail -- FastPFOR /
    // FastPFOR128 below one block, VariableByte alone); the decoder only for the latter (0.55 vs 0.88 ms; 0.43 vs 0.37 ms for
    // the BinaryPacking family, which it takes only when asked to: option short_list_codec = 2)
    const CodecTraits tr = codec_traits(codec);
    if (!encode && tr.blocks && !(ctx && ctx->shortListCodec >= 2)) return false;
    return nlists > 0 && sl_codec_ok(codec, delta) && (double)total_vals <= SL_MAXAVG * (double)nlists;
}

bool codec_fast_applicable(int codec_id, int64_t total_vals, int64_t nlists, bool encode)
{
    const int codec = codec_id & MCP_CODEC_MASK;
    const bool delta = (codec_id & MCP_FLAG_DELTA) != 0;
    (void)encode; // both directions
    const CodecTraits tr = codec_traits(codec);
    if ((tr.optpfd || (tr.xorbp && !delta)) && nlists > 0 && (double)total_vals < 0.95 * 128.0 * (double)nlists) return true; // flat_only_codec
    return batched_applicable(codec, delta, total_vals, nlists) || shortlist_applicable(nullptr, codec, delta, total_vals, nlists, encode);
}

// upper bound of the encoded size in bytes for the fast-path codecs: 33 words per 32-int block worst case, a 31-value
// VariableByte tail of 5 bytes each, count / header / trailing-zero words; FastPFOR: at most 32 bits per value plus the
// byte metadata and page words, and a tail of up to 255 values
int64_t codec_fast_max_bytes(int codec_id, int64_t total_vals, int64_t nlists)
{
    const CodecTraits tr = codec_traits(codec_id & MCP_CODEC_MASK);
    if (tr.fastpfor) return 4 * (total_vals + total_vals / 8 + 384 * nlists + 64);
    if (tr.optpfd) return 4 * (3 * total_vals + total_vals / 4 + 8 * nlists + 64); // a block: header + up to 254 S16 words + 4 x 20 packed (codec_max_words)
    if (!tr.blocks) return 4 * (total_vals + total_vals / 4 + 4 * nlists + 64); // VariableByte alone: 5 bytes per value at worst
    return 4 * (total_vals + total_vals / 32 + 44 * nlists + 64);
}

// XorBinaryPacking / OptPFD (128-int blocks) on short lists: no thread-per-list or batched kernel takes them, but lists below one
// block are Composition's 0 marker + VariableByte (of gaps, for XorBinaryPacking's integrated tail) like FastPFOR128's: the flat
// kernels serve the call, k_vb4_pages the tiles that hold a block.  (Delta before the integrated codec is not a combination anyone uses.)
static bool flat_only_codec(const mcp_ctx *ctx, int codec, bool delta, int64_t total_vals, int64_t nlists)
{
    const CodecTraits tr = codec_traits(codec);
    if (!(tr.xorbp || tr.optpfd) || (tr.xorbp && delta)) return false;
    return ctx->shortListCodec && ctx->flatCodec && nlists > 0 && (double)total_vals < 0.95 * 128.0 * (double)nlists;
}

// flat kernels on a FastPFOR codec: is a list of a full block something to plan for (average length within a third of
// the block) or an outlier (then the flat call fails over to the batched kernels as a whole)?
static bool flat_pages_likely(const CodecTraits &tr, int64_t total_vals, int64_t nlists)
{
    const int64_t blk = tr.fastpfor ? tr.fp_block : ((tr.xorbp || tr.optpfd) ? 128 : 0);
    return blk > 0 && nlists > 0 && 1.5 * (double)total_vals >= (double)blk * (double)nlists;
}

// ... or the rule (average length at the block size or beyond: the thread-per-list / flat kernels are not tried at all)
static bool flat_pages_common(int codec, int64_t total_vals, int64_t nlists)
{
    const CodecTraits tr = codec_traits(codec);
    return tr.fastpfor && batched_applicable(codec, false, total_vals, nlists) && (double)total_vals >= 0.95 * (double)tr.fp_block * (double)nlists;
}

// Single-pass encode of nlists lists; writes d_out_off[0..nlists] (bytes) and the stream.  *total_out = bytes needed;
// MCP_E_CAP when that exceeds cap (nothing useful is written then); MCP_E_UNSUPPORTED when a batch did not fit the
// FastPFOR kernel's staging areas (the caller then takes the warp-per-list kernels).
int codec_fast_encode(mcp_ctx *ctx, int codec_id, int framing, const int32_t *d_vals, const int64_t *d_list_off,
                      int64_t uniform_len, int64_t nlists, int64_t total_vals, int64_t *d_out_off, uint8_t *d_out, int64_t cap,
                      int64_t *total_out)
{
    if (nlists == 0) {
        MCP_CUDA(ctx, cudaMemsetAsync(d_out_off, 0, sizeof(int64_t), ctx->stream));
        *total_out = 0;
        return MCP_OK;
    }
    FastEncArgs a{};
    a.codec = codec_id & MCP_CODEC_MASK;
    a.delta = (codec_id & MCP_FLAG_DELTA) != 0;
    a.app = framing == MCP_FRAME_APP;
    a.vals = reinterpret_cast<const uint32_t *>(d_vals);
    a.list_off = d_list_off;
    a.uniform_len = uniform_len;
    a.nlists = nlists;
    a.out_off = d_out_off;
    a.out = d_out;
    a.cap = d_out ? cap : 0;
    const bool batched = batched_applicable(a.codec, a.delta, total_vals, nlists);
    const bool flat_only = flat_only_codec(ctx, a.codec, a.delta, total_vals, nlists);
    const bool shortl = flat_only || (shortlist_applicable(ctx, a.codec, a.delta, total_vals, nlists, true) && !flat_pages_common(a.codec, total_vals, nlists));
    if (!batched && !shortl) return MCP_E_UNSUPPORTED;
    if (codec_traits(a.codec).xorbp) a.delta = true; // what the flat kernels code: the integrated tail's gaps
    // pass 0: thread-per-list kernel (short lists); pass 1: thread-per-block kernels
    for (int pass = shortl ? 0 : 1; pass < 2; ++pass) {
        if (pass == 1 && !batched) { ++ctx->fastEncodeFallbacks; return MCP_E_UNSUPPORTED; }
        // lists that are all VariableByte (FastPFOR / FastPFOR128 below one block, VariableByte alone) take the flat kernel
        const bool flat = pass == 0 && !codec_traits(a.codec).blocks && ctx->flatCodec;
        if (pass == 0 && flat_only && !flat) return MCP_E_UNSUPPORTED;
        a.lists_per_batch = pass == 0 ? (flat ? vbf_lists_per_tile(total_vals, nlists, 0.93 * VB4::SPAN, VB4::MAXL) : 32)
                                      : fast_lists_per_batch(a.codec, total_vals, nlists, uniform_len);
        a.nbatches = (nlists + a.lists_per_batch - 1) / a.lists_per_batch;
        if (a.nbatches > 0x7fffffff) return fail(ctx, MCP_E_ARG, "encode: too many lists for one call");
        // scratch: [ticket u64][total i64][err i32][pad][deferred tiles i32][pad][state u64 x nbatches][deferred tile ids i32 x nbatches]
        const size_t sbytes = 32 + (size_t)a.nbatches * (sizeof(uint64_t) + sizeof(int));
        MCP_CUDA(ctx, ctx->dScanTmp.reserve(sbytes));
        MCP_CUDA(ctx, cudaMemsetAsync(ctx->dScanTmp.p, 0, sbytes, ctx->stream));
        unsigned char *sp = ctx->dScanTmp.as<unsigned char>();
        a.ticket = reinterpret_cast<unsigned long long *>(sp);
        a.total = reinterpret_cast<int64_t *>(sp + 8);
        a.err = reinterpret_cast<int *>(sp + 16);
        a.state = reinterpret_cast<uint64_t *>(sp + 32);
        {
            KScope ks(ctx, MCP_K_ENCODE);
            const CodecTraits tr = codec_traits(a.codec);
            if (pass == 0 && flat) {
                // size pass, scan of the tiles' word counts (in place, in the state words), emit pass
                // lists that reach a FastPFOR block are possible but rare at this average length (FastPFOR128 on ~100-value
                // lists): their tiles go to k_vb4_pages, two more launches; when they cannot occur short of an outlier
                // (FastPFOR on the same lists) the launches are saved and an outlier costs the rerun below
                const bool pages = flat_pages_likely(tr, total_vals, nlists);
                int *defer_cnt = pages ? reinterpret_cast<int *>(sp + 24) : nullptr;
                int *defer_tiles = pages ? reinterpret_cast<int *>(sp + 32 + (size_t)a.nbatches * sizeof(uint64_t)) : nullptr;
                const bool marker = tr.fastpfor || tr.xorbp || tr.optpfd; // a first stage that took nothing: Composition's 0 marker
                TileEncArgs ta{a, reinterpret_cast<int64_t *>(a.state), tr.sk ? 1 : (marker ? 2 : 0), tr.fastpfor ? tr.fp_block : (marker ? 128 : 0),
                               defer_cnt, defer_tiles};
                const unsigned pgrid = (unsigned)std::min<int64_t>(a.nbatches, (int64_t)ctx->prop.multiProcessorCount * 8);
                if (a.delta) k_vb4_size<true><<<(unsigned)a.nbatches, VB4_THREADS, 0, ctx->stream>>>(ta);
                else k_vb4_size<false><<<(unsigned)a.nbatches, VB4_THREADS, 0, ctx->stream>>>(ta);
                if (pages && tr.optpfd) k_vb4_pages<false, true><<<pgrid, VB4_THREADS, 0, ctx->stream>>>(ta);
                else if (pages) k_vb4_pages<false, false><<<pgrid, VB4_THREADS, 0, ctx->stream>>>(ta);
                k_vb4_scan<<<1, 1024, 0, ctx->stream>>>(ta.tile_words, a.nbatches, a.total, a.out_off + a.nlists, a.cap, a.err);
                if (a.delta) k_vb4_tile<VB4_THREADS, VB4_VPT, true, true><<<(unsigned)a.nbatches, VB4_THREADS, 0, ctx->stream>>>(ta);
                else k_vb4_tile<VB4_THREADS, VB4_VPT, false, true><<<(unsigned)a.nbatches, VB4_THREADS, 0, ctx->stream>>>(ta);
                if (pages && tr.optpfd) k_vb4_pages<true, true><<<pgrid, VB4_THREADS, 0, ctx->stream>>>(ta);
                else if (pages) k_vb4_pages<true, false><<<pgrid, VB4_THREADS, 0, ctx->stream>>>(ta);
            } else if (pass == 0) {
                const int64_t want = (a.nbatches + SL_WARPS - 1) / SL_WARPS, res = (int64_t)ctx->prop.multiProcessorCount * SL_CTAS;
                k_sl_encode<<<(unsigned)(want < res ? want : res), SL_WARPS * 32, 0, ctx->stream>>>(a);
            } else if (!tr.fastpfor) k_bp_encode_fast<<<(unsigned)a.nbatches, FE_THREADS, 0, ctx->stream>>>(a);
            else if (tr.fp_block == 256) k_fp_encode_fast<8><<<(unsigned)a.nbatches, FE_THREADS, 0, ctx->stream>>>(a);
            else k_fp_encode_fast<4><<<(unsigned)a.nbatches, FE_THREADS, 0, ctx->stream>>>(a);
            MCP_TRY(check_launch(ctx, "k_encode_fast"));
        }
        simulate_turn_end(ctx); // mcp_simulate with calls in flight: this call's last kernel is queued, the next call's may follow
        MCP_CUDA(ctx, cudaMemcpyAsync(ctx->hPinned, sp + 8, 16, cudaMemcpyDeviceToHost, ctx->stream));
        MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *total_out = *reinterpret_cast<int64_t *>(ctx->hPinned);
        const int flag = *reinterpret_cast<int *>(reinterpret_cast<char *>(ctx->hPinned) + 8);
        if (flag == 2) { // a batch did not fit this kernel: not an error, the next kind of kernel takes the call
            if (pass == 0) { ++ctx->shortListFallbacks; continue; }
            ++ctx->fastEncodeFallbacks;
            return MCP_E_UNSUPPORTED;
        }
        if (pass == 0) ++ctx->shortListCalls;
        ++ctx->fastEncodeCalls;
        if (flag != 0) return fail(ctx, MCP_E_CAP, "encode: output buffer too small");
        return MCP_OK;
    }
    return MCP_E_UNSUPPORTED;
}

int codec_fast_decode(mcp_ctx *ctx, int codec_id, int framing, const uint8_t *d_in, const int64_t *d_in_off, int64_t nlists,
                      int64_t total_out_vals, int32_t *d_out, const int64_t *d_out_off)
{
    if (nlists == 0) return MCP_OK;
    FastDecArgs a{};
    a.codec = codec_id & MCP_CODEC_MASK;
    a.delta = (codec_id & MCP_FLAG_DELTA) != 0;
    a.app = framing == MCP_FRAME_APP;
    a.in = d_in;
    a.in_off = d_in_off;
    a.nlists = nlists;
    a.out = reinterpret_cast<uint32_t *>(d_out);
    a.out_off = d_out_off;
    const bool batched = batched_applicable(a.codec, a.delta, total_out_vals, nlists);
    const bool flat_only = flat_only_codec(ctx, a.codec, a.delta, total_out_vals, nlists);
    const bool shortl = flat_only || (shortlist_applicable(ctx, a.codec, a.delta, total_out_vals, nlists, false) && !flat_pages_common(a.codec, total_out_vals, nlists));
    if (!batched && !shortl) return MCP_E_UNSUPPORTED;
    if (codec_traits(a.codec).xorbp) a.delta = true; // (see codec_fast_encode)
    MCP_CUDA(ctx, ctx->dErr.reserve(64));
    a.err = ctx->dErr.as<int>();
    for (int pass = shortl ? 0 : 1; pass < 2; ++pass) {
        if (pass == 1 && !batched) return MCP_E_UNSUPPORTED;
        MCP_CUDA(ctx, cudaMemsetAsync(ctx->dErr.p, 0, sizeof(int), ctx->stream));
        const bool flat = pass == 0 && !codec_traits(a.codec).blocks && ctx->flatCodec;
        if (pass == 0 && flat_only && !flat) return MCP_E_UNSUPPORTED;
        a.lists_per_batch = pass == 0 ? (flat ? vbf_lists_per_tile(total_out_vals, nlists, 0.88 * D4_VSPAN, D4_MAXL) : 32)
                                      : fast_lists_per_batch(a.codec, total_out_vals, nlists);
        const int64_t nbatches = (nlists + a.lists_per_batch - 1) / a.lists_per_batch;
        {
            KScope ks(ctx, MCP_K_DECODE);
            const CodecTraits tr = codec_traits(a.codec);
            if (pass == 0 && flat) {
                const bool marker = tr.fastpfor || tr.xorbp || tr.optpfd;
                const int hdr_mode = tr.sk ? 1 : (marker ? 2 : 0), page = tr.fastpfor ? tr.fp_block : (marker ? 128 : 0);
                int *defer_cnt = nullptr, *defer_tiles = nullptr;
                if (flat_pages_likely(tr, total_out_vals, nlists)) { // (see codec_fast_encode)
                    MCP_CUDA(ctx, ctx->dScanTmp.reserve(16 + (size_t)nbatches * sizeof(int)));
                    MCP_CUDA(ctx, cudaMemsetAsync(ctx->dScanTmp.p, 0, 16, ctx->stream));
                    defer_cnt = ctx->dScanTmp.as<int>();
                    defer_tiles = defer_cnt + 4;
                }
                if (a.delta) k_vb4_decode<true><<<(unsigned)nbatches, D4_THREADS, 0, ctx->stream>>>(a, hdr_mode, page, defer_cnt, defer_tiles);
                else k_vb4_decode<false><<<(unsigned)nbatches, D4_THREADS, 0, ctx->stream>>>(a, hdr_mode, page, defer_cnt, defer_tiles);
                const unsigned pgrid = (unsigned)std::min<int64_t>(nbatches, (int64_t)ctx->prop.multiProcessorCount * 8);
                if (defer_cnt && tr.optpfd) k_vb4_decode_pages<true><<<pgrid, D4_THREADS, 0, ctx->stream>>>(a, defer_cnt, defer_tiles);
                else if (defer_cnt) k_vb4_decode_pages<false><<<pgrid, D4_THREADS, 0, ctx->stream>>>(a, defer_cnt, defer_tiles);
            } else if (pass == 0) {
                const int64_t res = (int64_t)ctx->prop.multiProcessorCount * SL_DCTAS;
                k_sl_decode<<<(unsigned)(nbatches < res ? nbatches : res), 32, 0, ctx->stream>>>(a, nbatches);
            } else if (!tr.fastpfor) k_bp_decode_fast<<<(unsigned)nbatches, FE_THREADS, 0, ctx->stream>>>(a);
            else if (tr.fp_block == 256) k_fp_decode_fast<8><<<(unsigned)nbatches, FE_THREADS, 0, ctx->stream>>>(a);
            else k_fp_decode_fast<4><<<(unsigned)nbatches, FE_THREADS, 0, ctx->stream>>>(a);
            MCP_TRY(check_launch(ctx, "k_decode_fast"));
        }
        MCP_CUDA(ctx, cudaMemcpyAsync(ctx->hPinned, ctx->dErr.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const int flag = *reinterpret_cast<int *>(ctx->hPinned);
        if (flag == 2) { // a batch did not fit this kernel: the next kind of kernel (or the warp-per-list one) takes the call
            if (pass == 0) { ++ctx->shortListFallbacks; continue; }
            return MCP_E_UNSUPPORTED;
        }
        if (pass == 0) ++ctx->shortListCalls;
        if (flag != 0) {
            char msg[96];
            snprintf(msg, sizeof msg, "decode: malformed compressed stream (kernel pass %d, check %d)", pass, flag);
            return fail(ctx, MCP_E_FORMAT, msg);
        }
        return MCP_OK;
    }
    return MCP_E_UNSUPPORTED;
}

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
        if (codec_traits(a.codec).optpfd) k_encode<false, true><<<grid_for_lists(ctx, nlists), kEncWarps * 32, 0, ctx->stream>>>(a);
        else k_encode<false, false><<<grid_for_lists(ctx, nlists), kEncWarps * 32, 0, ctx->stream>>>(a);
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
    if (codec_traits(a.codec).optpfd) k_encode<true, true><<<grid_for_lists(ctx, nlists), kEncWarps * 32, 0, ctx->stream>>>(a);
    else k_encode<true, false><<<grid_for_lists(ctx, nlists), kEncWarps * 32, 0, ctx->stream>>>(a);
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
