/*
 * codec_oracle.c -- CPU restatement of the JavaFastPFOR codecs on the hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Written from the wire-format rules of
 * the vendored library (JavaFastPFOR 0.1.8-SNAPSHOT under /root/reference), one
 * generic width-parameterised bit loop instead of the library's generated unrolled
 * bodies.  All arithmetic is on uint32_t so that Java's int wrap-around and `>>>`
 * are reproduced exactly.  Citations are relative to
 * /root/reference/JavaFastPFOR/src/main/java/me/lemire/integercompression/.
 */
#include "oracle.h"

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

/* ------------------------------------------------------------------ Util */

/* Util.java:101-103  bits(i) = 32 - numberOfLeadingZeros(i) */
int orc_bits(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }

/* Util.java:28-33  OR-reduce then bits() */
int orc_maxbits(const uint32_t *in, int n)
{
    uint32_t m = 0;
    for (int i = 0; i < n; ++i) m |= in[i];
    return orc_bits(m);
}

/* Util.java:85-92  OR-reduce of successive differences (wrap-around) */
int orc_maxdiffbits(uint32_t init, const uint32_t *in, int n)
{
    uint32_t m = in[0] - init;
    for (int i = 1; i < n; ++i) m |= in[i] - in[i - 1];
    return orc_bits(m);
}

/* ------------------------------------------------------------ BitPacking
 * 32 values <-> b words.  Value i sits at bit offset i*b of a little-endian-by-word
 * bit stream (BitPacking.java:191-233 for b=10 shows the pattern; b=0 writes
 * nothing :150-153; b=32 is a copy :1449-1452).
 */
static void pack_stream(const uint32_t *v, uint32_t *out, int b, uint32_t mask)
{
    if (b == 0) return;
    if (b == 32) { memcpy(out, v, 32 * sizeof(uint32_t)); return; }
    for (int w = 0; w < b; ++w) out[w] = 0;
    for (int i = 0; i < 32; ++i) {
        const uint32_t x = v[i] & mask;
        const int bit = i * b, w = bit >> 5, s = bit & 31;
        out[w] |= x << s;
        if (s + b > 32) out[w + 1] |= x >> (32 - s);
    }
}

/* masked variant: only the low b bits of each input survive (BitPacking.java:42-148) */
void orc_fastpack(const uint32_t *in, uint32_t *out, int b)
{
    pack_stream(in, out, b, b >= 32 ? 0xffffffffu : ((1u << b) - 1u));
}

/*
 * unmasked variant (BitPacking.java:1706-1812): the caller guarantees the values
 * fit in b bits.  If they do not, the Java code ORs the stray high bits into the
 * following fields of the same word and `>>>`-spills them; reproduce that too so
 * the golden vectors for out-of-range inputs also match.
 */
void orc_fastpackwithoutmask(const uint32_t *in, uint32_t *out, int b)
{
    if (b == 0) return;
    if (b == 32) { memcpy(out, in, 32 * sizeof(uint32_t)); return; }
    for (int w = 0; w < b; ++w) out[w] = 0;
    for (int i = 0; i < 32; ++i) {
        const uint32_t x = in[i];
        const int bit = i * b, w = bit >> 5, s = bit & 31;
        out[w] |= x << s;
        if (s + b > 32) out[w + 1] |= x >> (32 - s);
    }
}

/* BitPacking.java:3021-3127; b=0 yields zeros (:3129-3132) */
void orc_fastunpack(const uint32_t *in, uint32_t *out, int b)
{
    if (b == 0) { memset(out, 0, 32 * sizeof(uint32_t)); return; }
    if (b == 32) { memcpy(out, in, 32 * sizeof(uint32_t)); return; }
    const uint32_t mask = (1u << b) - 1u;
    for (int i = 0; i < 32; ++i) {
        const int bit = i * b, w = bit >> 5, s = bit & 31;
        uint32_t x = in[w] >> s;
        if (s + b > 32) x |= in[w + 1] << (32 - s);
        out[i] = x & mask;
    }
}

/* IntegratedBitPacking.java:39-145: pack in[i]-in[i-1] unmasked; b=32 stores RAW values (:1446-1451) */
void orc_integratedpack(uint32_t init, const uint32_t *in, uint32_t *out, int b)
{
    if (b == 0) return;
    if (b == 32) { memcpy(out, in, 32 * sizeof(uint32_t)); return; }
    uint32_t d[32];
    d[0] = in[0] - init;
    for (int i = 1; i < 32; ++i) d[i] = in[i] - in[i - 1];
    orc_fastpackwithoutmask(d, out, b);
}

/* IntegratedBitPacking.java:1815-1921: prefix-sum while unpacking; b=0 repeats init (:1810-1813) */
void orc_integratedunpack(uint32_t init, const uint32_t *in, uint32_t *out, int b)
{
    if (b == 32) { memcpy(out, in, 32 * sizeof(uint32_t)); return; }
    uint32_t d[32];
    orc_fastunpack(in, d, b);
    uint32_t prev = init;
    for (int i = 0; i < 32; ++i) { prev += d[i]; out[i] = prev; }
}

/* ------------------------------------------------------------------ Delta */

void orc_delta(uint32_t *data, size_t n) /* Delta.java:24-28 */
{
    for (size_t i = n; i-- > 1;) data[i] -= data[i - 1];
}

void orc_inverse_delta(uint32_t *data, size_t n) /* Delta.java:84-88 */
{
    for (size_t i = 1; i < n; ++i) data[i] += data[i - 1];
}

/* ----------------------------------------------------------- VariableByte
 * VariableByte.java:38-77.  7-bit groups low first, bit 7 set on the LAST byte of a
 * value; the byte stream is zero-padded to a multiple of 4 and stored as
 * little-endian words.  `integrated` subtracts a running previous value first
 * (IntegratedVariableByte.java:184-225).
 */
static int64_t vb_encode(const uint32_t *in, int64_t n, uint32_t *out, int integrated, uint32_t *prev_io)
{
    if (n == 0) return 0;
    uint8_t *bytes = (uint8_t *)out; /* little-endian host: byte k of the stream is byte k of out[] */
    int64_t nb = 0;
    uint32_t prev = integrated ? *prev_io : 0;
    for (int64_t i = 0; i < n; ++i) {
        uint32_t v = in[i];
        if (integrated) { v = in[i] - prev; prev = in[i]; }
        while (v >= 128u) { bytes[nb++] = (uint8_t)(v & 127u); v >>= 7; }
        bytes[nb++] = (uint8_t)(v | 128u);
    }
    while (nb & 3) bytes[nb++] = 0;
    if (integrated) *prev_io = prev;
    return nb / 4;
}

/* VariableByte.uncompress :115-138 -- consumes exactly inlen words */
static int64_t vb_decode_all(const uint32_t *in, int64_t inlen, uint32_t *out, int64_t out_cap, int integrated)
{
    int64_t o = 0;
    uint32_t v = 0, prev = 0;
    int shift = 0;
    for (int64_t p = 0; p < inlen; ++p) {
        for (int s = 0; s < 32; s += 8) {
            const uint32_t c = (in[p] >> s) & 0xffu;
            v += (c & 127u) << (shift & 31); /* Java shifts are mod 32 */
            if (c & 128u) {
                if (o >= out_cap) return -1;
                if (integrated) { prev += v; out[o++] = prev; } else out[o++] = v;
                v = 0; shift = 0;
            } else
                shift += 7;
        }
    }
    return o;
}

/* VariableByte.headlessUncompress :180-203 -- stops after num values, rounds inpos up */
static int64_t vb_decode_n(const uint32_t *in, uint32_t *out, int64_t num, int integrated, uint32_t *prev_io)
{
    int64_t p = 0, o = 0;
    int s = 0, shift = 0;
    uint32_t v = 0, prev = integrated ? *prev_io : 0;
    while (o < num) {
        const uint32_t c = (in[p] >> s) & 0xffu;
        s += 8; p += s >> 5; s &= 31;
        v += (c & 127u) << (shift & 31);
        if (c & 128u) {
            if (integrated) { prev += v; out[o++] = prev; } else out[o++] = v;
            v = 0; shift = 0;
        } else
            shift += 7;
    }
    if (integrated) *prev_io = prev;
    return p + (s != 0 ? 1 : 0);
}

/* ---------------------------------------------------------- BinaryPacking
 * BinaryPacking.java:54-89 / IntegratedBinaryPacking.java:80-125.  n is already a
 * multiple of 32.  Groups of four blocks share one header word with the four widths
 * in its bytes (MSB = first block); leftover blocks get a whole header word each.
 */
static int64_t bp32_encode(const uint32_t *in, int64_t n, uint32_t *out, int integrated, uint32_t *init_io)
{
    int64_t o = 0, s = 0;
    uint32_t init = integrated ? *init_io : 0;
    for (; s + 128 <= n; s += 128) {
        int b[4];
        uint32_t ini[4];
        for (int j = 0; j < 4; ++j) {
            ini[j] = (j == 0) ? init : in[s + 32 * j - 1];
            b[j] = integrated ? orc_maxdiffbits(ini[j], in + s + 32 * j, 32) : orc_maxbits(in + s + 32 * j, 32);
        }
        out[o++] = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | (uint32_t)b[3];
        for (int j = 0; j < 4; ++j) {
            if (integrated) orc_integratedpack(ini[j], in + s + 32 * j, out + o, b[j]);
            else orc_fastpackwithoutmask(in + s + 32 * j, out + o, b[j]);
            o += b[j];
        }
        init = in[s + 127];
    }
    for (; s < n; s += 32) {
        const int b = integrated ? orc_maxdiffbits(init, in + s, 32) : orc_maxbits(in + s, 32);
        out[o++] = (uint32_t)b;
        if (integrated) orc_integratedpack(init, in + s, out + o, b);
        else orc_fastpackwithoutmask(in + s, out + o, b);
        o += b;
        init = in[s + 31];
    }
    if (integrated && n > 0) *init_io = in[n - 1];
    return o;
}

/* BinaryPacking.java:102-133 / IntegratedBinaryPacking.java:128-170; returns words consumed */
static int64_t bp32_decode(const uint32_t *in, uint32_t *out, int64_t n, int integrated, uint32_t *init_io)
{
    int64_t p = 0, s = 0;
    uint32_t init = integrated ? *init_io : 0;
    for (; s + 128 <= n; s += 128) {
        const uint32_t h = in[p++];
        const int b[4] = { (int)(h >> 24), (int)((h >> 16) & 255u), (int)((h >> 8) & 255u), (int)(h & 255u) };
        for (int j = 0; j < 4; ++j) {
            if (integrated) { orc_integratedunpack(init, in + p, out + s + 32 * j, b[j]); init = out[s + 32 * j + 31]; }
            else orc_fastunpack(in + p, out + s + 32 * j, b[j]);
            p += b[j];
        }
    }
    for (; s < n; s += 32) {
        const int b = (int)in[p++];
        if (integrated) { orc_integratedunpack(init, in + p, out + s, b); init = out[s + 31]; }
        else orc_fastunpack(in + p, out + s, b);
        p += b;
    }
    if (integrated) *init_io = init;
    return p;
}

/* ---------------------------------------------------------------- FastPFOR
 * FastPFOR.java:41-332 (BLOCK_SIZE 256) / FastPFOR128.java (BLOCK_SIZE 128).
 * The per-width exception buffers persist for the life of the instance and are
 * never cleared (only the fill pointers are, :143), which leaks stale values into
 * the padding of the last packed group of every exception stream.
 */
#define FP_PAGE 65536
#define FP_OVERHEAD_OF_EACH_EXCEPT 8

struct orc_fastpfor {
    int block;
    uint32_t *dtp[33]; /* dataTobePacked */
    int64_t cap[33];
    int64_t ptr[33];   /* dataPointers   */
    uint8_t *meta;     /* byteContainer  */
};

orc_fastpfor *orc_fastpfor_new(int block_size)
{
    if (block_size != 256 && block_size != 128) return NULL;
    orc_fastpfor *f = (orc_fastpfor *)calloc(1, sizeof(*f));
    f->block = block_size;
    for (int k = 1; k <= 32; ++k) { /* :74-75  new int[pageSize / 32 * 4] */
        f->cap[k] = FP_PAGE / 32 * 4;
        f->dtp[k] = (uint32_t *)calloc((size_t)f->cap[k], sizeof(uint32_t));
    }
    f->meta = (uint8_t *)malloc(3 * FP_PAGE / block_size + FP_PAGE + 8); /* :71-72 */
    return f;
}

void orc_fastpfor_free(orc_fastpfor *f)
{
    if (!f) return;
    for (int k = 1; k <= 32; ++k) free(f->dtp[k]);
    free(f->meta);
    free(f);
}

/* FastPFOR.java:107-134 */
static void fp_best_b(const uint32_t *in, int block, int *bestb, int *bestc, int *maxb)
{
    int freqs[33] = { 0 };
    for (int i = 0; i < block; ++i) freqs[orc_bits(in[i])]++;
    int b = 32;
    while (freqs[b] == 0) b--;
    *maxb = b; *bestb = b; *bestc = 0;
    int bestcost = b * block, c = 0;
    for (int cand = b - 1; cand >= 0; --cand) {
        c += freqs[cand + 1];
        if (c == block) break;
        int cost = c * FP_OVERHEAD_OF_EACH_EXCEPT + c * (*maxb - cand) + cand * block + 8;
        if (*maxb - cand == 1) cost -= c;
        if (cost < bestcost) { bestcost = cost; *bestb = cand; *bestc = c; }
    }
}

/* FastPFOR.java:136-212 */
static int64_t fp_encode_page(orc_fastpfor *f, const uint32_t *in, int64_t n, uint32_t *out)
{
    const int B = f->block;
    int64_t o = 1; /* out[0] = header, filled below */
    int64_t nmeta = 0;
    memset(f->ptr, 0, sizeof(f->ptr));
    for (int64_t s = 0; s + B <= n; s += B) {
        int b, c, maxb;
        fp_best_b(in + s, B, &b, &c, &maxb);
        f->meta[nmeta++] = (uint8_t)b;
        f->meta[nmeta++] = (uint8_t)c;
        if (c > 0) {
            f->meta[nmeta++] = (uint8_t)maxb;
            const int idx = maxb - b;
            if (f->ptr[idx] + c >= f->cap[idx]) { /* :156-164 grow, preserving contents */
                int64_t ns = 2 * (f->ptr[idx] + c);
                ns = (ns + 31) - (ns + 31) % 32;
                f->dtp[idx] = (uint32_t *)realloc(f->dtp[idx], (size_t)ns * sizeof(uint32_t));
                memset(f->dtp[idx] + f->cap[idx], 0, (size_t)(ns - f->cap[idx]) * sizeof(uint32_t));
                f->cap[idx] = ns;
            }
            for (int k = 0; k < B; ++k) {
                const uint32_t hi = (b >= 32) ? 0u : (in[s + k] >> b);
                if (hi != 0) {
                    f->meta[nmeta++] = (uint8_t)k;
                    f->dtp[idx][f->ptr[idx]++] = hi;
                }
            }
        }
        for (int k = 0; k < B; k += 32) { orc_fastpack(in + s + k, out + o, b); o += b; }
    }
    out[0] = (uint32_t)o;                 /* :182 offset of the metadata */
    out[o++] = (uint32_t)nmeta;           /* :186 bytesize (before padding) */
    {
        int64_t padded = nmeta;
        while (padded & 3) f->meta[padded++] = 0;
        memcpy(out + o, f->meta, (size_t)padded); /* little-endian words */
        o += padded / 4;
    }
    uint32_t bitmap = 0;
    for (int k = 2; k <= 32; ++k)
        if (f->ptr[k] != 0) bitmap |= 1u << (k - 1);
    out[o++] = bitmap;
    for (int k = 2; k <= 32; ++k) {
        if (f->ptr[k] == 0) continue;
        out[o++] = (uint32_t)f->ptr[k];
        int64_t j = 0;
        for (; j < f->ptr[k]; j += 32) { orc_fastpack(f->dtp[k] + j, out + o, k); o += k; }
        const int64_t overflow = j - f->ptr[k];
        o -= overflow * k / 32;
    }
    return o;
}

/* FastPFOR.java:233-307; returns words consumed */
static int64_t fp_decode_page(orc_fastpfor *f, const uint32_t *in, uint32_t *out, int64_t n)
{
    const int B = f->block;
    const int64_t wheremeta = in[0];
    int64_t e = wheremeta;
    const int64_t bytesize = in[e++];
    const uint8_t *meta = (const uint8_t *)(in + e);
    e += (bytesize + 3) / 4;
    const uint32_t bitmap = in[e++];
    for (int k = 2; k <= 32; ++k) {
        if (!(bitmap & (1u << (k - 1)))) continue;
        const int64_t size = in[e++];
        const int64_t rounded = (size + 31) - (size + 31) % 32;
        if (f->cap[k] < rounded) {
            free(f->dtp[k]);
            f->dtp[k] = (uint32_t *)calloc((size_t)rounded, sizeof(uint32_t));
            f->cap[k] = rounded;
        }
        /* unpack group by group; the last group may read past the stream's end in the
         * Java code too (it copies into a padded buffer, :262-272) -- emulate with a
         * bounded copy so we never read words that are not ours. */
        const int64_t words = (size * k + 31) / 32;
        int64_t j = 0, w = 0;
        for (; j < size; j += 32, w += k) {
            uint32_t tmp[32];
            const int64_t avail = words - w < k ? words - w : k;
            memset(tmp, 0, sizeof(tmp));
            memcpy(tmp, in + e + w, (size_t)avail * sizeof(uint32_t));
            orc_fastunpack(tmp, f->dtp[k] + j, k);
        }
        e += words;
    }
    memset(f->ptr, 0, sizeof(f->ptr));
    int64_t p = 1, m = 0;
    for (int64_t s = 0; s + B <= n; s += B) {
        const int b = meta[m++];
        const int c = meta[m++];
        for (int k = 0; k < B; k += 32) { orc_fastunpack(in + p, out + s + k, b); p += b; }
        if (c > 0) {
            const int maxb = meta[m++];
            const int idx = maxb - b;
            for (int k = 0; k < c; ++k) {
                const int pos = meta[m++];
                if (idx == 1) out[s + pos] |= 1u << b;
                else out[s + pos] |= f->dtp[idx][f->ptr[idx]++] << b;
            }
        }
    }
    return e;
}

/* FastPFOR.compress :309-317 + headlessCompress :92-105 */
int64_t orc_fastpfor_compress(orc_fastpfor *f, const int32_t *in_, int64_t n, int32_t *out_)
{
    const uint32_t *in = (const uint32_t *)in_;
    uint32_t *out = (uint32_t *)out_;
    const int64_t m = n - n % f->block;
    if (m == 0) return 0;
    int64_t o = 0;
    out[o++] = (uint32_t)m;
    for (int64_t s = 0; s < m;) {
        const int64_t sz = (m - s < FP_PAGE) ? m - s : FP_PAGE;
        o += fp_encode_page(f, in + s, sz, out + o);
        s += sz;
    }
    return o;
}

/* FastPFOR.uncompress :320-327 + headlessUncompress :222-231; returns ints produced */
int64_t orc_fastpfor_uncompress(orc_fastpfor *f, const int32_t *in_, int64_t inlen, int32_t *out_, int64_t *consumed)
{
    const uint32_t *in = (const uint32_t *)in_;
    uint32_t *out = (uint32_t *)out_;
    *consumed = 0;
    if (inlen == 0) return 0;
    int64_t m = in[0];
    m -= m % f->block;
    int64_t p = 1;
    for (int64_t s = 0; s < m;) {
        const int64_t sz = (m - s < FP_PAGE) ? m - s : FP_PAGE;
        p += fp_decode_page(f, in + p, out + s, sz);
        s += sz;
    }
    *consumed = p;
    return m;
}

/* --------------------------------------------------------- XorBinaryPacking
 * differential/XorBinaryPacking.java:31-71 (compress), :73-111 (uncompress), :118-127 (xorMaxBits), :129-139 (xorPack).
 * Blocks of 128: one header word (bits1<<24 | bits2<<16 | bits3<<8 | bits4), then four sub-blocks of 32 values, each
 * value XOR-ed with its predecessor (the first one of the list with 0) and packed with fastpackwithoutmask at the
 * sub-block's width = bits(OR of the XORs).  `ctx_io`: the last value seen (0 at the start). */
static int64_t xor128_encode(const uint32_t *in, int64_t m, uint32_t *out, uint32_t *ctx_io)
{
    int64_t o = 0;
    uint32_t ctx = *ctx_io;
    for (int64_t ip = 0; ip < m; ip += 128) {
        uint32_t work[4][32];
        int bits[4];
        for (int q = 0; q < 4; ++q) {
            uint32_t mask = 0, prev = q == 0 ? ctx : in[ip + 32 * q - 1];
            for (int i = 0; i < 32; ++i) { work[q][i] = in[ip + 32 * q + i] ^ prev; prev = in[ip + 32 * q + i]; mask |= work[q][i]; }
            bits[q] = orc_bits(mask);
        }
        out[o++] = ((uint32_t)bits[0] << 24) | ((uint32_t)bits[1] << 16) | ((uint32_t)bits[2] << 8) | (uint32_t)bits[3];
        for (int q = 0; q < 4; ++q) { orc_fastpackwithoutmask(work[q], out + o, bits[q]); o += bits[q]; }
        ctx = in[ip + 127];
    }
    *ctx_io = ctx;
    return o;
}

static int64_t xor128_decode(const uint32_t *in, uint32_t *out, int64_t m)
{
    int64_t p = 0;
    uint32_t ctx = 0;
    for (int64_t op = 0; op < m; op += 128) {
        const uint32_t h = in[p++];
        const int bits[4] = { (int)(h >> 24), (int)((h >> 16) & 0xFF), (int)((h >> 8) & 0xFF), (int)(h & 0xFF) };
        for (int q = 0; q < 4; ++q) {
            uint32_t work[32];
            orc_fastunpack(in + p, work, bits[q]);
            p += bits[q];
            for (int i = 0; i < 32; ++i) { ctx = work[i] ^ ctx; out[op + 32 * q + i] = ctx; }
        }
    }
    return p;
}

/* ------------------------------------------------------------------- S16
 * S16.java (Simple16 as adapted for NewPFD / OptPFD): 16 layouts of up to 28 values in the low 28 bits of a word, the
 * layout number in the top 4.  compressblock (:80-100) tries the layouts in order and takes the first one whose widths
 * hold the next min(S16_NUM, n) values; comparisons are on Java ints (signed). */
static const int S16_NUM[16] = { 28, 21, 21, 21, 14, 9, 8, 7, 6, 6, 5, 5, 4, 3, 2, 1 };
static const int S16_BITS[16][28] = {
    { 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1 },
    { 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1 },
    { 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1 },
    { 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2 },
    { 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2 },
    { 4, 3, 3, 3, 3, 3, 3, 3, 3 }, { 3, 4, 4, 4, 4, 3, 3, 3 },
    { 4, 4, 4, 4, 4, 4, 4 }, { 5, 5, 5, 5, 4, 4 },
    { 4, 4, 5, 5, 5, 5 }, { 6, 6, 6, 5, 5 }, { 5, 5, 6, 6, 6 },
    { 7, 7, 7, 7 }, { 10, 9, 9 }, { 14, 14 }, { 28 } };

/* one word: returns how many values it took (S16.compressblock / fakecompressblock); -1 = "Too big a number" */
static int s16_block(const int32_t *in, int n, uint32_t *word)
{
    for (int idx = 0; idx < 16; ++idx) {
        uint32_t w = (uint32_t)idx << 28;
        const int num = S16_NUM[idx] < n ? S16_NUM[idx] : n;
        int j = 0, bits = 0;
        while (j < num && in[j] < (int32_t)(1u << S16_BITS[idx][j])) {
            w |= (uint32_t)in[j] << bits;
            bits += S16_BITS[idx][j];
            ++j;
        }
        if (j == num) { if (word) *word = w; return num; }
    }
    return -1;
}

/* S16.compress (:30-44) / estimatecompress (:57-72): words written (out may be NULL to only count) */
static int s16_compress(const int32_t *in, int n, uint32_t *out)
{
    int pos = 0, words = 0;
    while (pos < n) {
        const int took = s16_block(in + pos, n - pos, out ? out + words : NULL);
        if (took < 0) return -1;
        pos += took;
        ++words;
    }
    return words;
}

/* S16.uncompress (:135-147) over `nwords` words, at most `n` values */
static void s16_uncompress(const uint32_t *in, int nwords, int32_t *out, int n)
{
    int pos = 0;
    for (int w = 0; w < nwords; ++w) {
        const int idx = (int)(in[w] >> 28);
        const int num = S16_NUM[idx] < n ? S16_NUM[idx] : n;
        int bits = 0;
        for (int j = 0; j < num; ++j) {
            out[pos + j] = (int32_t)((in[w] >> bits) & (0xffffffffu >> (32 - S16_BITS[idx][j])));
            bits += S16_BITS[idx][j];
        }
        n -= num;
        pos += num;
    }
}

/* ----------------------------------------------------------------- OptPFD
 * OptPFD.java:35-200.  Blocks of 128; per block one header word b | c << 8 | s << 16 (b = INDEX into `bits`, c =
 * exceptions, s = S16 words), the S16-coded exceptions (c high parts, then c positions), then 4 x fastpack at bits[b].
 * getBestBFromData (:61-98) as written: `mini` is computed from BIT WIDTHS (bits[invbits[mb]] - 28) but used as the first
 * INDEX of the search; the cost of a candidate is 4 bits[i] + S16 words against a start of 4 * 32; ties go to the later
 * candidate (<=). */
static const int OPT_BITS[17] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16, 20, 32 };
static const int OPT_INVBITS[33] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 14, 14, 15, 15, 15, 15, 16, 16, 16, 16, 16, 16, 16, 16,
                                     16, 16, 16, 16 };

static void optpfd_best(const uint32_t *in, int *bestb, int *bestc)
{
    int32_t exbuf[256];
    const int mb = orc_maxbits(in, 128);
    int mini = 0;
    if (mini + 28 < OPT_BITS[OPT_INVBITS[mb]]) mini = OPT_BITS[OPT_INVBITS[mb]] - 28;
    int besti = 16, bestcost = OPT_BITS[16] * 4, exceptcounter = 0;
    for (int i = mini; i < 16; ++i) {
        int tmp = 0;
        for (int k = 0; k < 128; ++k) tmp += (in[k] >> OPT_BITS[i]) != 0;
        if (tmp == 128) continue;
        for (int k = 0, c = 0; k < 128; ++k)
            if ((in[k] >> OPT_BITS[i]) != 0) { exbuf[tmp + c] = k; exbuf[c] = (int32_t)(in[k] >> OPT_BITS[i]); ++c; }
        const int thiscost = OPT_BITS[i] * 4 + s16_compress(exbuf, 2 * tmp, NULL);
        if (thiscost <= bestcost) { bestcost = thiscost; besti = i; exceptcounter = tmp; }
    }
    *bestb = besti;
    *bestc = exceptcounter;
}

static int64_t optpfd_encode(const uint32_t *in, int64_t m, uint32_t *out)
{
    int64_t o = 0;
    for (int64_t ip = 0; ip < m; ip += 128) {
        int b, c, s = 0;
        optpfd_best(in + ip, &b, &c);
        const int64_t remember = o++;
        if (c > 0) {
            int32_t exbuf[256];
            int cc = 0;
            for (int i = 0; i < 128; ++i)
                if ((in[ip + i] >> OPT_BITS[b]) != 0) { exbuf[cc + c] = i; exbuf[cc] = (int32_t)(in[ip + i] >> OPT_BITS[b]); ++cc; }
            s = s16_compress(exbuf, 2 * c, out + o);
            o += s;
        }
        out[remember] = (uint32_t)b | ((uint32_t)c << 8) | ((uint32_t)s << 16);
        for (int k = 0; k < 128; k += 32) { orc_fastpack(in + ip + k, out + o, OPT_BITS[b]); o += OPT_BITS[b]; }
    }
    return o;
}

static int64_t optpfd_decode(const uint32_t *in, uint32_t *out, int64_t m)
{
    int64_t p = 0;
    for (int64_t op = 0; op < m; op += 128) {
        const uint32_t h = in[p++];
        const int b = (int)(h & 0xFF), c = (int)((h >> 8) & 0xFF), s = (int)(h >> 16);
        int32_t exbuf[256 + 28];
        if (b > 16) return -1;
        s16_uncompress(in + p, s, exbuf, 2 * c);
        p += s;
        for (int k = 0; k < 128; k += 32) { orc_fastunpack(in + p, out + op + k, OPT_BITS[b]); p += OPT_BITS[b]; }
        for (int k = 0; k < c; ++k) out[op + (exbuf[k + c] & 127)] |= (uint32_t)exbuf[k] << (OPT_BITS[b] & 31);
    }
    return p;
}

/* ------------------------------------------------------------ compositions */

int64_t orc_max_compressed_words(int64_t n)
{
    /* worst case: every value needs 5 VByte bytes or 32 packed bits + headers;
     * the reference's own scratch is 4n+1024 (LoadPortfolios.java:611) */
    return 2 * n + 1024;
}

/* Composition.compress (Composition.java:37-51) for F1 in {BinaryPacking, FastPFOR[128]}, F2 = VariableByte;
 * IntegratedComposition.compress (IntegratedComposition.java:41-55). */
int64_t orc_compress(int codec_flags, const int32_t *in_, int64_t n, int32_t *out_)
{
    const int codec = codec_flags & ORC_CODEC_MASK;
    if (codec < 0 || codec >= ORC_CODEC_COUNT) return -1;
    uint32_t *out = (uint32_t *)out_;
    uint32_t *tmp = NULL;
    const uint32_t *in = (const uint32_t *)in_;
    if ((codec_flags & ORC_FLAG_DELTA) && n > 0) {
        tmp = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
        memcpy(tmp, in, (size_t)n * sizeof(uint32_t));
        orc_delta(tmp, (size_t)n);
        in = tmp;
    }
    int64_t o = 0;
    switch (codec) {
    case ORC_CODEC_BP32_VB: {
        if (n == 0) break;                       /* Composition.java:39-41 */
        const int64_t m = n - n % 32;            /* BinaryPacking.java:46 */
        if (m > 0) { out[o++] = (uint32_t)m; o += bp32_encode(in, m, out + o, 0, NULL); }
        else out[o++] = 0;                       /* Composition.java:45-48 */
        o += vb_encode(in + m, n - m, out + o, 0, NULL);
        break;
    }
    case ORC_CODEC_FASTPFOR256_VB:
    case ORC_CODEC_FASTPFOR128_VB: {
        if (n == 0) break;
        const int B = codec == ORC_CODEC_FASTPFOR256_VB ? 256 : 128;
        const int64_t m = n - n % B;
        if (m > 0) {
            orc_fastpfor *f = orc_fastpfor_new(B); /* parity contract: a fresh instance per list */
            o += orc_fastpfor_compress(f, (const int32_t *)in, n, (int32_t *)out);
            orc_fastpfor_free(f);
        } else
            out[o++] = 0;
        o += vb_encode(in + m, n - m, out + o, 0, NULL);
        break;
    }
    case ORC_CODEC_IBP32_IVB: {
        if (n == 0) break;
        const int64_t m = n - n % 32;
        uint32_t init = 0;
        if (m > 0) { out[o++] = (uint32_t)m; o += bp32_encode(in, m, out + o, 1, &init); }
        else out[o++] = 0;
        uint32_t zero = 0;                       /* IntegratedVariableByte.java:40: the tail restarts at 0 */
        o += vb_encode(in + m, n - m, out + o, 1, &zero);
        break;
    }
    case ORC_CODEC_SK_IBP32_IVB: {               /* IntegratedIntCompressor.java:38-46 */
        out[o++] = (uint32_t)n;
        if (n == 0) break;                       /* SkippableIntegratedComposition.java:50-51 */
        const int64_t m = n - n % 32;
        uint32_t init = 0;
        o += bp32_encode(in, m, out + o, 1, &init);
        o += vb_encode(in + m, n - m, out + o, 1, &init); /* initvalue is chained (:52,58) */
        break;
    }
    case ORC_CODEC_SK_BP32_VB: {                 /* IntCompressor.java:38-46 */
        out[o++] = (uint32_t)n;
        const int64_t m = n - n % 32;
        o += bp32_encode(in, m, out + o, 0, NULL);
        o += vb_encode(in + m, n - m, out + o, 0, NULL);
        break;
    }
    case ORC_CODEC_SK_FASTPFOR256_VB:
    case ORC_CODEC_SK_FASTPFOR128_VB: {          /* IntCompressor.java:38-46 over SkippableComposition.headlessCompress (:37-45) */
        out[o++] = (uint32_t)n;
        const int B = codec == ORC_CODEC_SK_FASTPFOR256_VB ? 256 : 128;
        const int64_t m = n - n % B;             /* FastPFOR.headlessCompress: greatestMultiple, no count word (FastPFOR.java:92-105) */
        if (m > 0) {
            orc_fastpfor *f = orc_fastpfor_new(B);
            for (int64_t s = 0; s < m;) {
                const int64_t sz = (m - s < FP_PAGE) ? m - s : FP_PAGE;
                o += fp_encode_page(f, in + s, sz, out + o);
                s += sz;
            }
            orc_fastpfor_free(f);
        }
        o += vb_encode(in + m, n - m, out + o, 0, NULL); /* VariableByte.headlessCompress (VariableByte.java:38-77) */
        break;
    }
    case ORC_CODEC_VB:
        o += vb_encode(in, n, out, 0, NULL);
        break;
    case ORC_CODEC_XOR128_IVB: {                 /* IntegratedComposition.java:41-55 over XorBinaryPacking.compress (:31-71) */
        if (n == 0) break;
        const int64_t m = n - n % 128;
        uint32_t ctx = 0;
        if (m > 0) { out[o++] = (uint32_t)m; o += xor128_encode(in, m, out + o, &ctx); }
        else out[o++] = 0;
        uint32_t zero = 0;                       /* IntegratedVariableByte.java:40: the tail's gaps restart at 0 */
        o += vb_encode(in + m, n - m, out + o, 1, &zero);
        break;
    }
    case ORC_CODEC_OPTPFD_VB: {                  /* Composition.java:37-51 over OptPFD.compress (:168-177) */
        if (n == 0) break;
        const int64_t m = n / 128 * 128;
        if (m > 0) { out[o++] = (uint32_t)m; o += optpfd_encode(in, m, out + o); }
        else out[o++] = 0;
        o += vb_encode(in + m, n - m, out + o, 0, NULL);
        break;
    }
    }
    free(tmp);
    return o;
}

int64_t orc_uncompress(int codec_flags, const int32_t *in_, int64_t inlen, int32_t *out_, int64_t out_cap)
{
    const int codec = codec_flags & ORC_CODEC_MASK;
    if (codec < 0 || codec >= ORC_CODEC_COUNT) return -1;
    const uint32_t *in = (const uint32_t *)in_;
    uint32_t *out = (uint32_t *)out_;
    int64_t n = 0;
    switch (codec) {
    case ORC_CODEC_BP32_VB:
    case ORC_CODEC_IBP32_IVB: {
        if (inlen == 0) break;                   /* Composition.java:56-57 */
        const int integ = codec == ORC_CODEC_IBP32_IVB;
        int64_t m = in[0];                       /* BinaryPacking.java:96 */
        m -= m % 32;
        if (m > out_cap) return -1;
        uint32_t init = 0;
        const int64_t p = 1 + bp32_decode(in + 1, out, m, integ, &init);
        const int64_t r = vb_decode_all(in + p, inlen - p, out + m, out_cap - m, integ);
        if (r < 0) return -1;
        n = m + r;
        break;
    }
    case ORC_CODEC_FASTPFOR256_VB:
    case ORC_CODEC_FASTPFOR128_VB: {
        if (inlen == 0) break;
        const int B = codec == ORC_CODEC_FASTPFOR256_VB ? 256 : 128;
        int64_t m = in[0];
        m -= m % B;
        if (m > out_cap) return -1;
        int64_t p = 1;
        if (m > 0) {
            orc_fastpfor *f = orc_fastpfor_new(B);
            orc_fastpfor_uncompress(f, (const int32_t *)in, inlen, (int32_t *)out, &p);
            orc_fastpfor_free(f);
        }
        const int64_t r = vb_decode_all(in + p, inlen - p, out + m, out_cap - m, 0);
        if (r < 0) return -1;
        n = m + r;
        break;
    }
    case ORC_CODEC_SK_IBP32_IVB:
    case ORC_CODEC_SK_BP32_VB: {
        if (inlen == 0) return -1;
        const int integ = codec == ORC_CODEC_SK_IBP32_IVB;
        n = in[0];
        if (n > out_cap) return -1;
        if (inlen - 1 == 0) break;
        const int64_t m = n - n % 32;
        uint32_t init = 0;
        const int64_t p = 1 + bp32_decode(in + 1, out, m, integ, &init);
        vb_decode_n(in + p, out + m, n - m, integ, &init);
        break;
    }
    case ORC_CODEC_SK_FASTPFOR256_VB:
    case ORC_CODEC_SK_FASTPFOR128_VB: {          /* IntCompressor.java:55-61 over SkippableComposition.headlessUncompress (:47-54) */
        if (inlen == 0) return -1;
        const int B = codec == ORC_CODEC_SK_FASTPFOR256_VB ? 256 : 128;
        n = in[0];
        if (n > out_cap) return -1;
        const int64_t m = n - n % B;             /* FastPFOR.headlessUncompress (FastPFOR.java:222-231) */
        int64_t p = 1;
        if (m > 0) {
            orc_fastpfor *f = orc_fastpfor_new(B);
            for (int64_t s = 0; s < m;) {
                const int64_t sz = (m - s < FP_PAGE) ? m - s : FP_PAGE;
                p += fp_decode_page(f, in + p, out + s, sz);
                s += sz;
            }
            orc_fastpfor_free(f);
        }
        uint32_t init = 0;
        vb_decode_n(in + p, out + m, n - m, 0, &init); /* VariableByte.headlessUncompress (VariableByte.java:180-203) */
        break;
    }
    case ORC_CODEC_VB: {
        const int64_t r = vb_decode_all(in, inlen, out, out_cap, 0);
        if (r < 0) return -1;
        n = r;
        break;
    }
    case ORC_CODEC_XOR128_IVB:
    case ORC_CODEC_OPTPFD_VB: {
        if (inlen == 0) break;
        const int xr = codec == ORC_CODEC_XOR128_IVB;
        int64_t m = in[0];                       /* XorBinaryPacking.java:79 / OptPFD.java:184 */
        m -= m % 128;
        if (m > out_cap) return -1;
        const int64_t q = xr ? xor128_decode(in + 1, out, m) : optpfd_decode(in + 1, out, m);
        if (q < 0) return -1;
        const int64_t p = 1 + q;
        const int64_t r = vb_decode_all(in + p, inlen - p, out + m, out_cap - m, xr);
        if (r < 0) return -1;
        n = m + r;
        break;
    }
    }
    if ((codec_flags & ORC_FLAG_DELTA) && n > 0) orc_inverse_delta(out, (size_t)n);
    return n;
}

/* ------------------------------------------------------------- app framing */

static void words_to_be(const uint32_t *w, int64_t n, uint8_t *b) /* BitManipulationHelper.java:59-64 */
{
    for (int64_t i = 0; i < n; ++i) {
        b[4 * i + 0] = (uint8_t)(w[i] >> 24);
        b[4 * i + 1] = (uint8_t)(w[i] >> 16);
        b[4 * i + 2] = (uint8_t)(w[i] >> 8);
        b[4 * i + 3] = (uint8_t)(w[i]);
    }
}

void orc_integers_to_bytes(const int32_t *in, int64_t n, uint8_t *bytes)
{
    words_to_be((const uint32_t *)in, n, bytes);
}

int64_t orc_get_compressed_instruments(int codec, const int32_t *in, int64_t n, uint8_t *bytes)
{
    const int64_t cap = orc_max_compressed_words(n) + 1;
    uint32_t *w = (uint32_t *)calloc((size_t)cap, sizeof(uint32_t)); /* Java arrays are zero-initialised */
    const int64_t o = orc_compress(codec, in, n, (int32_t *)w);
    if (o < 0) { free(w); return -1; }
    words_to_be(w, o + 1, bytes);                /* LoadPortfolios.java:622: outpos.get() + 1 */
    free(w);
    return 4 * (o + 1);
}

int64_t orc_get_decompressed_instruments(int codec, const uint8_t *bytes, int64_t nbytes, int32_t *out, int64_t out_cap)
{
    if (nbytes % 4 != 0) return -1;              /* BitManipulationHelper.java:69-72 */
    const int64_t nw = nbytes / 4;
    uint32_t *w = (uint32_t *)malloc((size_t)(nw + 1) * sizeof(uint32_t));
    for (int64_t i = 0; i < nw; ++i)
        w[i] = ((uint32_t)bytes[4 * i] << 24) | ((uint32_t)bytes[4 * i + 1] << 16) | ((uint32_t)bytes[4 * i + 2] << 8) | bytes[4 * i + 3];
    const int64_t n = orc_uncompress(codec, (const int32_t *)w, nw, out, out_cap);
    free(w);
    return n;
}

/* ------------------------------------------------------------------ batched */

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_batch_compress(int codec, int framing, const int32_t *vals, const int64_t *list_off, int64_t nlists,
                       uint8_t *out, int64_t stride_bytes, int64_t *out_len, int nthreads)
{
    int bad = 0;
    (void)nthreads;
orc_use_threads(nthreads);
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1) reduction(| : bad)
    for (int64_t i = 0; i < nlists; ++i) {
        const int64_t n = list_off[i + 1] - list_off[i];
        uint8_t *dst = out + i * stride_bytes;
        if (framing) {
            if (4 * (orc_max_compressed_words(n) + 1) > stride_bytes) {
                /* compress into a scratch first so a tight stride is still safe */
                uint8_t *tmp = (uint8_t *)malloc((size_t)(4 * (orc_max_compressed_words(n) + 1)));
                const int64_t len = orc_get_compressed_instruments(codec, vals + list_off[i], n, tmp);
                if (len < 0 || len > stride_bytes) bad |= 1; else { memcpy(dst, tmp, (size_t)len); out_len[i] = len; }
                free(tmp);
            } else {
                const int64_t len = orc_get_compressed_instruments(codec, vals + list_off[i], n, dst);
                if (len < 0) bad |= 1; else out_len[i] = len;
            }
        } else {
            uint32_t *tmp = (uint32_t *)malloc((size_t)orc_max_compressed_words(n) * sizeof(uint32_t));
            const int64_t o = orc_compress(codec, vals + list_off[i], n, (int32_t *)tmp);
            if (o < 0 || 4 * o > stride_bytes) bad |= 1; else { memcpy(dst, tmp, (size_t)(4 * o)); out_len[i] = 4 * o; }
            free(tmp);
        }
    }
    return bad ? -1 : 0;
}

int orc_batch_uncompress(int codec, int framing, const uint8_t *in, int64_t stride_bytes, const int64_t *in_len,
                         int64_t nlists, int32_t *out, const int64_t *out_off, int nthreads)
{
    int bad = 0;
    (void)nthreads;
orc_use_threads(nthreads);
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1) reduction(| : bad)
    for (int64_t i = 0; i < nlists; ++i) {
        const int64_t cap = out_off[i + 1] - out_off[i];
        int64_t n;
        if (framing) n = orc_get_decompressed_instruments(codec, in + i * stride_bytes, in_len[i], out + out_off[i], cap);
        else n = orc_uncompress(codec, (const int32_t *)(in + i * stride_bytes), in_len[i] / 4, out + out_off[i], cap);
        if (n != cap) bad |= 1;
    }
    return bad ? -1 : 0;
}
