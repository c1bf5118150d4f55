/*
 * weights_oracle.c -- CPU oracle of the loader's WEIGHTS column.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * The reference stores a portfolio's weights as Snappy.compress(double[]) -- the raw Snappy format over the
 * little-endian bytes of the doubles (LoadPortfolios.java:563-595, call sites :568,:580) -- if that is smaller than
 * 8 n bytes, else as big-endian doubles with algo "none" (:270-278, BitManipulationHelper.java:193-208).
 *
 * PARITY UNPINNED for the compressed bytes: snappy-java 1.1.10.4 is an un-vendored dependency
 * (maprdb-portfolio/pom.xml:50-56) and no reference test touches it.  What is pinned is the FORMAT (the published
 * Snappy format description: a varint length, then literal / copy elements): the decoder below accepts every valid
 * stream, so blobs written by snappy-java decode here and on the GPU; the encoder is this repository's own greedy
 * matcher (builder-defined, stated below) and writes valid streams that any Snappy decoder reads, but not the byte
 * choices snappy-java's matcher would make.
 *
 * Encoder, stated once for the oracle and the CUDA kernel:
 *   table of 1024 positions, h(x) = (x * 0x1e35a7bd) >> 22 over the 4 bytes at the cursor; the cursor advances one byte
 *   at a time; a candidate within 65535 bytes whose 4 bytes match starts a copy that is extended as far as the input
 *   matches; pending literal bytes are flushed first; copies are split like snappy's EmitCopy (64-byte pieces, a
 *   60-byte piece when 65..67 remain, the 1-byte-offset form for 4..11 bytes at offsets < 2048); positions inside a
 *   copy are not entered into the table.
 */
#include "oracle.h"

#include <string.h>

static uint32_t ld32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

static int64_t emit_literal(const uint8_t *src, int64_t len, uint8_t *out)
{
    int64_t o = 0;
    if (len <= 0) return 0;
    const uint32_t n = (uint32_t)(len - 1);
    if (n < 60) { if (out) out[o] = (uint8_t)(n << 2); ++o; }
    else {
        const int k = n < (1u << 8) ? 1 : n < (1u << 16) ? 2 : n < (1u << 24) ? 3 : 4;
        if (out) out[o] = (uint8_t)((59 + k) << 2);
        ++o;
        for (int i = 0; i < k; ++i) { if (out) out[o] = (uint8_t)(n >> (8 * i)); ++o; }
    }
    if (out) memcpy(out + o, src, (size_t)len);
    return o + len;
}

static int64_t emit_copy_upto64(int64_t offset, int64_t len, uint8_t *out)
{
    if (len < 12 && offset < 2048 && len >= 4) {
        if (out) { out[0] = (uint8_t)(1 | ((len - 4) << 2) | ((offset >> 8) << 5)); out[1] = (uint8_t)(offset & 0xff); }
        return 2;
    }
    if (out) { out[0] = (uint8_t)(2 | ((len - 1) << 2)); out[1] = (uint8_t)(offset & 0xff); out[2] = (uint8_t)(offset >> 8); }
    return 3;
}

static int64_t emit_copy(int64_t offset, int64_t len, uint8_t *out)
{
    int64_t o = 0;
    while (len >= 68) { o += emit_copy_upto64(offset, 64, out ? out + o : NULL); len -= 64; }
    if (len > 64) { o += emit_copy_upto64(offset, 60, out ? out + o : NULL); len -= 60; }
    return o + emit_copy_upto64(offset, len, out ? out + o : NULL);
}

/* raw Snappy stream of `len` bytes; out may be NULL to size.  Upper bound: 32 + len + len / 6 (snappy's MaxCompressedLength). */
int64_t orc_snappy_compress(const uint8_t *in, int64_t len, uint8_t *out)
{
    int64_t o = 0;
    uint64_t v = (uint64_t)len;
    do { uint8_t b = (uint8_t)(v & 0x7f); v >>= 7; if (v) b |= 0x80; if (out) out[o] = b; ++o; } while (v);
    int32_t table[1024];
    for (int i = 0; i < 1024; ++i) table[i] = -1;
    int64_t pos = 0, lit = 0;
    while (pos + 4 <= len) {
        const uint32_t x = ld32(in + pos);
        const uint32_t h = (x * 0x1e35a7bdu) >> 22;
        const int64_t cand = table[h];
        table[h] = (int32_t)pos;
        if (cand >= 0 && pos - cand <= 65535 && ld32(in + cand) == x) {
            int64_t m = 4;
            while (pos + m < len && in[cand + m] == in[pos + m]) ++m;
            o += emit_literal(in + lit, pos - lit, out ? out + o : NULL);
            o += emit_copy(pos - cand, m, out ? out + o : NULL);
            pos += m;
            lit = pos;
        } else
            ++pos;
    }
    o += emit_literal(in + lit, len - lit, out ? out + o : NULL);
    return o;
}

/* any valid raw Snappy stream (all four element kinds); returns the bytes produced or -1 if malformed / out_cap too small */
int64_t orc_snappy_uncompress(const uint8_t *in, int64_t inlen, uint8_t *out, int64_t out_cap)
{
    int64_t p = 0, o = 0;
    uint64_t total = 0;
    int shift = 0;
    for (;;) {
        if (p >= inlen || shift > 35) return -1;
        const uint8_t b = in[p++];
        total |= (uint64_t)(b & 0x7f) << shift;
        shift += 7;
        if (!(b & 0x80)) break;
    }
    if ((int64_t)total > out_cap) return -1;
    while (p < inlen) {
        const uint8_t tag = in[p++];
        int64_t len, offset = 0;
        if ((tag & 3) == 0) {
            len = (tag >> 2) + 1;
            if (len > 60) {
                const int k = (int)len - 60;
                if (p + k > inlen) return -1;
                uint32_t n = 0;
                for (int i = 0; i < k; ++i) n |= (uint32_t)in[p + i] << (8 * i);
                p += k;
                len = (int64_t)n + 1;
            }
            if (p + len > inlen || o + len > (int64_t)total) return -1;
            memcpy(out + o, in + p, (size_t)len);
            p += len;
            o += len;
            continue;
        }
        if ((tag & 3) == 1) {
            if (p + 1 > inlen) return -1;
            len = ((tag >> 2) & 7) + 4;
            offset = ((int64_t)(tag >> 5) << 8) | in[p];
            p += 1;
        } else if ((tag & 3) == 2) {
            if (p + 2 > inlen) return -1;
            len = (tag >> 2) + 1;
            offset = in[p] | ((int64_t)in[p + 1] << 8);
            p += 2;
        } else {
            if (p + 4 > inlen) return -1;
            len = (tag >> 2) + 1;
            offset = (int64_t)ld32(in + p);
            p += 4;
        }
        if (offset <= 0 || offset > o || o + len > (int64_t)total) return -1;
        for (int64_t i = 0; i < len; ++i) out[o + i] = out[o - offset + i]; /* overlapping copies repeat the pattern */
        o += len;
    }
    return o == (int64_t)total ? o : -1;
}

/* LoadPortfolios.getCompressedWeights (:563-595): Snappy.compress(double[]) = the raw stream of the doubles' native
 * (little-endian) bytes.  Returns bytes written (out may be NULL to size). */
int64_t orc_get_compressed_weights(const double *w, int64_t n, uint8_t *out)
{
    return orc_snappy_compress((const uint8_t *)w, 8 * n, out);
}

/* Snappy.uncompressDoubleArray (:480,:580): returns the number of doubles, -1 if malformed */
int64_t orc_get_decompressed_weights(const uint8_t *in, int64_t inlen, double *out, int64_t out_cap)
{
    const int64_t r = orc_snappy_uncompress(in, inlen, (uint8_t *)out, 8 * out_cap);
    return r < 0 || (r & 7) ? -1 : r / 8;
}

/* BitManipulationHelper.doublesToBytes (:193-208): big-endian IEEE bits -- what the loader stores under algo "none" */
void orc_doubles_to_bytes(const double *w, int64_t n, uint8_t *bytes)
{
    for (int64_t i = 0; i < n; ++i) {
        uint64_t bits;
        memcpy(&bits, w + i, 8);
        for (int k = 0; k < 8; ++k) bytes[8 * i + k] = (uint8_t)(bits >> (56 - 8 * k));
    }
}
