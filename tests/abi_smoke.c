/*
 * abi_smoke.c -- the C ABI exercised by a C compiler, not only through ctypes: links libmcp.so, runs the
 * stored-document known-answer vector of the reference (setup.txt:47-48: the 17 instruments of portfolios.json:1 and the
 * 36 bytes the loader stored for them) through mcp_get_compressed_instruments / mcp_get_decompressed_instruments and
 * IntegerCODEC-style mcp_codec_compress / mcp_codec_uncompress, and a small mcp_simulate.
 *
 *   gcc -std=c99 -Iinclude tests/abi_smoke.c -L monte-carlo-portfolios_b200 -lmcp -lm -Wl,-rpath,$PWD/monte-carlo-portfolios_b200 -o abi_smoke
 *
 * Exit code 0 = everything matched; 77 = no GPU (mcp_create said MCP_E_CUDA: there is no CPU fallback to test).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mcp.h"

static const int32_t kat_ints[17] = {719, 5, 3990, 4112, 3328, 78, 1163, 135, 263, 67, 67, 1860, 75, 1342, 1145, 38, 1216};
static const uint8_t kat_bytes[36] = {0x00, 0x00, 0x00, 0x00, 0x16, 0x85, 0x85, 0x4f, 0x00, 0xa0, 0x10, 0x9f, 0x89, 0x0b, 0xce, 0x9a, 0x82, 0x07,
                                      0x81, 0x07, 0x8e, 0x44, 0xc3, 0xc3, 0x79, 0x8a, 0x3e, 0xcb, 0x89, 0x40, 0xa6, 0x88, 0x00, 0x00, 0x00, 0x00};

#define CHECK(cond, what)                                                        \
    do {                                                                         \
        if (!(cond)) {                                                           \
            fprintf(stderr, "abi_smoke: FAILED %s (line %d): %s\n", what, __LINE__, ctx ? mcp_last_error(ctx) : ""); \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(void)
{
    mcp_ctx *ctx = NULL;
    int rc = mcp_create(0, &ctx);
    if (rc == MCP_E_CUDA) { fprintf(stderr, "abi_smoke: no GPU (%s)\n", mcp_strerror(rc)); return 77; }
    CHECK(rc == MCP_OK && ctx, "mcp_create");

    /* LoadPortfolios.getCompressedInstruments / getDecompressedInstruments */
    uint8_t blob[256];
    int64_t nbytes = 0;
    rc = mcp_get_compressed_instruments(ctx, MCP_CODEC_BP32_VB, kat_ints, 17, blob, sizeof blob, &nbytes);
    CHECK(rc == MCP_OK && nbytes == 36 && memcmp(blob, kat_bytes, 36) == 0, "stored-document known answer (encode)");
    int32_t back[17];
    rc = mcp_get_decompressed_instruments(ctx, MCP_CODEC_BP32_VB, kat_bytes, 36, 17, back);
    CHECK(rc == MCP_OK && memcmp(back, kat_ints, sizeof back) == 0, "stored-document known answer (decode)");

    /* IntegerCODEC.compress / uncompress: native words; the app framing is their big-endian image plus a zero int */
    int32_t words[64], nwords = 0, ncount = 0;
    rc = mcp_codec_compress(ctx, MCP_CODEC_BP32_VB, kat_ints, 17, words, 64, &nwords);
    CHECK(rc == MCP_OK && nwords == 8, "mcp_codec_compress");
    for (int i = 0; i < 8; ++i) {
        const uint32_t be = ((uint32_t)kat_bytes[4 * i] << 24) | ((uint32_t)kat_bytes[4 * i + 1] << 16) | ((uint32_t)kat_bytes[4 * i + 2] << 8) | kat_bytes[4 * i + 3];
        CHECK((uint32_t)words[i] == be, "words = big-endian image of the stored bytes");
    }
    rc = mcp_codec_uncompress(ctx, MCP_CODEC_BP32_VB, words, nwords, back, 17, &ncount);
    CHECK(rc == MCP_OK && ncount == 17 && memcmp(back, kat_ints, sizeof back) == 0, "mcp_codec_uncompress");
    rc = mcp_codec_compress(ctx, MCP_CODEC_BP32_VB, kat_ints, 17, words, 3, &nwords);
    CHECK(rc == MCP_E_CAP, "too small an output array is MCP_E_CAP (Java: ArrayIndexOutOfBoundsException)");

    /* the whole path on a tiny problem: 3 portfolios x 64 trials x 16 factors */
    enum { P = 3, N = 64, F = 16 };
    static double E[P * F], S[N * F], mean[P], var[P], vl[P], cv[P];
    int32_t vt[P];
    int64_t off[P + 1], enc_len = 0;
    static uint8_t enc[4096];
    for (int i = 0; i < P * F; ++i) E[i] = (double)((i * 37) % 11) - 5.0;
    for (int i = 0; i < N * F; ++i) S[i] = (double)((i * 53) % 17) * 0.125 - 1.0;
    rc = mcp_simulate(ctx, E, P, S, N, F, 0.9, MCP_CODEC_FASTPFOR256_VB | MCP_FLAG_DELTA, MCP_FRAME_APP, mean, var, vl, vt, cv, NULL, off,
                      enc, sizeof enc, &enc_len);
    CHECK(rc == MCP_OK && off[0] == 0 && off[P] == enc_len && enc_len > 0, "mcp_simulate");
    const int64_t t = mcp_tail_size(N, 0.9);
    for (int p = 0; p < P; ++p) {
        /* the VaR trial's loss is the reported VaR loss: V = fma chain, loss = 0 - V */
        double acc = 0.0;
        for (int k = 0; k < F; ++k) acc = __builtin_fma(E[p * F + k], S[vt[p] * F + k], acc);
        CHECK(0.0 - acc == vl[p], "VaR loss = loss of the VaR trial");
        int32_t lst[64];
        rc = mcp_get_decompressed_instruments(ctx, MCP_CODEC_FASTPFOR256_VB | MCP_FLAG_DELTA, enc + off[p], off[p + 1] - off[p], (int32_t)t, lst);
        CHECK(rc == MCP_OK, "breach list decodes");
        int seen = 0;
        for (int i = 0; i < t; ++i) { CHECK(i == 0 || lst[i] > lst[i - 1], "breach list ascending"); seen |= lst[i] == vt[p]; }
        CHECK(seen, "the VaR trial is in the breach list");
    }
    mcp_destroy(ctx);
    printf("abi_smoke ok: known answer, IntegerCODEC round trip, MCP_E_CAP, mcp_simulate\n");
    return 0;
}
