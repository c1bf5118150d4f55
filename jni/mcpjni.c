/*
 * mcpjni.c -- the JNI half of java/com/mapr/db/data/McpNative.java: one C function per static native, each a thin call
 * into libmcp.so (include/mcp.h).  The reference has no native seam; this is the glue a
 * maintainer adds.
 *
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include mcpjni.c -L.. -lmcp -o libmcpjni.so
 *
 * Error mapping (mcp.h): MCP_E_CAP -> ArrayIndexOutOfBoundsException (what the Java codecs throw when `out` is too
 * small, ExampleTest.java:55-57), MCP_E_FORMAT / MCP_E_ARG -> IllegalArgumentException (BitPacking.java:143-145),
 * everything else -> IllegalStateException with mcp_last_error's text.
 * Arrays are pinned with GetPrimitiveArrayCritical: no JNI call is made between Get and Release.
 */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "mcp.h"

#define CTX(h) ((mcp_ctx *)(intptr_t)(h))

static void throw_rc(JNIEnv *e, mcp_ctx *ctx, int rc, const char *what)
{
    const char *cls = rc == MCP_E_CAP ? "java/lang/ArrayIndexOutOfBoundsException"
                    : (rc == MCP_E_FORMAT || rc == MCP_E_ARG) ? "java/lang/IllegalArgumentException"
                                                              : "java/lang/IllegalStateException";
    char msg[256];
    const char *detail = ctx ? mcp_last_error(ctx) : "";
    strncpy(msg, what, sizeof msg - 1);
    msg[sizeof msg - 1] = 0;
    strncat(msg, ": ", sizeof msg - strlen(msg) - 1);
    strncat(msg, mcp_strerror(rc), sizeof msg - strlen(msg) - 1);
    strncat(msg, " ", sizeof msg - strlen(msg) - 1);
    strncat(msg, detail, sizeof msg - strlen(msg) - 1);
    (*e)->ThrowNew(e, (*e)->FindClass(e, cls), msg);
}

JNIEXPORT jlong JNICALL Java_com_mapr_db_data_McpNative_create(JNIEnv *e, jclass c, jint device)
{
    mcp_ctx *ctx = NULL;
    const int rc = mcp_create(device, &ctx);
    (void)c;
    if (rc != MCP_OK) { throw_rc(e, NULL, rc, "mcp_create (a B200 is required; there is no CPU fallback)"); return 0; }
    return (jlong)(intptr_t)ctx;
}

JNIEXPORT void JNICALL Java_com_mapr_db_data_McpNative_destroy(JNIEnv *e, jclass c, jlong h)
{
    (void)e; (void)c;
    mcp_destroy(CTX(h));
}

typedef int (*codec_fn)(mcp_ctx *, int, const int32_t *, int32_t, int32_t *, int32_t, int32_t *);

static jint codec_call(JNIEnv *e, jlong h, jint codec, jintArray in, jint inpos, jint inlength, jintArray out, jint outpos,
                       codec_fn fn, const char *what)
{
    const jint cap = (*e)->GetArrayLength(e, out) - outpos;
    int32_t n = 0;
    if (inpos < 0 || inlength < 0 || inpos + inlength > (*e)->GetArrayLength(e, in) || cap < 0) {
        throw_rc(e, NULL, MCP_E_CAP, what);
        return 0;
    }
    jint *pi = (jint *)(*e)->GetPrimitiveArrayCritical(e, in, 0);
    jint *po = (jint *)(*e)->GetPrimitiveArrayCritical(e, out, 0);
    const int rc = fn(CTX(h), codec, (const int32_t *)pi + inpos, inlength, (int32_t *)po + outpos, cap, &n);
    (*e)->ReleasePrimitiveArrayCritical(e, out, po, 0);
    (*e)->ReleasePrimitiveArrayCritical(e, in, pi, JNI_ABORT);
    if (rc != MCP_OK) { throw_rc(e, CTX(h), rc, what); return 0; }
    return n;
}

JNIEXPORT jint JNICALL Java_com_mapr_db_data_McpNative_compress(JNIEnv *e, jclass c, jlong h, jint codec, jintArray in, jint inpos,
                                                                jint inlength, jintArray out, jint outpos)
{
    (void)c;
    return codec_call(e, h, codec, in, inpos, inlength, out, outpos, mcp_codec_compress, "mcp_codec_compress");
}

JNIEXPORT jint JNICALL Java_com_mapr_db_data_McpNative_uncompress(JNIEnv *e, jclass c, jlong h, jint codec, jintArray in, jint inpos,
                                                                  jint inlength, jintArray out, jint outpos)
{
    (void)c;
    return codec_call(e, h, codec, in, inpos, inlength, out, outpos, mcp_codec_uncompress, "mcp_codec_uncompress");
}

JNIEXPORT jint JNICALL Java_com_mapr_db_data_McpNative_headlessCompress(JNIEnv *e, jclass c, jlong h, jint codec, jintArray in,
                                                                        jint inpos, jint inlength, jintArray out, jint outpos)
{
    (void)c;
    return codec_call(e, h, codec, in, inpos, inlength, out, outpos, mcp_codec_headless_compress, "mcp_codec_headless_compress");
}

JNIEXPORT jint JNICALL Java_com_mapr_db_data_McpNative_headlessUncompress(JNIEnv *e, jclass c, jlong h, jint codec, jintArray in,
                                                                          jint inpos, jint inlength, jintArray out, jint outpos,
                                                                          jint num)
{
    int32_t used = 0;
    (void)c;
    if (inpos < 0 || inlength < 0 || inpos + inlength > (*e)->GetArrayLength(e, in) || num < 0 ||
        outpos < 0 || outpos + num > (*e)->GetArrayLength(e, out)) {
        throw_rc(e, NULL, MCP_E_CAP, "mcp_codec_headless_uncompress");
        return 0;
    }
    jint *pi = (jint *)(*e)->GetPrimitiveArrayCritical(e, in, 0);
    jint *po = (jint *)(*e)->GetPrimitiveArrayCritical(e, out, 0);
    const int rc = mcp_codec_headless_uncompress(CTX(h), codec, (const int32_t *)pi + inpos, inlength, (int32_t *)po + outpos, num, &used);
    (*e)->ReleasePrimitiveArrayCritical(e, out, po, 0);
    (*e)->ReleasePrimitiveArrayCritical(e, in, pi, JNI_ABORT);
    if (rc != MCP_OK) { throw_rc(e, CTX(h), rc, "mcp_codec_headless_uncompress"); return 0; }
    return used;
}

JNIEXPORT jbyteArray JNICALL Java_com_mapr_db_data_McpNative_getCompressedInstruments(JNIEnv *e, jclass c, jlong h, jint codec,
                                                                                       jintArray instruments)
{
    const jint n = (*e)->GetArrayLength(e, instruments);
    const int64_t cap = mcp_codec_max_bytes(codec, MCP_FRAME_APP, n);
    int64_t nbytes = 0;
    uint8_t *buf;
    jbyteArray out;
    (void)c;
    if (cap < 0 || !(buf = (uint8_t *)malloc((size_t)cap + 1))) { throw_rc(e, NULL, cap < 0 ? MCP_E_ARG : MCP_E_NOMEM, "getCompressedInstruments"); return NULL; }
    jint *pi = (jint *)(*e)->GetPrimitiveArrayCritical(e, instruments, 0);
    const int rc = mcp_get_compressed_instruments(CTX(h), codec, (const int32_t *)pi, n, buf, cap, &nbytes);
    (*e)->ReleasePrimitiveArrayCritical(e, instruments, pi, JNI_ABORT);
    if (rc != MCP_OK) { free(buf); throw_rc(e, CTX(h), rc, "mcp_get_compressed_instruments"); return NULL; }
    out = (*e)->NewByteArray(e, (jsize)nbytes);
    if (out) (*e)->SetByteArrayRegion(e, out, 0, (jsize)nbytes, (const jbyte *)buf);
    free(buf);
    return out;
}

JNIEXPORT jintArray JNICALL Java_com_mapr_db_data_McpNative_getDecompressedInstruments(JNIEnv *e, jclass c, jlong h, jint codec,
                                                                                        jbyteArray blob, jint n)
{
    const jint nbytes = (*e)->GetArrayLength(e, blob);
    jintArray out = (*e)->NewIntArray(e, n < 0 ? 0 : n);
    (void)c;
    if (!out || n < 0) { if (n < 0) throw_rc(e, NULL, MCP_E_ARG, "getDecompressedInstruments"); return NULL; }
    jbyte *pb = (jbyte *)(*e)->GetPrimitiveArrayCritical(e, blob, 0);
    jint *po = (jint *)(*e)->GetPrimitiveArrayCritical(e, out, 0);
    const int rc = mcp_get_decompressed_instruments(CTX(h), codec, (const uint8_t *)pb, nbytes, n, (int32_t *)po);
    (*e)->ReleasePrimitiveArrayCritical(e, out, po, 0);
    (*e)->ReleasePrimitiveArrayCritical(e, blob, pb, JNI_ABORT);
    if (rc != MCP_OK) { throw_rc(e, CTX(h), rc, "mcp_get_decompressed_instruments"); return NULL; }
    return out;
}

JNIEXPORT jbyteArray JNICALL Java_com_mapr_db_data_McpNative_encodeBatch(JNIEnv *e, jclass c, jlong h, jint codec, jint framing,
                                                                          jintArray vals, jlongArray listOff, jlongArray outOff)
{
    const jsize nlists = (*e)->GetArrayLength(e, listOff) - 1;
    int64_t total = 0, cap;
    uint8_t *buf;
    jbyteArray out = NULL;
    int rc;
    (void)c;
    if (nlists < 0 || (*e)->GetArrayLength(e, outOff) < nlists + 1) { throw_rc(e, NULL, MCP_E_ARG, "encodeBatch"); return NULL; }
    /* generous first guess (the reference's own convention is 4n + 1024 ints); MCP_E_CAP reports the exact need */
    cap = 5 * (int64_t)(*e)->GetArrayLength(e, vals) + 16 * (int64_t)nlists + 64;
    for (;;) {
        if (!(buf = (uint8_t *)malloc((size_t)cap))) { throw_rc(e, NULL, MCP_E_NOMEM, "encodeBatch"); return NULL; }
        jint *pv = (jint *)(*e)->GetPrimitiveArrayCritical(e, vals, 0);
        jlong *pl = (jlong *)(*e)->GetPrimitiveArrayCritical(e, listOff, 0);
        jlong *po = (jlong *)(*e)->GetPrimitiveArrayCritical(e, outOff, 0);
        rc = mcp_codec_encode(CTX(h), codec, framing, (const int32_t *)pv, (const int64_t *)pl, nlists, (int64_t *)po, buf, cap, &total);
        (*e)->ReleasePrimitiveArrayCritical(e, outOff, po, 0);
        (*e)->ReleasePrimitiveArrayCritical(e, listOff, pl, JNI_ABORT);
        (*e)->ReleasePrimitiveArrayCritical(e, vals, pv, JNI_ABORT);
        if (rc == MCP_E_CAP && total > cap) { free(buf); cap = total; continue; }
        break;
    }
    if (rc != MCP_OK) { free(buf); throw_rc(e, CTX(h), rc, "mcp_codec_encode"); return NULL; }
    out = (*e)->NewByteArray(e, (jsize)total);
    if (out) (*e)->SetByteArrayRegion(e, out, 0, (jsize)total, (const jbyte *)buf);
    free(buf);
    return out;
}

JNIEXPORT jbyteArray JNICALL Java_com_mapr_db_data_McpNative_simulate(JNIEnv *e, jclass c, jlong h, jdoubleArray E, jlong P,
                                                                       jdoubleArray S, jlong N, jint F, jdouble alpha, jint codec,
                                                                       jdoubleArray stats, jintArray varTrial, jlongArray encOff)
{
    const int64_t t = mcp_tail_size(N, alpha);
    const int64_t cap = P * mcp_codec_max_bytes(codec, MCP_FRAME_APP, t);
    int64_t enc_len = 0;
    uint8_t *buf;
    jbyteArray out = NULL;
    (void)c;
    if (P < 0 || N < 0 || F <= 0 || (*e)->GetArrayLength(e, E) < P * F || (*e)->GetArrayLength(e, S) < N * F ||
        (*e)->GetArrayLength(e, stats) < 4 * P || (*e)->GetArrayLength(e, varTrial) < P || (*e)->GetArrayLength(e, encOff) < P + 1) {
        throw_rc(e, NULL, MCP_E_ARG, "simulate");
        return NULL;
    }
    if (!(buf = (uint8_t *)malloc((size_t)(cap > 0 ? cap : 1)))) { throw_rc(e, NULL, MCP_E_NOMEM, "simulate"); return NULL; }
    jdouble *pe = (jdouble *)(*e)->GetPrimitiveArrayCritical(e, E, 0);
    jdouble *ps = (jdouble *)(*e)->GetPrimitiveArrayCritical(e, S, 0);
    jdouble *pt = (jdouble *)(*e)->GetPrimitiveArrayCritical(e, stats, 0);
    jint *pv = (jint *)(*e)->GetPrimitiveArrayCritical(e, varTrial, 0);
    jlong *po = (jlong *)(*e)->GetPrimitiveArrayCritical(e, encOff, 0);
    const int rc = mcp_simulate(CTX(h), pe, P, ps, N, F, alpha, codec, MCP_FRAME_APP, pt, pt + P, pt + 2 * P, (int32_t *)pv, pt + 3 * P,
                                NULL, (int64_t *)po, buf, cap, &enc_len);
    (*e)->ReleasePrimitiveArrayCritical(e, encOff, po, 0);
    (*e)->ReleasePrimitiveArrayCritical(e, varTrial, pv, 0);
    (*e)->ReleasePrimitiveArrayCritical(e, stats, pt, 0);
    (*e)->ReleasePrimitiveArrayCritical(e, S, ps, JNI_ABORT);
    (*e)->ReleasePrimitiveArrayCritical(e, E, pe, JNI_ABORT);
    if (rc != MCP_OK) { free(buf); throw_rc(e, CTX(h), rc, "mcp_simulate"); return NULL; }
    out = (*e)->NewByteArray(e, (jsize)enc_len);
    if (out) (*e)->SetByteArrayRegion(e, out, 0, (jsize)enc_len, (const jbyte *)buf);
    free(buf);
    return out;
}

JNIEXPORT jlong JNICALL Java_com_mapr_db_data_McpNative_tailSize(JNIEnv *e, jclass c, jlong N, jdouble alpha)
{
    (void)e; (void)c;
    return mcp_tail_size(N, alpha);
}
