package com.mapr.db.data;

/**
 * JNI surface of libmcp.so (include/mcp.h), one static native per C entry point the Java side needs.
 * Implemented by jni/mcpjni.c (libmcpjni.so, which links libmcp.so).  Handles are the mcp_ctx pointer as a long.
 *
 * The reference has no native seam at all (SURVEY.md section 0): this class, NativeIntegerCodec,
 * NativeSkippableCodec and SimulatePortfolios are what a maintainer adds; the reference's own classes stay as they are.
 * There is no CPU fallback: create() throws IllegalStateException without a usable GPU.
 */
final class McpNative {
    static {
        System.loadLibrary("mcpjni");
    }

    /* codec ids of include/mcp.h (MCP_CODEC_*), optionally OR-ed with FLAG_DELTA */
    static final int BP32_VB = 0, FASTPFOR256_VB = 1, FASTPFOR128_VB = 2, IBP32_IVB = 3, SK_IBP32_IVB = 4, VB = 5,
            SK_BP32_VB = 6, SK_FASTPFOR256_VB = 7, SK_FASTPFOR128_VB = 8, FLAG_DELTA = 0x100;
    static final int FRAME_WORDS = 0, FRAME_APP = 1;

    private McpNative() {
    }

    static native long create(int device);

    static native void destroy(long ctx);

    /** IntegerCODEC.compress on in[inpos, inpos+inlength) into out[outpos..): returns the ints written. */
    static native int compress(long ctx, int codecId, int[] in, int inpos, int inlength, int[] out, int outpos);

    /** IntegerCODEC.uncompress: returns the ints produced. */
    static native int uncompress(long ctx, int codecId, int[] in, int inpos, int inlength, int[] out, int outpos);

    /** SkippableIntegerCODEC.headlessCompress: returns the ints written. */
    static native int headlessCompress(long ctx, int codecId, int[] in, int inpos, int inlength, int[] out, int outpos);

    /** SkippableIntegerCODEC.headlessUncompress of exactly num values: returns the ints consumed. */
    static native int headlessUncompress(long ctx, int codecId, int[] in, int inpos, int inlength, int[] out, int outpos,
            int num);

    /** LoadPortfolios.getCompressedInstruments: the big-endian blob with its trailing zero int. */
    static native byte[] getCompressedInstruments(long ctx, int codecId, int[] instruments);

    /** LoadPortfolios.getDecompressedInstruments. */
    static native int[] getDecompressedInstruments(long ctx, int codecId, byte[] blob, int numberOfIntegers);

    /**
     * Batched encode (mcp_codec_encode): list i is vals[listOff[i] .. listOff[i+1]).  Returns the encoded bytes;
     * outOff[0..nlists] receives the byte offset of every list's blob.
     */
    static native byte[] encodeBatch(long ctx, int codecId, int framing, int[] vals, long[] listOff, long[] outOff);

    /**
     * mcp_simulate: E is P x F row-major, S is N x F row-major.  stats = double[4 * P] (mean, variance, VaR loss, CVaR,
     * each P long), varTrial = int[P], encOff = long[P + 1]; returns the encoded breach lists (FRAME_APP blobs).
     */
    static native byte[] simulate(long ctx, double[] E, long P, double[] S, long N, int F, double alpha, int codecId,
            double[] stats, int[] varTrial, long[] encOff);

    static native long tailSize(long N, double alpha);
}
