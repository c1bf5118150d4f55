package com.mapr.db.data;

import me.lemire.integercompression.IntWrapper;
import me.lemire.integercompression.SkippableIntegerCODEC;

/**
 * SkippableIntegerCODEC (.../SkippableIntegerCODEC.java:43-66) over libmcp.so: the stream WITHOUT any count word, the
 * caller supplies the number of values on the way back.  Drop-in for the composition behind IntCompressor:
 *
 * <pre>
 * new IntCompressor(new NativeSkippableCodec(NativeSkippableCodec.SK_FASTPFOR256_VB))
 *     // for: new IntCompressor(new SkippableComposition(new FastPFOR(), new VariableByte()))
 * </pre>
 */
public final class NativeSkippableCodec implements SkippableIntegerCODEC, AutoCloseable {
    public static final int SK_BP32_VB = McpNative.SK_BP32_VB, SK_FASTPFOR256_VB = McpNative.SK_FASTPFOR256_VB,
            SK_FASTPFOR128_VB = McpNative.SK_FASTPFOR128_VB, SK_IBP32_IVB = McpNative.SK_IBP32_IVB;

    private final int codecId;
    private long ctx;

    public NativeSkippableCodec(int codecId) {
        this.codecId = codecId;
        this.ctx = McpNative.create(0);
    }

    @Override
    public void headlessCompress(int[] in, IntWrapper inpos, int inlength, int[] out, IntWrapper outpos) {
        int written = McpNative.headlessCompress(ctx, codecId, in, inpos.get(), inlength, out, outpos.get());
        inpos.add(inlength);
        outpos.add(written);
    }

    @Override
    public void headlessUncompress(int[] in, IntWrapper inpos, int inlength, int[] out, IntWrapper outpos, int num) {
        int consumed = McpNative.headlessUncompress(ctx, codecId, in, inpos.get(), inlength, out, outpos.get(), num);
        inpos.add(consumed);
        outpos.add(num);
    }

    @Override
    public void close() {
        if (ctx != 0) {
            McpNative.destroy(ctx);
            ctx = 0;
        }
    }
}
