package com.mapr.db.data;

import me.lemire.integercompression.IntWrapper;
import me.lemire.integercompression.IntegerCODEC;

/**
 * IntegerCODEC (JavaFastPFOR/src/main/java/me/lemire/integercompression/IntegerCODEC.java:36-58) over libmcp.so.
 *
 * Drop-in for the codec object the loader holds (LoadPortfolios.java:725-726):
 *
 * <pre>
 * private IntegerCODEC integerCodec = new NativeIntegerCodec(NativeIntegerCodec.BP32_VB);
 * </pre>
 *
 * produces, int for int, what {@code new Composition(new BinaryPacking(), new VariableByte())} produces.  Position
 * wrappers advance exactly as the Java codecs advance them (inpos += inlength, outpos += ints written); an output
 * array that is too small raises ArrayIndexOutOfBoundsException, as the Java codecs do (ExampleTest.java:55-57).
 * Not thread-safe, like FastPFOR itself (FastPFOR.java:36-37): one instance per thread.
 *
 * A call is one GPU round trip: right for parity and for occasional lists, wrong for a bulk load -- there collect
 * the lists and call {@link #compressAll(int[], long[], long[])} once (mcp_codec_encode, CSR in / CSR out).
 */
public final class NativeIntegerCodec implements IntegerCODEC, AutoCloseable {
    public static final int BP32_VB = McpNative.BP32_VB, FASTPFOR256_VB = McpNative.FASTPFOR256_VB,
            FASTPFOR128_VB = McpNative.FASTPFOR128_VB, IBP32_IVB = McpNative.IBP32_IVB, VB = McpNative.VB,
            FLAG_DELTA = McpNative.FLAG_DELTA;

    private final int codecId;
    private long ctx;

    public NativeIntegerCodec(int codecId) {
        this(codecId, 0);
    }

    public NativeIntegerCodec(int codecId, int device) {
        this.codecId = codecId;
        this.ctx = McpNative.create(device);
    }

    @Override
    public void compress(int[] in, IntWrapper inpos, int inlength, int[] out, IntWrapper outpos) {
        int written = McpNative.compress(ctx, codecId, in, inpos.get(), inlength, out, outpos.get());
        inpos.add(inlength);
        outpos.add(written);
    }

    @Override
    public void uncompress(int[] in, IntWrapper inpos, int inlength, int[] out, IntWrapper outpos) {
        int produced = McpNative.uncompress(ctx, codecId, in, inpos.get(), inlength, out, outpos.get());
        inpos.add(inlength);
        outpos.add(produced);
    }

    /** LoadPortfolios.getCompressedInstruments (LoadPortfolios.java:608-623) in one native call. */
    public byte[] getCompressedInstruments(int[] instruments) {
        return McpNative.getCompressedInstruments(ctx, codecId, instruments);
    }

    /** LoadPortfolios.getDecompressedInstruments (LoadPortfolios.java:625-662). */
    public int[] getDecompressedInstruments(byte[] compressedInstruments, int numberOfIntegers) {
        return McpNative.getDecompressedInstruments(ctx, codecId, compressedInstruments, numberOfIntegers);
    }

    /** Many lists in one call: list i = vals[listOff[i] .. listOff[i+1]); blob i = result[outOff[i] .. outOff[i+1]). */
    public byte[] compressAll(int[] vals, long[] listOff, long[] outOff) {
        return McpNative.encodeBatch(ctx, codecId, McpNative.FRAME_APP, vals, listOff, outOff);
    }

    @Override
    public String toString() {
        return "NativeIntegerCodec(" + codecId + ")";
    }

    @Override
    public void close() {
        if (ctx != 0) {
            McpNative.destroy(ctx);
            ctx = 0;
        }
    }
}
