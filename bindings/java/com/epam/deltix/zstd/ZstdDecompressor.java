/*
 * Drop-in for the reference's com.epam.deltix.zstd.ZstdDecompressor
 * (reference: java/src/main/java/com/epam/deltix/zstd/ZstdDecompressor.java:18-34): the two public methods keep their
 * signatures and error behaviour, the work goes to libzstdb200 through JNI (bindings/java/zstdb200_jni.c), and a
 * batched overload is added — the GPU path pays for itself only on batches.
 *
 * NOT COMPILED in the build image (no JDK there): INTEGRATION.md says how a maintainer builds it.
 */
package com.epam.deltix.zstd;

public class ZstdDecompressor implements AutoCloseable {
    static { System.loadLibrary("zstdb200_jni"); }

    /** One native context per instance: the reference's class holds per-instance tables for the same reason
     *  (ZstdFrameDecompressor.java:137-155) — one instance per thread. */
    private long ctx = create0(256L << 20);

    /** Same contract as the reference: bytes written; RuntimeException on malformed input (Util.java:32-40);
     *  0 when maxOutputLength == 0 (ZstdFrameDecompressor.java:164-166). */
    public int decompress(final byte[] input, final int inputOffset, final int inputLength,
                          final byte[] output, final int outputOffset, final int maxOutputLength) {
        if (maxOutputLength == 0) {
            return 0;
        }
        final int r = decompress0(ctx, input, inputOffset, inputLength, output, outputOffset, maxOutputLength);
        if (isError0(r)) {
            throw new RuntimeException(errorName0(r) + ": offset=" + inputOffset);
        }
        return r;
    }

    /** Batched overload: frame i is input[i][inputOffset[i] .. +inputLength[i]) and goes to
     *  output[i][outputOffset[i] .. +maxOutputLength[i]).  result[i] = bytes written, or the reference's error code
     *  ((int) -(code), ZStdErrors.cs:92-95) — one bad frame does not fail the batch. */
    public void decompress(final byte[][] input, final int[] inputOffset, final int[] inputLength,
                           final byte[][] output, final int[] outputOffset, final int[] maxOutputLength, final int[] result) {
        if (decompressBatch0(ctx, input, inputOffset, inputLength, output, outputOffset, maxOutputLength, result) != 0) {
            throw new RuntimeException("zstdb200: " + lastError0(ctx));
        }
    }

    public static boolean isError(final int result) { return isError0(result); }

    /** ZstdFrameDecompressor.java:922-926: content size from the frame header, -1 when the header does not carry it;
     *  RuntimeException on a bad magic number (:928-940). */
    public static long getDecompressedSize(final byte[] input, final int offset, final int length) {
        return getDecompressedSize0(input, offset, length);
    }

    @Override
    public void close() {
        if (ctx != 0) { destroy0(ctx); ctx = 0; }
    }

    private static native long create0(long maxBatchBytes);
    private static native void destroy0(long ctx);
    private static native String lastError0(long ctx);
    private static native boolean isError0(int code);
    private static native String errorName0(int code);
    private static native int decompress0(long ctx, byte[] in, int inOff, int inLen, byte[] out, int outOff, int maxLen);
    private static native int decompressBatch0(long ctx, byte[][] in, int[] inOff, int[] inLen, byte[][] out, int[] outOff, int[] maxLen, int[] result);
    private static native long getDecompressedSize0(byte[] in, int off, int len);
}
