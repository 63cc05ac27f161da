/* zstdb200_jni.c — JNI glue between com.epam.deltix.zstd.ZstdDecompressor (bindings/java/...) and libzstdb200
 * (include/zstdb200.h).  It replaces the reference's pure-Java path behind
 * ZstdDecompressor.decompress(byte[],int,int,byte[],int,int) and getDecompressedSize (ZstdDecompressor.java:22-33).
 *
 * Build (on a box with a JDK):  cc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *                               bindings/java/zstdb200_jni.c -o libzstdb200_jni.so -Lzstandard_b200 -lzstdb200
 * The build image has no JDK; tests/test_bindings.py compiles this file against a minimal stand-in jni.h
 * (tests/jni_stub/jni.h) and links it against libzstdb200.so, which checks the C and the ABI use, nothing more.
 *
 * Arrays are pinned with Get/ReleasePrimitiveArrayCritical for the duration of the call — the same contract as the
 * C# shim's `fixed`: pinned for the VM, pageable for CUDA, so the library stages them. */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "zstdb200.h"

#define FN(name) Java_com_epam_deltix_zstd_ZstdDecompressor_##name

static const char* code_name(uint32_t r) {
  switch (0u - r) {                                   /* ZStdErrors.cs:61-90 */
    case 1: return "GENERIC"; case 10: return "prefix_unknown"; case 14: return "frameParameter_unsupported";
    case 16: return "frameParameter_windowTooLarge"; case 20: return "corruption_detected"; case 22: return "checksum_wrong";
    case 30: return "dictionary_corrupted"; case 32: return "dictionary_wrong"; case 44: return "tableLog_tooLarge";
    case 46: return "maxSymbolValue_tooLarge"; case 48: return "maxSymbolValue_tooSmall"; case 70: return "dstSize_tooSmall";
    case 72: return "srcSize_wrong"; default: return "error";
  }
}

JNIEXPORT jlong JNICALL FN(create0)(JNIEnv* env, jclass cls, jlong maxBatchBytes) {
  zstdb200_ctx* ctx = NULL;
  (void)env; (void)cls;
  if (zstdb200_create(&ctx, NULL, 0, (size_t)maxBatchBytes) != 0) return 0;
  return (jlong)(intptr_t)ctx;
}
JNIEXPORT void JNICALL FN(destroy0)(JNIEnv* env, jclass cls, jlong ctx) { (void)env; (void)cls; zstdb200_destroy((zstdb200_ctx*)(intptr_t)ctx); }
JNIEXPORT jstring JNICALL FN(lastError0)(JNIEnv* env, jclass cls, jlong ctx) {
  (void)cls; return (*env)->NewStringUTF(env, zstdb200_last_error((zstdb200_ctx*)(intptr_t)ctx));
}
JNIEXPORT jboolean JNICALL FN(isError0)(JNIEnv* env, jclass cls, jint code) { (void)env; (void)cls; return zstdb200_is_error((uint32_t)code) ? JNI_TRUE : JNI_FALSE; }
JNIEXPORT jstring JNICALL FN(errorName0)(JNIEnv* env, jclass cls, jint code) { (void)cls; return (*env)->NewStringUTF(env, code_name((uint32_t)code)); }

JNIEXPORT jint JNICALL FN(decompress0)(JNIEnv* env, jclass cls, jlong ctx, jbyteArray in, jint inOff, jint inLen,
                                       jbyteArray out, jint outOff, jint maxLen) {
  (void)cls;
  if (inOff < 0 || inLen < 0 || outOff < 0 || maxLen < 0 || inOff + (jlong)inLen > (*env)->GetArrayLength(env, in) ||
      outOff + (jlong)maxLen > (*env)->GetArrayLength(env, out)) return (jint)(0u - 1u);                 /* GENERIC */
  jbyte* s = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, in, NULL);
  jbyte* d = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, out, NULL);
  uint32_t r = 0u - 1u;
  if (s && d) r = zstdb200_decompress((zstdb200_ctx*)(intptr_t)ctx, d + outOff, (uint32_t)maxLen, s + inOff, (uint32_t)inLen);
  if (d) (*env)->ReleasePrimitiveArrayCritical(env, out, d, 0);
  if (s) (*env)->ReleasePrimitiveArrayCritical(env, in, s, JNI_ABORT);
  return (jint)r;
}

/* The batch pins every array of the call at once: critical sections may nest, and nothing below calls back into the
 * VM before the last release (the descriptor tables live on the C heap, not on the stack). */
JNIEXPORT jint JNICALL FN(decompressBatch0)(JNIEnv* env, jclass cls, jlong ctx, jobjectArray in, jintArray inOff, jintArray inLen,
                                            jobjectArray out, jintArray outOff, jintArray maxLen, jintArray result) {
  (void)cls;
  const jsize n = (*env)->GetArrayLength(env, in);
  if ((*env)->GetArrayLength(env, out) != n || (*env)->GetArrayLength(env, inOff) < n || (*env)->GetArrayLength(env, inLen) < n ||
      (*env)->GetArrayLength(env, outOff) < n || (*env)->GetArrayLength(env, maxLen) < n || (*env)->GetArrayLength(env, result) < n) return 1;
  if (n == 0) return 0;
  if ((*env)->EnsureLocalCapacity(env, 2 * n + 16) != 0) return 1;               /* two local references per item */
  const void** src = (const void**)calloc((size_t)n, sizeof(void*)); void** dst = (void**)calloc((size_t)n, sizeof(void*));
  uint32_t* srcSize = (uint32_t*)calloc((size_t)n, 4); uint32_t* dstCap = (uint32_t*)calloc((size_t)n, 4); uint32_t* res = (uint32_t*)calloc((size_t)n, 4);
  jbyteArray* ia = (jbyteArray*)calloc((size_t)n, sizeof(jbyteArray)); jbyteArray* oa = (jbyteArray*)calloc((size_t)n, sizeof(jbyteArray));
  jbyte** ip = (jbyte**)calloc((size_t)n, sizeof(jbyte*)); jbyte** op = (jbyte**)calloc((size_t)n, sizeof(jbyte*));
  jint rc = 1;
  if (src && dst && srcSize && dstCap && res && ia && oa && ip && op) {
    jint* io = (*env)->GetIntArrayElements(env, inOff, NULL); jint* il = (*env)->GetIntArrayElements(env, inLen, NULL);
    jint* oo = (*env)->GetIntArrayElements(env, outOff, NULL); jint* ml = (*env)->GetIntArrayElements(env, maxLen, NULL);
    int ok = io && il && oo && ml;
    for (jsize i = 0; ok && i < n; i++) {
      ia[i] = (jbyteArray)(*env)->GetObjectArrayElement(env, in, i); oa[i] = (jbyteArray)(*env)->GetObjectArrayElement(env, out, i);
      if (!ia[i] || !oa[i] || io[i] < 0 || il[i] < 0 || oo[i] < 0 || ml[i] < 0 || io[i] + (jlong)il[i] > (*env)->GetArrayLength(env, ia[i]) ||
          oo[i] + (jlong)ml[i] > (*env)->GetArrayLength(env, oa[i])) ok = 0;
    }
    for (jsize i = 0; ok && i < n; i++) {
      ip[i] = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, ia[i], NULL); op[i] = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, oa[i], NULL);
      if (!ip[i] || !op[i]) ok = 0;
      else { src[i] = ip[i] + io[i]; srcSize[i] = (uint32_t)il[i]; dst[i] = op[i] + oo[i]; dstCap[i] = (uint32_t)ml[i]; }
    }
    if (ok) rc = zstdb200_decompress_batch((zstdb200_ctx*)(intptr_t)ctx, src, srcSize, dst, dstCap, res, (size_t)n);
    for (jsize i = n; i-- > 0;) {
      if (op[i]) (*env)->ReleasePrimitiveArrayCritical(env, oa[i], op[i], 0);
      if (ip[i]) (*env)->ReleasePrimitiveArrayCritical(env, ia[i], ip[i], JNI_ABORT);
    }
    if (io) (*env)->ReleaseIntArrayElements(env, inOff, io, JNI_ABORT);
    if (il) (*env)->ReleaseIntArrayElements(env, inLen, il, JNI_ABORT);
    if (oo) (*env)->ReleaseIntArrayElements(env, outOff, oo, JNI_ABORT);
    if (ml) (*env)->ReleaseIntArrayElements(env, maxLen, ml, JNI_ABORT);
    if (rc == 0) (*env)->SetIntArrayRegion(env, result, 0, n, (const jint*)res);
  }
  free(src); free(dst); free(srcSize); free(dstCap); free(res); free(ia); free(oa); free(ip); free(op);
  return rc;
}

/* ZstdFrameDecompressor.java:922-940: magic check (RuntimeException), then the header's content size, -1 when absent */
JNIEXPORT jlong JNICALL FN(getDecompressedSize0)(JNIEnv* env, jclass cls, jbyteArray in, jint off, jint len) {
  (void)cls;
  if (off < 0 || len < 0 || off + (jlong)len > (*env)->GetArrayLength(env, in) || len < 4) {
    (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/RuntimeException"), "Not enough input bytes");
    return 0;
  }
  uint8_t head[18]; const jint take = len < 18 ? len : 18;
  (*env)->GetByteArrayRegion(env, in, off, take, (jbyte*)head);
  const uint32_t magic = (uint32_t)head[0] | (uint32_t)head[1] << 8 | (uint32_t)head[2] << 16 | (uint32_t)head[3] << 24;
  if (magic != 0xFD2FB528u) {
    (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/RuntimeException"), "Invalid magic prefix");
    return 0;
  }
  if (take >= 5 && (head[4] >> 6) == 0 && !(head[4] & 0x20)) return -1;          /* no content size field */
  return (jlong)zstdb200_get_decompressed_size(head, (uint32_t)take);
}
