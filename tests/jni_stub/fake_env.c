/* A JNIEnv made of plain C arrays — TEST AID ONLY.  It lets tests/test_bindings.py drive the JNI glue
 * (bindings/java/zstdb200_jni.c) without a JVM: byte[] / int[] / byte[][] are FakeArray objects, strings are C strings,
 * ThrowNew records the message.  Enough to check the glue's argument handling, pinning order and ABI calls. */
#define _GNU_SOURCE
#include <jni.h>
#include <stdlib.h>
#include <string.h>

typedef struct FakeArray { jsize len; int kind; /* 0 bytes, 1 ints, 2 objects */ void* data; int pinned; } FakeArray;
static char g_thrown[256];
static int g_critical_depth, g_violations;   /* JNI calls made inside a critical region (forbidden by the JNI spec) */

static void outside(void) { if (g_critical_depth) g_violations++; }
static jclass f_FindClass(JNIEnv* e, const char* n) { (void)e; outside(); return (jclass)n; }
static jint f_ThrowNew(JNIEnv* e, jclass c, const char* m) { (void)e; (void)c; outside(); strncpy(g_thrown, m, sizeof g_thrown - 1); return 0; }
static jstring f_NewStringUTF(JNIEnv* e, const char* s) { (void)e; outside(); return (jstring)strdup(s ? s : ""); }
static jint f_EnsureLocalCapacity(JNIEnv* e, jint n) { (void)e; (void)n; outside(); return 0; }
static jsize f_GetArrayLength(JNIEnv* e, jarray a) { (void)e; outside(); return ((FakeArray*)a)->len; }
static jobject f_GetObjectArrayElement(JNIEnv* e, jobjectArray a, jsize i) { (void)e; outside(); return ((jobject*)((FakeArray*)a)->data)[i]; }
static jint* f_GetIntArrayElements(JNIEnv* e, jintArray a, jboolean* c) { (void)e; outside(); if (c) *c = 0; return (jint*)((FakeArray*)a)->data; }
static void f_ReleaseIntArrayElements(JNIEnv* e, jintArray a, jint* p, jint m) { (void)e; (void)a; (void)p; (void)m; outside(); }
static void f_GetByteArrayRegion(JNIEnv* e, jbyteArray a, jsize s, jsize n, jbyte* b) { (void)e; outside(); memcpy(b, (jbyte*)((FakeArray*)a)->data + s, (size_t)n); }
static void f_SetIntArrayRegion(JNIEnv* e, jintArray a, jsize s, jsize n, const jint* b) { (void)e; outside(); memcpy((jint*)((FakeArray*)a)->data + s, b, (size_t)n * 4); }
static void* f_GetCritical(JNIEnv* e, jarray a, jboolean* c) { (void)e; if (c) *c = 0; ((FakeArray*)a)->pinned++; g_critical_depth++; return ((FakeArray*)a)->data; }
static void f_ReleaseCritical(JNIEnv* e, jarray a, void* p, jint m) { (void)e; (void)p; (void)m; ((FakeArray*)a)->pinned--; g_critical_depth--; }

static const struct JNINativeInterface_ g_table = {
  f_FindClass, f_ThrowNew, f_NewStringUTF, f_EnsureLocalCapacity, f_GetArrayLength, f_GetObjectArrayElement, f_GetIntArrayElements,
  f_ReleaseIntArrayElements, f_GetByteArrayRegion, f_SetIntArrayRegion, f_GetCritical, f_ReleaseCritical,
};
static JNIEnv g_env = &g_table;

JNIEnv* fake_env(void) { return &g_env; }
FakeArray* fake_array(int kind, jsize len, void* data) { FakeArray* a = (FakeArray*)calloc(1, sizeof *a); a->kind = kind; a->len = len; a->data = data; return a; }
const char* fake_thrown(void) { return g_thrown; }
void fake_clear(void) { g_thrown[0] = 0; }
int fake_violations(void) { return g_violations; }
int fake_pinned(const FakeArray* a) { return a->pinned; }
const char* fake_string(jstring s) { return (const char*)s; }
