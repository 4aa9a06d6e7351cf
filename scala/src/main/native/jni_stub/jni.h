/* Minimal stand-in for <jni.h>: only what sparkb200_jni.c uses, so that the shim can be TYPE-CHECKED against include/spark_b200.h
   on a machine without a JDK (tests/test_capi_cpu.py).  A real build uses the JDK header instead. */
#include <stdint.h>
#include <stddef.h>
typedef int32_t jint; typedef int64_t jlong; typedef int8_t jbyte; typedef uint8_t jboolean; typedef jint jsize;
typedef void *jobject; typedef jobject jclass; typedef jobject jarray; typedef jarray jintArray; typedef jarray jlongArray; typedef jarray jbyteArray; typedef jarray jbooleanArray; typedef jarray jobjectArray;
#define JNIEXPORT
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv *, const char *);
  jint (*ThrowNew)(JNIEnv *, jclass, const char *);
  jsize (*GetArrayLength)(JNIEnv *, jarray);
  jint *(*GetIntArrayElements)(JNIEnv *, jintArray, jboolean *);
  jlong *(*GetLongArrayElements)(JNIEnv *, jlongArray, jboolean *);
  jboolean *(*GetBooleanArrayElements)(JNIEnv *, jbooleanArray, jboolean *);
  void (*ReleaseIntArrayElements)(JNIEnv *, jintArray, jint *, jint);
  void (*ReleaseLongArrayElements)(JNIEnv *, jlongArray, jlong *, jint);
  void (*ReleaseBooleanArrayElements)(JNIEnv *, jbooleanArray, jboolean *, jint);
  jbyteArray (*NewByteArray)(JNIEnv *, jsize);
  void (*SetByteArrayRegion)(JNIEnv *, jbyteArray, jsize, jsize, const jbyte *);
  void (*GetByteArrayRegion)(JNIEnv *, jbyteArray, jsize, jsize, jbyte *);
  jlongArray (*NewLongArray)(JNIEnv *, jsize);
  void (*SetLongArrayRegion)(JNIEnv *, jlongArray, jsize, jsize, const jlong *);
  jobject (*GetObjectArrayElement)(JNIEnv *, jobjectArray, jsize);
};
