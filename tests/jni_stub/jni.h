/* tests/jni_stub/jni.h -- a MINIMAL stand-in for the JDK's <jni.h>, only so that jni/swb_jni.c can be compiled and
 * link-checked in an image without a JDK (tests/test_abi_and_host.py).  It declares exactly the JNI types and the
 * JNIEnv entries swb_jni.c uses, with the JNI specification's signatures; the member ORDER is not the real
 * function table's, so an object built against this header must never be loaded into a JVM. */
#ifndef SWB_JNI_STUB_H
#define SWB_JNI_STUB_H
#include <stdint.h>
#define SWB_JNI_STUB 1
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
typedef int32_t jint; typedef int64_t jlong; typedef int8_t jbyte; typedef uint8_t jboolean; typedef jint jsize;
typedef void *jobject; typedef jobject jclass; typedef jobject jstring; typedef jobject jarray; typedef jobject jthrowable;
typedef jarray jbyteArray; typedef jarray jintArray; typedef jarray jlongArray; typedef jarray jobjectArray;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
    jclass (*FindClass)(JNIEnv *, const char *);
    jint (*ThrowNew)(JNIEnv *, jclass, const char *);
    jstring (*NewStringUTF)(JNIEnv *, const char *);
    jsize (*GetArrayLength)(JNIEnv *, jarray);
    void *(*GetPrimitiveArrayCritical)(JNIEnv *, jarray, jboolean *);
    void (*ReleasePrimitiveArrayCritical)(JNIEnv *, jarray, void *, jint);
    jbyteArray (*NewByteArray)(JNIEnv *, jsize);
    jintArray (*NewIntArray)(JNIEnv *, jsize);
    jlongArray (*NewLongArray)(JNIEnv *, jsize);
    void (*GetByteArrayRegion)(JNIEnv *, jbyteArray, jsize, jsize, jbyte *);
    void (*SetByteArrayRegion)(JNIEnv *, jbyteArray, jsize, jsize, const jbyte *);
    void (*SetIntArrayRegion)(JNIEnv *, jintArray, jsize, jsize, const jint *);
    void (*SetLongArrayRegion)(JNIEnv *, jlongArray, jsize, jsize, const jlong *);
    jobjectArray (*NewObjectArray)(JNIEnv *, jsize, jclass, jobject);
    void (*SetObjectArrayElement)(JNIEnv *, jobjectArray, jsize, jobject);
};
#endif
