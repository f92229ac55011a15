/* TEST INFRASTRUCTURE — a minimal stand-in for the JDK's <jni.h> (not in this image): just the types
 * and JNIEnv entries java/kao_jni.c uses, with the JNI specification's signatures, so that the shim
 * can at least be type-checked against include/kao.h (tests/test_host.py).  Never shipped. */
#ifndef KAO_TEST_JNI_STUB_H
#define KAO_TEST_JNI_STUB_H
#include <stdint.h>
typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef int16_t jshort;
typedef uint8_t jboolean;
typedef jint jsize;
typedef struct _jobject *jobject;
typedef jobject jclass, jstring, jthrowable, jarray, jbyteArray, jshortArray, jintArray, jlongArray;
typedef struct _jmethodID *jmethodID;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
    jclass (*FindClass)(JNIEnv *, const char *);
    jint (*Throw)(JNIEnv *, jthrowable);
    jobject (*NewObject)(JNIEnv *, jclass, jmethodID, ...);
    jmethodID (*GetMethodID)(JNIEnv *, jclass, const char *, const char *);
    jstring (*NewStringUTF)(JNIEnv *, const char *);
    jbyte *(*GetByteArrayElements)(JNIEnv *, jbyteArray, jboolean *);
    jshort *(*GetShortArrayElements)(JNIEnv *, jshortArray, jboolean *);
    jint *(*GetIntArrayElements)(JNIEnv *, jintArray, jboolean *);
    void (*ReleaseByteArrayElements)(JNIEnv *, jbyteArray, jbyte *, jint);
    void (*ReleaseShortArrayElements)(JNIEnv *, jshortArray, jshort *, jint);
    void (*ReleaseIntArrayElements)(JNIEnv *, jintArray, jint *, jint);
    void (*SetLongArrayRegion)(JNIEnv *, jlongArray, jsize, jsize, const jlong *);
};
#endif
