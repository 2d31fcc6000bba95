/* Minimal stand-in for <jni.h> -- TEST INFRASTRUCTURE ONLY (tests/test_abi.py::test_jni_shim_compiles).
 * The image has no JDK, so jni/mc_jni.cpp could never go through a compiler; this header declares just the JNI types and the
 * JNIEnv members the shim uses, with the signatures of the JNI specification (Java SE "JNI Functions"), so that the shim is
 * at least type-checked against include/mc_b200.h.  It is not a JNI implementation. */
#ifndef MC_TEST_JNI_STUB_H
#define MC_TEST_JNI_STUB_H
#include <stdint.h>
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
typedef int32_t jint;
typedef int64_t jlong;
typedef int16_t jshort;
typedef uint8_t jboolean;
typedef int8_t jbyte;
typedef double jdouble;
typedef jint jsize;
class _jobject {};
class _jclass : public _jobject {};
class _jstring : public _jobject {};
class _jarray : public _jobject {};
class _jobjectArray : public _jarray {};
class _jlongArray : public _jarray {};
class _jshortArray : public _jarray {};
class _jbyteArray : public _jarray {};
typedef _jobject* jobject;
typedef _jclass* jclass;
typedef _jstring* jstring;
typedef _jarray* jarray;
typedef _jobjectArray* jobjectArray;
typedef _jlongArray* jlongArray;
typedef _jshortArray* jshortArray;
typedef _jbyteArray* jbyteArray;
struct JNIEnv {
    jclass FindClass(const char*);
    jint ThrowNew(jclass, const char*);
    jsize GetArrayLength(jarray);
    jobject GetObjectArrayElement(jobjectArray, jsize);
    const char* GetStringUTFChars(jstring, jboolean*);
    void ReleaseStringUTFChars(jstring, const char*);
    void DeleteLocalRef(jobject);
    jlong* GetLongArrayElements(jlongArray, jboolean*);
    void ReleaseLongArrayElements(jlongArray, jlong*, jint);
    jshortArray NewShortArray(jsize);
    void SetShortArrayRegion(jshortArray, jsize, jsize, const jshort*);
    jbyte* GetByteArrayElements(jbyteArray, jboolean*);
    void ReleaseByteArrayElements(jbyteArray, jbyte*, jint);
    jbyteArray NewByteArray(jsize);
    void SetByteArrayRegion(jbyteArray, jsize, jsize, const jbyte*);
};
#endif
