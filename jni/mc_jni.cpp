// JNI shim: the ~100 lines the reference's maintainers would add to call libmcgpu from the Java tools.
//
// SOURCE ONLY in this image (no JDK, no <jni.h>): it is compiled only where <jni.h> exists
//   g++ -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include mc_jni.cpp -L../metacherchant_b200/csrc -lmcgpu -o libmcjni.so
// Correctness of everything below the shim is proven through the C ABI + CLI tests; the Java side is described in
// INTEGRATION.md (class io.NativeCounter, used from EnvironmentFinderMain.loadInput()/runImpl()).
#if defined(__has_include)
#if __has_include(<jni.h>)
#define MC_HAVE_JNI 1
#endif
#endif

#ifdef MC_HAVE_JNI
#include <jni.h>

#include <string>
#include <vector>

#include "mc_b200.h"

namespace {
// ExecutionFailedException(String) is what the reference throws from loadReads (src/io/IOUtils.java:292)
void throw_execution_failed(JNIEnv* env) {
    jclass cls = env->FindClass("ru/ifmo/genetics/utils/tool/ExecutionFailedException");
    if (cls) env->ThrowNew(cls, mc_last_error());
}
std::vector<std::string> to_strings(JNIEnv* env, jobjectArray arr) {
    std::vector<std::string> out;
    jsize n = env->GetArrayLength(arr);
    for (jsize i = 0; i < n; i++) {
        jstring s = (jstring)env->GetObjectArrayElement(arr, i);
        const char* c = env->GetStringUTFChars(s, nullptr);
        out.emplace_back(c);
        env->ReleaseStringUTFChars(s, c);
        env->DeleteLocalRef(s);
    }
    return out;
}
std::vector<const char*> c_strs(const std::vector<std::string>& v) {
    std::vector<const char*> p;
    for (auto& s : v) p.push_back(s.c_str());
    return p;
}
}  // namespace

extern "C" {

// long io.NativeCounter.createContext(int device)
JNIEXPORT jlong JNICALL Java_io_NativeCounter_createContext(JNIEnv* env, jclass, jint device) {
    mc_ctx* ctx = nullptr;
    if (mc_ctx_create(device, &ctx) != MC_OK) { throw_execution_failed(env); return 0; }
    return (jlong)ctx;
}
JNIEXPORT void JNICALL Java_io_NativeCounter_destroyContext(JNIEnv*, jclass, jlong ctx) { mc_ctx_destroy((mc_ctx*)ctx); }

// long io.NativeCounter.loadReads(long ctx, String[] files, int k)   ~ IOUtils.loadReads (src/io/IOUtils.java:217)
JNIEXPORT jlong JNICALL Java_io_NativeCounter_loadReads(JNIEnv* env, jclass, jlong ctx, jobjectArray files, jint k) {
    std::vector<std::string> paths = to_strings(env, files);
    std::vector<const char*> p = c_strs(paths);
    mc_table* t = nullptr;
    if (mc_count_files((mc_ctx*)ctx, p.data(), (int)p.size(), k, &t) != MC_OK) { throw_execution_failed(env); return 0; }
    return (jlong)t;
}
// long io.NativeCounter.size(long table)                               ~ BigLong2ShortHashMap.size()
JNIEXPORT jlong JNICALL Java_io_NativeCounter_size(JNIEnv*, jclass, jlong table) { return (jlong)mc_table_size((mc_table*)table); }
// short[] io.NativeCounter.get(long table, long[] keys)                ~ BigLong2ShortHashMap.get(long), batched
JNIEXPORT jshortArray JNICALL Java_io_NativeCounter_get(JNIEnv* env, jclass, jlong table, jlongArray keys) {
    jsize n = env->GetArrayLength(keys);
    jlong* k = env->GetLongArrayElements(keys, nullptr);
    std::vector<int16_t> out((size_t)n);
    int rc = mc_table_get((mc_table*)table, (const uint64_t*)k, (size_t)n, out.data());
    env->ReleaseLongArrayElements(keys, k, JNI_ABORT);
    if (rc != MC_OK) { throw_execution_failed(env); return nullptr; }
    jshortArray res = env->NewShortArray(n);
    env->SetShortArrayRegion(res, 0, n, (const jshort*)out.data());
    return res;
}
JNIEXPORT void JNICALL Java_io_NativeCounter_freeTable(JNIEnv*, jclass, jlong table) { mc_table_free((mc_table*)table); }

// void io.NativeCounter.findEnvironments(long ctx, long table, String[] seqs, String[] names, int k, int minCov,
//        long maxKmers, long maxRadius, boolean bothDirs, int chunkLength, String outDir)
//   ~ one OneSequenceCalculator(...).run() per target (src/tools/EnvironmentFinderMain.java:218-225)
JNIEXPORT void JNICALL Java_io_NativeCounter_findEnvironments(JNIEnv* env, jclass, jlong ctx, jlong table, jobjectArray seqs, jobjectArray names,
                                                              jint k, jint minCov, jlong maxKmers, jlong maxRadius, jboolean bothDirs,
                                                              jint chunkLength, jstring outDir) {
    std::vector<std::string> s = to_strings(env, seqs), nm = to_strings(env, names);
    std::vector<const char*> sp = c_strs(s), np = c_strs(nm);
    const char* od = env->GetStringUTFChars(outDir, nullptr);
    mc_table* solid = nullptr;
    int rc = mc_table_filter((mc_ctx*)ctx, (mc_table*)table, minCov, &solid);
    if (rc == MC_OK) {
        rc = mc_env_find((mc_ctx*)ctx, solid, sp.data(), np.data(), (int)sp.size(), k, minCov, maxKmers, maxRadius, bothDirs ? 1 : 0, chunkLength, od, nullptr);
        mc_table_free(solid);
    }
    env->ReleaseStringUTFChars(outDir, od);
    if (rc != MC_OK) throw_execution_failed(env);
}

// void io.NativeCounter.findEnvironmentsEx(long ctx, long table, String[] seqs, String[] names, String[] hicSeqs, int k, int minCov,
//        long maxKmers, long maxRadius, boolean bothDirs, int chunkLength, boolean merge, boolean trim, String outDir)
//   ~ the same loop with --merge (one OneSequenceCalculator over every target and Hi-C sequence, output <out>/merged/),
//     --hicseq and --trim (src/tools/EnvironmentFinderMain.java:146-150,218-233, src/algo/OneSequenceCalculator.java:168-192,208-216,241-262)
JNIEXPORT void JNICALL Java_io_NativeCounter_findEnvironmentsEx(JNIEnv* env, jclass, jlong ctx, jlong table, jobjectArray seqs, jobjectArray names,
                                                                jobjectArray hicSeqs, jint k, jint minCov, jlong maxKmers, jlong maxRadius,
                                                                jboolean bothDirs, jint chunkLength, jboolean merge, jboolean trim, jstring outDir) {
    std::vector<std::string> s = to_strings(env, seqs), nm = to_strings(env, names), hic = to_strings(env, hicSeqs);
    std::vector<const char*> sp = c_strs(s), np = c_strs(nm), hp = c_strs(hic);
    const char* od = env->GetStringUTFChars(outDir, nullptr);
    mc_table* solid = nullptr;
    int rc = mc_table_filter((mc_ctx*)ctx, (mc_table*)table, minCov, &solid);
    if (rc == MC_OK) {
        rc = mc_env_find_ex((mc_ctx*)ctx, solid, sp.data(), np.data(), (int)sp.size(), hp.data(), (int)hp.size(), k, minCov, maxKmers, maxRadius,
                            bothDirs ? 1 : 0, chunkLength, merge ? 1 : 0, trim ? 1 : 0, od, nullptr);
        mc_table_free(solid);
    }
    env->ReleaseStringUTFChars(outDir, od);
    if (rc != MC_OK) throw_execution_failed(env);
}

// void io.NativeCounter.multiEnvironment(long ctx, String[] envFiles, String seqFasta, int geneId, String outDir)
//   ~ MultiSequenceCalculator(...).run() + printGene + printProbability (src/tools/EnvironmentFinderMultiMain.java:81-91)
JNIEXPORT void JNICALL Java_io_NativeCounter_multiEnvironment(JNIEnv* env, jclass, jlong ctx, jobjectArray envFiles, jstring seqFasta, jint geneId,
                                                              jstring outDir) {
    std::vector<std::string> e = to_strings(env, envFiles);
    std::vector<const char*> ep = c_strs(e);
    const char* sf = env->GetStringUTFChars(seqFasta, nullptr);
    const char* od = env->GetStringUTFChars(outDir, nullptr);
    int rc = mc_env_multi((mc_ctx*)ctx, ep.data(), (int)ep.size(), sf, geneId, od);
    env->ReleaseStringUTFChars(seqFasta, sf);
    env->ReleaseStringUTFChars(outDir, od);
    if (rc != MC_OK) throw_execution_failed(env);
}

// byte[] io.NativeCounter.classifyReads(long ctx, long table, byte[] fastq, double z, double foundThreshold, boolean correction)
//   ~ ReadsFinderInGraph.findRead / findReadWithCorrection for every read of the buffer (src/algo/ReadsFinderInGraph.java:37-123);
//     PairFinder.run (src/algo/PairFinder.java:30-55) keeps its bookkeeping on the Java side, fed from the returned flags
JNIEXPORT jbyteArray JNICALL Java_io_NativeCounter_classifyReads(JNIEnv* env, jclass, jlong ctx, jlong table, jbyteArray fastq, jdouble z,
                                                                 jdouble foundThreshold, jboolean correction) {
    const jsize len = env->GetArrayLength(fastq);
    jbyte* buf = env->GetByteArrayElements(fastq, nullptr);
    size_t n = 0;
    int rc = mc_classify_bytes((mc_ctx*)ctx, (const mc_table*)table, buf, (size_t)len, MC_FMT_FASTQ, z, foundThreshold, correction ? 1 : 0, nullptr, 0, &n);
    std::vector<uint8_t> found(n);
    if (rc == MC_OK && n)
        rc = mc_classify_bytes((mc_ctx*)ctx, (const mc_table*)table, buf, (size_t)len, MC_FMT_FASTQ, z, foundThreshold, correction ? 1 : 0, found.data(), n, &n);
    env->ReleaseByteArrayElements(fastq, buf, JNI_ABORT);
    if (rc != MC_OK) { throw_execution_failed(env); return nullptr; }
    jbyteArray res = env->NewByteArray((jsize)n);
    env->SetByteArrayRegion(res, 0, (jsize)n, (const jbyte*)found.data());
    return res;
}

}  // extern "C"
#endif  // MC_HAVE_JNI
