/*
 * mrs_jni.c -- JNI shim between shared.NativeEngine (java/shared/NativeEngine.java) and the C ABI of libmrs_b200.so
 * (include/mrs_b200.h).  Java 8.  One function per native method; handles travel as jlong; a non-zero mrs_status
 * becomes a RuntimeException carrying mrs_last_error().  No arithmetic happens here.
 *
 * Reference surface served (src/main/scala/shared/predictions.scala = "P:"): load P:35-49, the baseline family
 * P:94-237 and its Spark twins P:246-391, the personalized / kNN family P:400-649, recommendations P:651-674, MAE
 * P:69-86 / P:256-258 -- through the Scala facade in scala/shared/predictions.scala.
 *
 * Build (where a JDK exists; there is none in the image this was written in, so the file is checked against a minimal
 * jni.h stub by tests/test_jni_shim.py):
 *   gcc -shared -fPIC -O2 -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude jni/mrs_jni.c \
 *       -Lmovie-recommender-system_b200 -lmrs_b200 -o libmrs_jni.so
 */
#include <jni.h>
#include <stdint.h>
#include <string.h>

#include "mrs_b200.h"

#define H(type, x) ((type*)(intptr_t)(x))
#define J(p) ((jlong)(intptr_t)(p))

static void throw_last(JNIEnv* env) {
  jclass ex = (*env)->FindClass(env, "java/lang/RuntimeException");
  if (ex) (*env)->ThrowNew(env, ex, mrs_last_error());
}
#define CHECK(call, fail)                 \
  do {                                    \
    if ((call) != MRS_OK) {               \
      throw_last(env);                    \
      return fail;                        \
    }                                     \
  } while (0)

/* ------------------------------------------------------------------ engine */
JNIEXPORT jlong JNICALL Java_shared_NativeEngine_engineCreate(JNIEnv* env, jclass c, jint device) {
  mrs_engine* e = NULL;
  CHECK(mrs_engine_create(device, NULL, &e), 0);
  return J(e);
}
JNIEXPORT void JNICALL Java_shared_NativeEngine_engineDestroy(JNIEnv* env, jclass c, jlong e) { mrs_engine_destroy(H(mrs_engine, e)); }
JNIEXPORT jstring JNICALL Java_shared_NativeEngine_version(JNIEnv* env, jclass c) { return (*env)->NewStringUTF(env, mrs_version()); }

/* ------------------------------------------------------------------ rating sets (P:35-49, .collect() at predict/Baseline.scala:40-42) */
JNIEXPORT jlong JNICALL Java_shared_NativeEngine_ratingsFromCoo(JNIEnv* env, jclass c, jlong eng, jintArray users, jintArray items,
                                                                jdoubleArray ratings) {
  const jsize n = (*env)->GetArrayLength(env, users);
  if ((*env)->GetArrayLength(env, items) != n || (*env)->GetArrayLength(env, ratings) != n) {
    jclass ex = (*env)->FindClass(env, "java/lang/IllegalArgumentException");
    if (ex) (*env)->ThrowNew(env, ex, "users, items, ratings must have the same length");
    return 0;
  }
  jint* u = (jint*)(*env)->GetPrimitiveArrayCritical(env, users, NULL);
  jint* i = (jint*)(*env)->GetPrimitiveArrayCritical(env, items, NULL);
  jdouble* r = (jdouble*)(*env)->GetPrimitiveArrayCritical(env, ratings, NULL);
  mrs_ratings* out = NULL;
  int32_t s = MRS_ERR_NOMEM;
  if (u && i && r) s = mrs_ratings_from_coo(H(mrs_engine, eng), (const int32_t*)u, (const int32_t*)i, (const double*)r, (int64_t)n, 0, 0, &out);
  if (r) (*env)->ReleasePrimitiveArrayCritical(env, ratings, r, JNI_ABORT);
  if (i) (*env)->ReleasePrimitiveArrayCritical(env, items, i, JNI_ABORT);
  if (u) (*env)->ReleasePrimitiveArrayCritical(env, users, u, JNI_ABORT);
  if (s != MRS_OK) { throw_last(env); return 0; }
  return J(out);
}

/* compact form: the rating as one byte = 2 x rating (half-star data), 9 bytes per Rating over PCIe instead of 16 */
JNIEXPORT jlong JNICALL Java_shared_NativeEngine_ratingsFromCodes(JNIEnv* env, jclass c, jlong eng, jintArray users, jintArray items,
                                                                  jbyteArray codes) {
  const jsize n = (*env)->GetArrayLength(env, users);
  if ((*env)->GetArrayLength(env, items) != n || (*env)->GetArrayLength(env, codes) != n) {
    jclass ex = (*env)->FindClass(env, "java/lang/IllegalArgumentException");
    if (ex) (*env)->ThrowNew(env, ex, "users, items, codes must have the same length");
    return 0;
  }
  jint* u = (jint*)(*env)->GetPrimitiveArrayCritical(env, users, NULL);
  jint* i = (jint*)(*env)->GetPrimitiveArrayCritical(env, items, NULL);
  jbyte* k = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, codes, NULL);
  mrs_ratings* out = NULL;
  int32_t s = MRS_ERR_NOMEM;
  if (u && i && k) s = mrs_ratings_from_coo_codes(H(mrs_engine, eng), (const int32_t*)u, (const int32_t*)i, (const uint8_t*)k, (int64_t)n, 0, 0, &out);
  if (k) (*env)->ReleasePrimitiveArrayCritical(env, codes, k, JNI_ABORT);
  if (i) (*env)->ReleasePrimitiveArrayCritical(env, items, i, JNI_ABORT);
  if (u) (*env)->ReleasePrimitiveArrayCritical(env, users, u, JNI_ABORT);
  if (s != MRS_OK) { throw_last(env); return 0; }
  return J(out);
}

/* text of a ratings file as read by the JVM (or an HDFS block): parsed on the device with the rules of P:35-49 */
JNIEXPORT jlong JNICALL Java_shared_NativeEngine_ratingsFromText(JNIEnv* env, jclass c, jlong eng, jbyteArray text, jstring sep) {
  const jsize n = (*env)->GetArrayLength(env, text);
  const char* sp = (*env)->GetStringUTFChars(env, sep, NULL);
  jbyte* t = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, text, NULL);
  mrs_ratings* out = NULL;
  int32_t s = MRS_ERR_NOMEM;
  if (t && sp) s = mrs_ratings_from_text(H(mrs_engine, eng), (const char*)t, (int64_t)n, sp, &out);
  if (t) (*env)->ReleasePrimitiveArrayCritical(env, text, t, JNI_ABORT);
  if (sp) (*env)->ReleaseStringUTFChars(env, sep, sp);
  if (s != MRS_OK) { throw_last(env); return 0; }
  return J(out);
}

JNIEXPORT jlong JNICALL Java_shared_NativeEngine_ratingsFromFile(JNIEnv* env, jclass c, jlong eng, jstring path, jstring sep) {
  const char* p = (*env)->GetStringUTFChars(env, path, NULL);
  const char* sp = (*env)->GetStringUTFChars(env, sep, NULL);
  mrs_ratings* out = NULL;
  int32_t s = MRS_ERR_NOMEM;
  if (p && sp) s = mrs_ratings_from_file(H(mrs_engine, eng), p, sp, &out);
  if (sp) (*env)->ReleaseStringUTFChars(env, sep, sp);
  if (p) (*env)->ReleaseStringUTFChars(env, path, p);
  if (s != MRS_OK) { throw_last(env); return 0; }
  return J(out);
}

JNIEXPORT jlong JNICALL Java_shared_NativeEngine_ratingsCount(JNIEnv* env, jclass c, jlong r) {
  int64_t n = 0;
  CHECK(mrs_ratings_info(H(mrs_ratings, r), &n, NULL, NULL, NULL), 0);
  return (jlong)n;
}
JNIEXPORT void JNICALL Java_shared_NativeEngine_ratingsDestroy(JNIEnv* env, jclass c, jlong r) { mrs_ratings_destroy(H(mrs_ratings, r)); }

/* ------------------------------------------------------------------ fit (eager part of P:101-214 / P:265-368) */
JNIEXPORT jlong JNICALL Java_shared_NativeEngine_fit(JNIEnv* env, jclass c, jlong eng, jlong train) {
  mrs_model* m = NULL;
  CHECK(mrs_fit(H(mrs_engine, eng), H(mrs_ratings, train), &m), 0);
  return J(m);
}
JNIEXPORT void JNICALL Java_shared_NativeEngine_modelDestroy(JNIEnv* env, jclass c, jlong m) { mrs_model_destroy(H(mrs_model, m)); }

/* average P:94 / getGlobalAvg P:265 */
JNIEXPORT jdouble JNICALL Java_shared_NativeEngine_globalAvg(JNIEnv* env, jclass c, jlong m) {
  double out = 0.0;
  CHECK(mrs_model_scalar(H(mrs_model, m), MRS_GLOBAL_AVG, &out), 0.0);
  return out;
}
/* one entry of usersAvg / itemsAvg / itemsAvgDev with the reference's fallbacks (P:126, P:147, P:197) */
JNIEXPORT jdouble JNICALL Java_shared_NativeEngine_modelLookup(JNIEnv* env, jclass c, jlong m, jint vecKind, jint id) {
  double out = 0.0;
  CHECK(mrs_model_lookup(H(mrs_model, m), vecKind, id, &out, NULL), 0.0);
  return out;
}
/* table size (max id + 1) of usersAvg / itemsAvg / itemsAvgDev */
JNIEXPORT jint JNICALL Java_shared_NativeEngine_modelVectorSize(JNIEnv* env, jclass c, jlong m, jint vecKind) {
  int64_t n = 0;
  CHECK(mrs_model_vector(H(mrs_model, m), vecKind, NULL, NULL, 0, &n), 0);
  return (jint)n;
}
/* whole table, direct-indexed by id; counts[id] == 0 marks ids that do not occur in the train set (P:113, P:134, P:176) */
JNIEXPORT void JNICALL Java_shared_NativeEngine_modelVector(JNIEnv* env, jclass c, jlong m, jint vecKind, jdoubleArray valsOut,
                                                            jintArray countsOut) {
  const jsize cap = (*env)->GetArrayLength(env, valsOut);
  jdouble* v = (jdouble*)(*env)->GetPrimitiveArrayCritical(env, valsOut, NULL);
  jint* k = (jint*)(*env)->GetPrimitiveArrayCritical(env, countsOut, NULL);
  int64_t n = 0;
  int32_t s = MRS_ERR_NOMEM;
  if (v && k) s = mrs_model_vector(H(mrs_model, m), vecKind, (double*)v, (int32_t*)k, (int64_t)cap, &n);
  if (k) (*env)->ReleasePrimitiveArrayCritical(env, countsOut, k, 0);
  if (v) (*env)->ReleasePrimitiveArrayCritical(env, valsOut, v, 0);
  if (s != MRS_OK) throw_last(env);
}

/* ------------------------------------------------------------------ similarity (P:400-481, P:596-649) */
JNIEXPORT jlong JNICALL Java_shared_NativeEngine_fitSimilarity(JNIEnv* env, jclass c, jlong m, jint simKind, jint k) {
  mrs_sim* s = NULL;
  CHECK(mrs_fit_similarity(H(mrs_model, m), simKind, k, &s), 0);
  return J(s);
}
JNIEXPORT void JNICALL Java_shared_NativeEngine_simSetK(JNIEnv* env, jclass c, jlong s, jint k) {
  if (mrs_sim_set_k(H(mrs_sim, s), k) != MRS_OK) throw_last(env);
}
JNIEXPORT jdouble JNICALL Java_shared_NativeEngine_similarity(JNIEnv* env, jclass c, jlong s, jint u, jint v) {
  double out = 0.0;
  CHECK(mrs_similarity(H(mrs_sim, s), u, v, &out), 0.0);
  return out;
}
/* getNeighbors P:596-617: first min(k, candidates) neighbours of u, (similarity desc, id asc); returns how many */
JNIEXPORT jint JNICALL Java_shared_NativeEngine_neighbors(JNIEnv* env, jclass c, jlong s, jint u, jint k, jintArray idsOut,
                                                          jdoubleArray simsOut) {
  const jsize cap = (*env)->GetArrayLength(env, idsOut);
  jint* ids = (jint*)(*env)->GetPrimitiveArrayCritical(env, idsOut, NULL);
  jdouble* sims = (jdouble*)(*env)->GetPrimitiveArrayCritical(env, simsOut, NULL);
  int32_t n = 0, st = MRS_ERR_NOMEM;
  if (ids && sims) st = mrs_neighbors(H(mrs_sim, s), u, k, (int32_t*)ids, (double*)sims, (int32_t)cap, &n);
  if (sims) (*env)->ReleasePrimitiveArrayCritical(env, simsOut, sims, 0);
  if (ids) (*env)->ReleasePrimitiveArrayCritical(env, idsOut, ids, 0);
  if (st != MRS_OK) { throw_last(env); return 0; }
  return n;
}
/* computeNormalizeDeviation P:155-169 (which = 0) / preprocessedRating P:470-481 (which = 1), user-major order; returns the count */
JNIEXPORT jlong JNICALL Java_shared_NativeEngine_entryValues(JNIEnv* env, jclass c, jlong s, jint which, jintArray usersOut, jintArray itemsOut,
                                                             jdoubleArray valsOut) {
  int64_t n = 0;
  if (!usersOut) {  /* size query */
    CHECK(mrs_sim_entry_values(H(mrs_sim, s), which, NULL, NULL, NULL, 0, &n), 0);
    return (jlong)n;
  }
  const jsize cap = (*env)->GetArrayLength(env, valsOut);
  jint* u = (jint*)(*env)->GetPrimitiveArrayCritical(env, usersOut, NULL);
  jint* i = (jint*)(*env)->GetPrimitiveArrayCritical(env, itemsOut, NULL);
  jdouble* v = (jdouble*)(*env)->GetPrimitiveArrayCritical(env, valsOut, NULL);
  int32_t st = MRS_ERR_NOMEM;
  if (u && i && v) st = mrs_sim_entry_values(H(mrs_sim, s), which, (int32_t*)u, (int32_t*)i, (double*)v, (int64_t)cap, &n);
  if (v) (*env)->ReleasePrimitiveArrayCritical(env, valsOut, v, 0);
  if (i) (*env)->ReleasePrimitiveArrayCritical(env, itemsOut, i, 0);
  if (u) (*env)->ReleasePrimitiveArrayCritical(env, usersOut, u, 0);
  if (st != MRS_OK) { throw_last(env); return 0; }
  return (jlong)n;
}
JNIEXPORT void JNICALL Java_shared_NativeEngine_simDestroy(JNIEnv* env, jclass c, jlong s) { mrs_sim_destroy(H(mrs_sim, s)); }

/* ------------------------------------------------------------------ prediction, MAE, recommendations */
JNIEXPORT jdouble JNICALL Java_shared_NativeEngine_predictOne(JNIEnv* env, jclass c, jlong m, jlong sim, jint predKind, jint u, jint i) {
  double out = 0.0;
  int32_t uu = u, ii = i;
  CHECK(mrs_predict(H(mrs_model, m), H(mrs_sim, sim), predKind, &uu, &ii, 1, &out), 0.0);
  return out;
}
JNIEXPORT void JNICALL Java_shared_NativeEngine_predict(JNIEnv* env, jclass c, jlong m, jlong sim, jint predKind, jintArray users,
                                                        jintArray items, jdoubleArray out) {
  const jsize n = (*env)->GetArrayLength(env, users);
  jint* u = (jint*)(*env)->GetPrimitiveArrayCritical(env, users, NULL);
  jint* i = (jint*)(*env)->GetPrimitiveArrayCritical(env, items, NULL);
  jdouble* o = (jdouble*)(*env)->GetPrimitiveArrayCritical(env, out, NULL);
  int32_t st = MRS_ERR_NOMEM;
  if (u && i && o) st = mrs_predict(H(mrs_model, m), H(mrs_sim, sim), predKind, (const int32_t*)u, (const int32_t*)i, (int64_t)n, (double*)o);
  if (o) (*env)->ReleasePrimitiveArrayCritical(env, out, o, 0);
  if (i) (*env)->ReleasePrimitiveArrayCritical(env, items, i, JNI_ABORT);
  if (u) (*env)->ReleasePrimitiveArrayCritical(env, users, u, JNI_ABORT);
  if (st != MRS_OK) throw_last(env);
}
/* MAE P:69-86 / MeanAbsoluteErrorSpark P:256-258: one fused native call */
JNIEXPORT jdouble JNICALL Java_shared_NativeEngine_mae(JNIEnv* env, jclass c, jlong m, jlong sim, jint predKind, jlong test) {
  double out = 0.0;
  CHECK(mrs_mae(H(mrs_model, m), H(mrs_sim, sim), predKind, H(mrs_ratings, test), &out), 0.0);
  return out;
}
/* recommendations P:651-674; returns how many were written */
JNIEXPORT jint JNICALL Java_shared_NativeEngine_recommend(JNIEnv* env, jclass c, jlong m, jlong sim, jint predKind, jint user, jint n,
                                                          jintArray itemsOut, jdoubleArray scoresOut) {
  jint* it = (jint*)(*env)->GetPrimitiveArrayCritical(env, itemsOut, NULL);
  jdouble* sc = (jdouble*)(*env)->GetPrimitiveArrayCritical(env, scoresOut, NULL);
  int32_t w = 0, st = MRS_ERR_NOMEM;
  if (it && sc) st = mrs_recommend(H(mrs_model, m), H(mrs_sim, sim), predKind, user, n, (int32_t*)it, (double*)sc, &w);
  if (sc) (*env)->ReleasePrimitiveArrayCritical(env, scoresOut, sc, 0);
  if (it) (*env)->ReleasePrimitiveArrayCritical(env, itemsOut, it, 0);
  if (st != MRS_OK) { throw_last(env); return 0; }
  return w;
}

/* ------------------------------------------------------------------ several GPUs from this one JVM (distributed/DistributedBaseline.scala:30-47) */
JNIEXPORT jlong JNICALL Java_shared_NativeEngine_multiCreate(JNIEnv* env, jclass c, jintArray deviceIds) {
  const jsize n = (*env)->GetArrayLength(env, deviceIds);
  jint* d = (jint*)(*env)->GetPrimitiveArrayCritical(env, deviceIds, NULL);
  mrs_multi* out = NULL;
  int32_t st = MRS_ERR_NOMEM;
  if (d) st = mrs_multi_create((const int32_t*)d, (int32_t)n, &out);
  if (d) (*env)->ReleasePrimitiveArrayCritical(env, deviceIds, d, JNI_ABORT);
  if (st != MRS_OK) { throw_last(env); return 0; }
  return J(out);
}
JNIEXPORT void JNICALL Java_shared_NativeEngine_multiLoad(JNIEnv* env, jclass c, jlong multi, jintArray trU, jintArray trI, jdoubleArray trR,
                                                          jintArray teU, jintArray teI, jdoubleArray teR) {
  const jsize n = (*env)->GetArrayLength(env, trU), nt = (*env)->GetArrayLength(env, teU);
  jint* a = (jint*)(*env)->GetPrimitiveArrayCritical(env, trU, NULL);
  jint* b = (jint*)(*env)->GetPrimitiveArrayCritical(env, trI, NULL);
  jdouble* r = (jdouble*)(*env)->GetPrimitiveArrayCritical(env, trR, NULL);
  jint* x = (jint*)(*env)->GetPrimitiveArrayCritical(env, teU, NULL);
  jint* y = (jint*)(*env)->GetPrimitiveArrayCritical(env, teI, NULL);
  jdouble* z = (jdouble*)(*env)->GetPrimitiveArrayCritical(env, teR, NULL);
  int32_t st = MRS_ERR_NOMEM;
  if (a && b && r && x && y && z)
    st = mrs_multi_load(H(mrs_multi, multi), (const int32_t*)a, (const int32_t*)b, (const double*)r, (int64_t)n, (const int32_t*)x, (const int32_t*)y,
                        (const double*)z, (int64_t)nt);
  if (z) (*env)->ReleasePrimitiveArrayCritical(env, teR, z, JNI_ABORT);
  if (y) (*env)->ReleasePrimitiveArrayCritical(env, teI, y, JNI_ABORT);
  if (x) (*env)->ReleasePrimitiveArrayCritical(env, teU, x, JNI_ABORT);
  if (r) (*env)->ReleasePrimitiveArrayCritical(env, trR, r, JNI_ABORT);
  if (b) (*env)->ReleasePrimitiveArrayCritical(env, trI, b, JNI_ABORT);
  if (a) (*env)->ReleasePrimitiveArrayCritical(env, trU, a, JNI_ABORT);
  if (st != MRS_OK) throw_last(env);
}
JNIEXPORT jdouble JNICALL Java_shared_NativeEngine_multiBaselineMae(JNIEnv* env, jclass c, jlong multi) {
  double out = 0.0;
  CHECK(mrs_multi_baseline_mae(H(mrs_multi, multi), &out), 0.0);
  return out;
}
JNIEXPORT jlong JNICALL Java_shared_NativeEngine_multiModel(JNIEnv* env, jclass c, jlong multi, jint slot) {
  mrs_model* m = NULL;
  CHECK(mrs_multi_model(H(mrs_multi, multi), slot, &m), 0);
  return J(m);
}
JNIEXPORT jint JNICALL Java_shared_NativeEngine_multiOwner(JNIEnv* env, jclass c, jlong multi, jint user) {
  int32_t slot = 0;
  CHECK(mrs_multi_owner(H(mrs_multi, multi), user, &slot), 0);
  return slot;
}
JNIEXPORT void JNICALL Java_shared_NativeEngine_multiDestroy(JNIEnv* env, jclass c, jlong multi) { mrs_multi_destroy(H(mrs_multi, multi)); }
