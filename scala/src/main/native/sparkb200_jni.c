/*
 * sparkb200_jni.c -- mechanical JNI shim between org.apache.spark.sql.b200.Native and libsparkb200.so.
 * Every function: unpack Java arrays, call the sb_* entry point, throw B200Exception on a non-zero code.
 * Build (needs a JDK for jni.h; none in this image):
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../../../include \
 *       sparkb200_jni.c -L../../../../spark_b200 -lsparkb200 -o libsparkb200_jni.so
 */
#include <jni.h>
#include <stdlib.h>
#include "spark_b200.h"

static void throw_if(JNIEnv *env, int rc) {
  if (rc == SB_OK) return;
  jclass cls = (*env)->FindClass(env, "org/apache/spark/sql/b200/B200Exception");
  (*env)->ThrowNew(env, cls, sb_last_error());   /* task fails -> Spark's retry policy applies */
}

JNIEXPORT void JNICALL Java_org_apache_spark_sql_b200_Native_init(JNIEnv *env, jclass c, jint dev) {
  throw_if(env, sb_init(dev));
}

JNIEXPORT jlong JNICALL Java_org_apache_spark_sql_b200_Native_streamCreate(JNIEnv *env, jclass c) {
  sb_stream *s = NULL;
  throw_if(env, sb_stream_create(&s));
  return (jlong)(intptr_t)s;
}

JNIEXPORT jlong JNICALL Java_org_apache_spark_sql_b200_Native_tableImportHost(
    JNIEnv *env, jclass c, jintArray types, jlongArray lengths, jlongArray nulls, jlongArray data, jlongArray validity,
    jlongArray offsets, jlong stream) {
  jsize n = (*env)->GetArrayLength(env, types);
  jint *t = (*env)->GetIntArrayElements(env, types, NULL);
  jlong *len = (*env)->GetLongArrayElements(env, lengths, NULL), *nc = (*env)->GetLongArrayElements(env, nulls, NULL);
  jlong *d = (*env)->GetLongArrayElements(env, data, NULL), *v = (*env)->GetLongArrayElements(env, validity, NULL);
  jlong *o = (*env)->GetLongArrayElements(env, offsets, NULL);
  sb_column *cols = (sb_column *)calloc((size_t)n, sizeof(sb_column));
  for (jsize i = 0; i < n; i++) {
    cols[i].type = t[i]; cols[i].length = len[i]; cols[i].null_count = nc[i];
    cols[i].data = (const void *)(intptr_t)d[i];            /* OffHeapColumnVector / ArrowBuf address */
    cols[i].validity = (const uint8_t *)(intptr_t)v[i];
    cols[i].offsets = (const int32_t *)(intptr_t)o[i];
  }
  sb_table *out = NULL;
  int rc = sb_table_import_host(cols, n, (sb_stream *)(intptr_t)stream, &out);
  free(cols);
  (*env)->ReleaseIntArrayElements(env, types, t, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, lengths, len, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, nulls, nc, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, data, d, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, validity, v, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, offsets, o, JNI_ABORT);
  throw_if(env, rc);
  return (jlong)(intptr_t)out;
}

JNIEXPORT void JNICALL Java_org_apache_spark_sql_b200_Native_tableRelease(JNIEnv *env, jclass c, jlong t) {
  throw_if(env, sb_table_release((sb_table *)(intptr_t)t));
}

JNIEXPORT jlong JNICALL Java_org_apache_spark_sql_b200_Native_hashPartition(JNIEnv *env, jclass c, jlong table, jintArray keyCols,
                                                                            jint n, jlong stream, jlongArray offsetsOut) {
  jsize nk = (*env)->GetArrayLength(env, keyCols);
  jint *k = (*env)->GetIntArrayElements(env, keyCols, NULL);
  jlong *offs = (*env)->GetLongArrayElements(env, offsetsOut, NULL);
  sb_table *out = NULL;
  int rc = sb_hash_partition((const sb_table *)(intptr_t)table, (const int32_t *)k, nk, n, (sb_stream *)(intptr_t)stream, &out,
                             (int64_t *)offs);
  (*env)->ReleaseIntArrayElements(env, keyCols, k, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, offsetsOut, offs, 0);
  throw_if(env, rc);
  return (jlong)(intptr_t)out;
}

JNIEXPORT jlong JNICALL Java_org_apache_spark_sql_b200_Native_hashAggregate(
    JNIEnv *env, jclass c, jlong table, jint mode, jintArray keyCols, jintArray funcs, jlongArray inputExprs, jlong filterExpr,
    jlong expectedGroups, jlong stream) {
  jsize nk = (*env)->GetArrayLength(env, keyCols), na = (*env)->GetArrayLength(env, funcs);
  jint *k = (*env)->GetIntArrayElements(env, keyCols, NULL), *f = (*env)->GetIntArrayElements(env, funcs, NULL);
  jlong *in = (*env)->GetLongArrayElements(env, inputExprs, NULL);
  sb_agg_spec *specs = (sb_agg_spec *)calloc((size_t)(na ? na : 1), sizeof(sb_agg_spec));
  for (jsize i = 0; i < na; i++) {
    specs[i].func = f[i];
    if (in[i]) specs[i].input = *(const sb_expr *)(intptr_t)in[i];
  }
  sb_agg_plan plan = {mode, nk, (const int32_t *)k, na, 0, specs, (const sb_expr *)(intptr_t)filterExpr, expectedGroups};
  sb_table *out = NULL;
  int rc = sb_hash_aggregate((const sb_table *)(intptr_t)table, &plan, (sb_stream *)(intptr_t)stream, &out);
  free(specs);
  (*env)->ReleaseIntArrayElements(env, keyCols, k, JNI_ABORT);
  (*env)->ReleaseIntArrayElements(env, funcs, f, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, inputExprs, in, JNI_ABORT);
  throw_if(env, rc);
  return (jlong)(intptr_t)out;
}

JNIEXPORT jlong JNICALL Java_org_apache_spark_sql_b200_Native_joinBuild(JNIEnv *env, jclass c, jlong table, jintArray keyCols, jlong stream) {
  jsize nk = (*env)->GetArrayLength(env, keyCols);
  jint *k = (*env)->GetIntArrayElements(env, keyCols, NULL);
  sb_hash_table *ht = NULL;
  int rc = sb_join_build((const sb_table *)(intptr_t)table, (const int32_t *)k, nk, (sb_stream *)(intptr_t)stream, &ht);
  (*env)->ReleaseIntArrayElements(env, keyCols, k, JNI_ABORT);
  throw_if(env, rc);
  return (jlong)(intptr_t)ht;
}

JNIEXPORT jlong JNICALL Java_org_apache_spark_sql_b200_Native_joinProbe(JNIEnv *env, jclass c, jlong rel, jlong probe, jintArray keyCols,
                                                                        jint joinType, jlong stream) {
  jsize nk = (*env)->GetArrayLength(env, keyCols);
  jint *k = (*env)->GetIntArrayElements(env, keyCols, NULL);
  sb_table *out = NULL;
  int rc = sb_join_probe((const sb_hash_table *)(intptr_t)rel, (const sb_table *)(intptr_t)probe, (const int32_t *)k, nk, joinType,
                         (sb_stream *)(intptr_t)stream, &out);
  (*env)->ReleaseIntArrayElements(env, keyCols, k, JNI_ABORT);
  throw_if(env, rc);
  return (jlong)(intptr_t)out;
}
/* sort / topN / filterProject / allToAll / allGather / expr* follow the same pattern (arrays in, one sb_* call, throw_if). */
