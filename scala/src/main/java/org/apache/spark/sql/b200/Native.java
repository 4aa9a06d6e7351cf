package org.apache.spark.sql.b200;

/**
 * JNI face of libsparkb200.so (include/spark_b200.h).  One static native method per C entry point the
 * operators use; handles are jlong, errors surface as B200Exception (non-zero return + sb_last_error()).
 * Compiled only where a JDK exists -- this image has none (see DESIGN.md); the C side is exercised through
 * ctypes with identical signatures.
 */
public final class Native {
  static { System.loadLibrary("sparkb200_jni"); }   // links against libsparkb200.so
  private Native() {}

  public static native void init(int deviceOrdinal);
  public static native void shutdown();
  public static native byte[] commGetUniqueId();
  public static native void commInit(int rank, int nranks, byte[] uniqueId);
  public static native long streamCreate();
  public static native void streamDestroy(long stream);

  /** columns: parallel arrays describing Arrow buffers (type, length, nullCount, data/validity/offsets addresses). */
  public static native long tableImportHost(int[] types, long[] lengths, long[] nullCounts, long[] data, long[] validity,
                                            long[] offsets, long stream);
  public static native long tableNumRows(long table);
  public static native void tableExportHost(long table, int column, long data, long validity, long offsets, long stream);
  public static native void tableRelease(long table);

  public static native long filterProject(long table, long predicateExpr, long[] projectionExprs, long stream);
  /** returns the partitioned table; offsetsOut receives numPartitions + 1 boundaries. */
  public static native long hashPartition(long table, int[] keyCols, int numPartitions, long stream, long[] offsetsOut);
  public static native long roundRobinPartition(long table, int start, int numPartitions, long stream, long[] offsetsOut);
  public static native long hashAggregate(long table, int mode, int[] keyCols, int[] funcs, long[] inputExprs, long filterExpr,
                                          long expectedGroups, long stream);
  public static native long sort(long table, int[] cols, boolean[] ascending, boolean[] nullsFirst, long stream);
  public static native long topN(long table, int[] cols, boolean[] ascending, boolean[] nullsFirst, long k, long stream);
  public static native long joinBuild(long table, int[] keyCols, long stream);
  public static native long joinProbe(long relation, long probe, int[] keyCols, int joinType, long stream);
  public static native void hashTableRelease(long relation);
  public static native long allToAll(long table, long[] partOffsets, int numPartitions, long stream, long[] outPartOffsets);
  public static native long allGather(long table, long stream);

  /** postfix sb_expr program built by ExprCompiler; returns a native handle freed with exprFree. */
  public static native long exprCreate(int[] ops, int[] vtypes, int[] args, long[] literals, int outType);
  public static native void exprFree(long expr);
}
