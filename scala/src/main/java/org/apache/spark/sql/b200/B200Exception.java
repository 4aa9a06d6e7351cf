package org.apache.spark.sql.b200;

/** Thrown by the JNI shim when an sb_* call returns non-zero: the task fails and Spark's retry policy applies
 *  (the reference's operators throw QueryExecutionErrors / SparkException the same way). */
public class B200Exception extends RuntimeException {
  private final int code;

  public B200Exception(String message) {          // the constructor ThrowNew uses: message = sb_last_error()
    super(message);
    this.code = -1;
  }

  public B200Exception(int code, String message) {
    super(message);
    this.code = code;
  }

  public int code() { return code; }
}
