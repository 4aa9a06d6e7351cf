package org.apache.spark.sql.b200

import org.apache.spark.TaskContext
import org.apache.spark.rdd.RDD
import org.apache.spark.sql.catalyst.InternalRow
import org.apache.spark.sql.execution.SparkPlan
import org.apache.spark.sql.execution.vectorized.OffHeapColumnVector
import org.apache.spark.sql.types._
import org.apache.spark.sql.vectorized.{ColumnarArray, ColumnarBatch, ColumnarMap, ColumnVector}
import org.apache.spark.unsafe.Platform
import org.apache.spark.unsafe.types.UTF8String

/** Common contract of the GPU operators (SparkPlan.scala:92 supportsColumnar, :232 executeColumnar, :359 doExecuteColumnar). */
trait GpuExec extends SparkPlan {
  override def supportsColumnar: Boolean = true
  override def supportsRowBased: Boolean = false
  override protected def doExecute(): RDD[InternalRow] =
    throw new IllegalStateException(s"$nodeName is columnar only: there is no CPU fallback")

  /** One CUDA stream per task thread (sb_stream); destroyed when the task completes, with everything the task still owns. */
  protected def taskStream(): Long = GpuExec.taskStream()
}

object GpuExec {
  private val streams = new ThreadLocal[java.lang.Long]

  def taskStream(): Long = {
    val cur = streams.get()
    if (cur != null) return cur
    val s = Native.streamCreate()
    streams.set(s)
    Option(TaskContext.get()).foreach(_.addTaskCompletionListener[Unit] { _ =>   // TaskContext.scala:130-141
      Native.streamDestroy(s)
      streams.remove()
    })
    s
  }

  /** sb_* type id of a Catalyst type, or -1 (widths follow the row -> column converters, Columnar.scala:290-326). */
  def typeId(dt: DataType): Int = dt match {
    case BooleanType => 1
    case ByteType => 2
    case ShortType => 3
    case IntegerType => 4
    case LongType => 5
    case FloatType => 6
    case DoubleType => 7
    case DateType => 8
    case TimestampType | TimestampNTZType => 9
    case d: DecimalType if d.precision <= 18 => 10      // unscaled long (Decimal.scala: compact form), SB_DECIMAL64
    case _: DecimalType => 12                           // 128-bit unscaled value, SB_DECIMAL128: results of SUM / AVG, payload only
    case StringType | BinaryType => 11
    case _ => -1
  }
}

/** A ColumnarBatch whose columns live in HBM behind one sb_table handle.  Whoever creates a batch closes it
 *  (SparkPlan.scala:355-358): close() == sb_table_release. */
final class DeviceBatch(val table: Long, val types: Array[DataType])
  extends ColumnarBatch(
    types.indices.map(i => new DeviceColumnVector(table, i, types(i)): ColumnVector).toArray,
    Native.tableNumRows(table).toInt) {
  private var open = true
  override def close(): Unit = if (open) { open = false; Native.tableRelease(table) }
}

/**
 * ColumnVector view of one column of a device table.  Operators of this plugin never read it element-wise: they hand the
 * table handle to the next sb_* call.  A CPU operator that does read it (only possible after DeviceToHostExec, which turns the
 * batch into OffHeapColumnVectors) never sees this class; the getters below therefore fail loudly instead of copying
 * element by element across PCIe.
 */
final class DeviceColumnVector(val table: Long, val ordinal: Int, dt: DataType) extends ColumnVector(dt) {
  private def hostOnly(): Nothing =
    throw new UnsupportedOperationException("device-resident column: insert DeviceToHostExec (ColumnarToRow) before reading rows")
  override def close(): Unit = ()                      // the DeviceBatch owns the table
  override def hasNull: Boolean = Native.columnNullCount(table, ordinal) != 0
  override def numNulls: Int = Native.columnNullCount(table, ordinal).toInt
  override def isNullAt(rowId: Int): Boolean = hostOnly()
  override def getBoolean(rowId: Int): Boolean = hostOnly()
  override def getByte(rowId: Int): Byte = hostOnly()
  override def getShort(rowId: Int): Short = hostOnly()
  override def getInt(rowId: Int): Int = hostOnly()
  override def getLong(rowId: Int): Long = hostOnly()
  override def getFloat(rowId: Int): Float = hostOnly()
  override def getDouble(rowId: Int): Double = hostOnly()
  override def getArray(rowId: Int): ColumnarArray = hostOnly()
  override def getMap(ordinal: Int): ColumnarMap = hostOnly()
  override def getDecimal(rowId: Int, precision: Int, scale: Int): Decimal = hostOnly()
  override def getUTF8String(rowId: Int): UTF8String = hostOnly()
  override def getBinary(rowId: Int): Array[Byte] = hostOnly()
  override def getChild(ordinal: Int): ColumnVector = hostOnly()
}

/** Host <-> device movement of whole batches (the bodies of HostToDeviceExec / DeviceToHostExec). */
object DeviceTransfer {
  /**
   * OffHeapColumnVector keeps nulls as ONE BYTE PER ROW (1 = NULL, OffHeapColumnVector.java:67-76) while the C ABI takes an Arrow
   * validity BITMAP (1 = valid, LSB first); the bytes are packed into a pinned scratch bitmap here (ADVICE round 1: passing the
   * byte array as a bitmap was wrong).  Values / offsets are passed by address: the off-heap buffers are plain native memory.
   */
  def toDevice(batch: ColumnarBatch, types: Array[DataType], stream: Long): DeviceBatch = batch match {
    case d: DeviceBatch => d
    case _ =>
      val n = batch.numRows()
      val ncols = batch.numCols()
      val typeIds = new Array[Int](ncols)
      val scales = types.map { case d: DecimalType => (d.precision << 8) | d.scale; case _ => 0 }
      val lengths = Array.fill[Long](ncols)(n.toLong)
      val nullCounts = new Array[Long](ncols)
      val data = new Array[Long](ncols)
      val validity = new Array[Long](ncols)
      val offsets = new Array[Long](ncols)
      val scratch = new scala.collection.mutable.ArrayBuffer[Long]()
      try {
        var c = 0
        while (c < ncols) {
          val v = batch.column(c) match {
            case o: OffHeapColumnVector => o
            case other => throw new B200Exception(5, s"column $c is a ${other.getClass.getName}; the GPU path imports OffHeapColumnVector " +
              "batches (spark.sql.columnVector.offheap.enabled=true) or ArrowColumnVector buffers")
          }
          typeIds(c) = GpuExec.typeId(types(c))
          if (typeIds(c) < 0 || typeIds(c) == 12)    // OffHeapColumnVector keeps decimal(p > 18) as byte arrays, not as 16-byte values
            throw new B200Exception(5, s"type ${types(c)} is not supported as an input of the GPU path")
          nullCounts(c) = v.numNulls()
          data(c) = v.valuesNativeAddress()
          if (v.hasNull) {
            val bitmapBytes = (n + 7) / 8 + 8
            val bm = Native.hostAlloc(bitmapBytes)
            scratch += bm
            Platform.setMemory(bm, 0.toByte, bitmapBytes)
            var r = 0
            while (r < n) {
              if (!v.isNullAt(r)) {   // the nulls buffer has no public address (OffHeapColumnVector.java:70): one byte per row, 1 = NULL
                val addr = bm + (r >> 3)
                Platform.putByte(null, addr, (Platform.getByte(null, addr) | (1 << (r & 7))).toByte)
              }
              r += 1
            }
            validity(c) = bm
          }
          c += 1
        }
        val table = Native.tableImportHost(typeIds, scales, lengths, nullCounts, data, validity, offsets, stream)
        Native.streamSynchronize(stream)          // the scratch bitmaps and the source vectors may go away after this call
        new DeviceBatch(table, types)
      } finally scratch.foreach(Native.hostFree)
  }

  /** DeviceBatch -> OffHeapColumnVector batch (ColumnarToRowExec reads those row by row). */
  def toHost(batch: ColumnarBatch, types: Array[DataType], stream: Long): ColumnarBatch = batch match {
    case d: DeviceBatch =>
      val n = d.numRows()
      val vectors = OffHeapColumnVector.allocateColumns(math.max(n, 1), StructType(types.zipWithIndex.map { case (t, i) =>
        StructField(s"c$i", t) }))
      var c = 0
      while (c < types.length) {
        val v = vectors(c)
        val bitmapBytes = (n + 7) / 8 + 8
        val bm = Native.hostAlloc(bitmapBytes)
        try {
          val wide = types(c) match { case dt: DecimalType if dt.precision > 18 => true; case _ => false }
          val staging = if (wide) Native.hostAlloc(16L * math.max(n, 1)) else 0L   // decimal(p > 18): 16-byte little-endian values
          val nulls = Native.tableExportHost(d.table, c, if (wide) staging else v.valuesNativeAddress(), bm, 0L, stream)   // returns the NULL count
          if (wide) {
            val dt = types(c).asInstanceOf[DecimalType]
            val bytes = new Array[Byte](16)
            var r = 0
            while (r < n) {
              var k = 0
              while (k < 16) { bytes(15 - k) = Platform.getByte(null, staging + 16L * r + k); k += 1 }   // BigInteger wants big endian
              v.putDecimal(r, Decimal(new java.math.BigDecimal(new java.math.BigInteger(bytes), dt.scale), dt.precision, dt.scale), dt.precision)
              r += 1
            }
            Native.hostFree(staging)
          }
          if (nulls > 0) {
            var r = 0
            while (r < n) {
              if (((Platform.getByte(null, bm + (r >> 3)) >> (r & 7)) & 1) == 0) v.putNull(r)
              r += 1
            }
          }
        } finally Native.hostFree(bm)
        c += 1
      }
      val out = new ColumnarBatch(vectors.map(_.asInstanceOf[ColumnVector]), n)
      d.close()
      out
    case other => other
  }
}
