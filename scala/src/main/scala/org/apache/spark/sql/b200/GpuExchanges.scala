package org.apache.spark.sql.b200

import java.util.UUID
import java.util.concurrent.ConcurrentHashMap

import scala.concurrent.{ExecutionContext, Future}

import org.apache.spark.{MapOutputStatistics, Partition, TaskContext}
import org.apache.spark.broadcast.Broadcast
import org.apache.spark.rdd.RDD
import org.apache.spark.sql.catalyst.InternalRow
import org.apache.spark.sql.catalyst.plans.logical.Statistics
import org.apache.spark.sql.catalyst.plans.physical._
import org.apache.spark.sql.execution._
import org.apache.spark.sql.execution.exchange.{BroadcastExchangeLike, ShuffleExchangeLike, ShuffleOrigin}
import org.apache.spark.sql.execution.metric.{SQLMetric, SQLMetrics}
import org.apache.spark.sql.types.DataType
import org.apache.spark.sql.vectorized.ColumnarBatch

/**
 * Executor-local registry of exchange outputs that stay in HBM between the stage that produced them and the stage that reads
 * them (the role shuffle files + the block manager play for the reference: SortShuffleManager / BlockStoreShuffleReader).
 * Entry: the rows this executor received, grouped by source rank and partition-contiguous inside every source block, plus the
 * boundaries needed to cut out one reducer partition.  Freed by B200Plugin when the shuffle is unregistered / the app ends.
 */
object DeviceShuffleStore {
  final case class Entry(table: Long, types: Array[DataType], firstPartition: Int, lastPartition: Int,
                         sourceBlockOffsets: Array[Array[Long]])      // [source rank][owned partitions + 1]: row offsets inside `table`
  private val entries = new ConcurrentHashMap[Int, Entry]()
  def put(shuffleId: Int, e: Entry): Unit = Option(entries.put(shuffleId, e)).foreach(old => Native.tableRelease(old.table))
  def get(shuffleId: Int): Entry = Option(entries.get(shuffleId)).getOrElse(
    throw new B200Exception(2, s"device shuffle $shuffleId is not resident on this executor (lost executor: the query must be retried, " +
      "like the reference's pipelined shuffle, ShuffleExchangeExec.scala:558-566)"))
  def remove(shuffleId: Int): Unit = Option(entries.remove(shuffleId)).foreach(e => Native.tableRelease(e.table))
  def clear(): Unit = { val it = entries.keySet().iterator(); while (it.hasNext) remove(it.next()) }
}

/**
 * ShuffleExchangeExec on the GPUs (SQLX/exchange/ShuffleExchangeExec.scala:190).  It must stay a ShuffleExchangeLike or AQE
 * refuses the plan (AdaptiveSparkPlanExec.scala:977-988); every abstract member of the trait (:55-151) is implemented.
 *
 * Materialisation = ONE barrier stage with one task per GPU executor (RDD.barrier: all tasks run together, which a collective
 * needs): each task partitions the rows of its input partitions on its GPU (sb_hash_partition / sb_round_robin_partition /
 * sb_range_partition: Murmur3 pmod ids bit-identical to HashPartitioning.partitionIdExpression, stable regrouping), the buckets
 * cross NVLink with sb_all_to_all (rank r owns the contiguous partition range [ceil(r n / R), ceil((r + 1) n / R))), and the
 * received table stays in HBM in DeviceShuffleStore.  The task reports the bytes per reducer partition (sb_map_output_statistics:
 * MapOutputStatistics.bytesByPartitionId over ALL map sides), which is what mapOutputStatisticsFuture completes with.
 *
 * getShuffleRDD(specs) has one partition per spec (AQEShuffleReadExec.scala:268-284 reads it as RDD[ColumnarBatch]); a
 * CoalescedPartitionSpec is served from the store of the executor that owns its reducer range.  Limitation (stated in
 * INTEGRATION.md): a coalesced range that straddles two executors' ownership ranges is rejected -- serving it needs a
 * one-sided peer read of the other executor's table, which the C ABI does not expose yet; PartialReducer / PartialMapper
 * specs (skew join splitting) are rejected for the same reason, so GpuSupport only takes an exchange when
 * spark.sql.adaptive.skewJoin.enabled is false.
 */
case class GpuShuffleExchangeExec(
    override val outputPartitioning: Partitioning,
    child: SparkPlan,
    shuffleOrigin: ShuffleOrigin,
    advisoryPartitionSize: Option[Long]) extends ShuffleExchangeLike with GpuExec {

  override lazy val metrics: Map[String, SQLMetric] = Map(
    "dataSize" -> SQLMetrics.createSizeMetric(sparkContext, "data size"),
    "numPartitions" -> SQLMetrics.createMetric(sparkContext, "number of partitions"))

  /** one executor per GPU: spark.executor.instances executors each holding one gpu resource (B200Plugin checks this at start-up) */
  private def numGpus: Int = B200Plugin.numExecutors(sparkContext.getConf)

  override def numMappers: Int = numGpus
  override def numPartitions: Int = outputPartitioning.numPartitions
  override lazy val shuffleId: Int = sparkContext.newShuffleId()

  private lazy val childTypes: Array[DataType] = child.output.map(_.dataType).toArray

  /** The barrier stage; returns per task (rank, bytesByPartitionId as seen by every rank, rows received). */
  @transient private lazy val materialised: Future[Array[(Int, Array[Long], Long)]] = {
    val n = numPartitions
    val types = childTypes
    val sid = shuffleId
    val part = GpuSupport.describePartitioning(outputPartitioning, child.output)      // serialisable: kind + key ordinals / bounds
    val input = child.executeColumnar().coalesce(numGpus, shuffle = false)
    val stage = input.barrier().mapPartitions { batches =>
      val ctx = org.apache.spark.BarrierTaskContext.get()
      val stream = GpuExec.taskStream()
      val rank = ctx.partitionId()
      val in = GpuSupport.concatToDevice(batches, types, stream, orEmpty = true)
      val offs = new Array[Long](n + 1)
      val parted = try part.apply(in.table, n, stream, offs) finally in.close()        // sb_*_partition
      try {
        val world = ctx.getTaskInfos().length
        val bytes = Native.mapOutputStatistics(parted, offs, n, stream)                // collective: all-gather of the counts
        val counts = Native.exchangeCounts(offs, n, world, stream)                     // rows of every (source rank, partition)
        val recvOffs = new Array[Long](n + 1)
        val received = Native.allToAll(parted, offs, n, stream, recvOffs)              // collective: NVLink all-to-all
        val lo = (rank.toLong * n + world - 1) / world
        val hi = ((rank + 1).toLong * n + world - 1) / world
        DeviceShuffleStore.put(sid, DeviceShuffleStore.Entry(received, types, lo.toInt, hi.toInt,
          GpuSupport.sourceBlockOffsets(counts, n, lo.toInt, hi.toInt, world)))
        Iterator.single((rank, bytes, Native.tableNumRows(received)))
      } finally Native.tableRelease(parted)
    }
    // a plain job on a helper thread: AQE only needs the Future (ShuffleExchangeLike.submitShuffleJob, :118-121)
    Future(stage.collect())(ExecutionContext.global)
  }

  override protected def mapOutputStatisticsFuture: Future[MapOutputStatistics] =
    materialised.map { results =>
      val bytes = results.head._2                                                       // every rank computed the same global vector
      longMetric("dataSize") += bytes.sum
      longMetric("numPartitions") += numPartitions
      new MapOutputStatistics(shuffleId, bytes)
    }(ExecutionContext.global)

  override def runtimeStatistics: Statistics = {
    val results = scala.concurrent.Await.result(materialised, scala.concurrent.duration.Duration.Inf)
    Statistics(sizeInBytes = BigInt(results.head._2.sum), rowCount = Some(BigInt(results.map(_._3).sum)), isRuntime = true)
  }

  override def getShuffleRDD(partitionSpecs: Array[ShufflePartitionSpec]): RDD[_] = {
    scala.concurrent.Await.result(materialised, scala.concurrent.duration.Duration.Inf)
    new DeviceShuffleReadRDD(sparkContext, shuffleId, numPartitions, numGpus, partitionSpecs)
  }

  override protected def doExecuteColumnar(): RDD[ColumnarBatch] = {
    // without AQE: one output partition per reducer partition, exactly as HashPartitioning(n) promises (ADVICE round 1: the sketch
    // returned one batch per executor holding a RANGE of partition ids while still advertising HashPartitioning(n))
    getShuffleRDD(Array.tabulate[ShufflePartitionSpec](numPartitions)(i => CoalescedPartitionSpec(i, i + 1)))
      .asInstanceOf[RDD[ColumnarBatch]]
  }
  override protected def withNewChildInternal(c: SparkPlan): SparkPlan = copy(child = c)
}

/** One partition per ShufflePartitionSpec; runs on the executor that owns the spec's reducer range and slices its stored table. */
class DeviceShuffleReadRDD(
    sc: org.apache.spark.SparkContext, shuffleId: Int, numReducers: Int, numGpus: Int, specs: Array[ShufflePartitionSpec])
  extends RDD[ColumnarBatch](sc, Nil) {
  private case class SpecPartition(index: Int, spec: ShufflePartitionSpec) extends Partition
  private def ownerOf(reducer: Int): Int = {      // inverse of the contiguous ownership of sb_all_to_all / sb_exchange_plan
    var r = 0
    while (r + 1 < numGpus && ((r + 1).toLong * numReducers + numGpus - 1) / numGpus <= reducer) r += 1
    r
  }
  override protected def getPartitions: Array[Partition] = specs.zipWithIndex.map { case (s, i) => SpecPartition(i, s): Partition }
  override protected def getPreferredLocations(split: Partition): Seq[String] = split.asInstanceOf[SpecPartition].spec match {
    case CoalescedPartitionSpec(start, _, _) => B200Plugin.executorLocation(ownerOf(start)).toSeq
    case _ => Nil
  }
  override def compute(split: Partition, context: TaskContext): Iterator[ColumnarBatch] = split.asInstanceOf[SpecPartition].spec match {
    case CoalescedPartitionSpec(start, end, _) =>
      val e = DeviceShuffleStore.get(shuffleId)
      if (start < e.firstPartition || end > e.lastPartition)
        throw new B200Exception(5, s"reducer range [$start, $end) is not owned by this executor ([${e.firstPartition}, ${e.lastPartition})): " +
          "coalesced ranges that straddle executors need a peer read (see INTEGRATION.md)")
      val stream = GpuExec.taskStream()
      // the range's rows are one slice per source rank; stitch them into one batch (sb_table_slice + sb_table_concat)
      val slices = e.sourceBlockOffsets.map { offs =>
        Native.tableSlice(e.table, offs(start - e.firstPartition), offs(end - e.firstPartition), stream)
      }
      try Iterator.single(new DeviceBatch(Native.tableConcat(slices, stream), e.types): ColumnarBatch)
      finally slices.foreach(Native.tableRelease)
    case other => throw new B200Exception(5, s"shuffle partition spec $other is not supported by the device shuffle")
  }
}

/**
 * BroadcastExchangeExec on the GPUs (SQLX/exchange/BroadcastExchangeExec.scala:124): every executor ends up with the whole
 * relation in its own HBM (sb_all_gather over NVLink when the child is spread over the GPUs); what is "broadcast" through Spark
 * is only the shuffle-store key, the rows never visit the driver (the reference collects them to the driver, :177-260).
 * Implements every member of BroadcastExchangeLike (:45-90).
 */
case class GpuBroadcastExchangeExec(mode: BroadcastMode, child: SparkPlan) extends BroadcastExchangeLike with GpuExec {
  override val runId: UUID = UUID.randomUUID
  override lazy val metrics: Map[String, SQLMetric] = Map(
    "dataSize" -> SQLMetrics.createSizeMetric(sparkContext, "data size"),
    "numOutputRows" -> SQLMetrics.createMetric(sparkContext, "number of output rows"))
  override def outputPartitioning: Partitioning = BroadcastPartitioning(mode)
  private lazy val storeKey: Int = sparkContext.newShuffleId()
  private lazy val types: Array[DataType] = child.output.map(_.dataType).toArray

  @transient private lazy val gathered: java.util.concurrent.Future[Broadcast[Any]] = {
    val sid = storeKey
    val tps = types
    val numGpus = B200Plugin.numExecutors(sparkContext.getConf)
    val task = new java.util.concurrent.FutureTask[Broadcast[Any]](() => {
      val rows = child.executeColumnar().coalesce(numGpus, shuffle = false).barrier().mapPartitions { batches =>
        val stream = GpuExec.taskStream()
        val in = GpuSupport.concatToDevice(batches, tps, stream, orEmpty = true)
        val all = try Native.allGather(in.table, stream) finally in.close()              // collective: every rank gets every row
        DeviceShuffleStore.put(sid, DeviceShuffleStore.Entry(all, tps, 0, 1, Array(Array(0L, Native.tableNumRows(all)))))
        Iterator.single(Native.tableNumRows(all))
      }.collect()
      longMetric("numOutputRows") += rows.headOption.getOrElse(0L)
      sparkContext.broadcast[Any](Array(sid.toLong))                                      // the handle every executor resolves locally
    })
    BroadcastExchangeLikeThreads.pool.execute(task)
    task
  }
  override def relationFuture: java.util.concurrent.Future[Broadcast[Any]] = gathered
  override protected def completionFuture: Future[Broadcast[Any]] = Future(gathered.get())(ExecutionContext.global)
  override def runtimeStatistics: Statistics = {
    gathered.get()
    Statistics(sizeInBytes = BigInt(metrics("dataSize").value), rowCount = Some(BigInt(metrics("numOutputRows").value)), isRuntime = true)
  }
  override protected def doPrepare(): Unit = { gathered }
  override protected def doExecuteBroadcast[T](): Broadcast[T] = gathered.get().asInstanceOf[Broadcast[T]]
  override protected def doExecuteColumnar(): RDD[ColumnarBatch] =
    throw new IllegalStateException("a broadcast exchange is consumed through executeBroadcast")
  override protected def withNewChildInternal(c: SparkPlan): SparkPlan = copy(child = c)
}

object GpuBroadcastExchangeExec {
  /** the relation as this executor holds it (a retained view of the stored table; the caller closes the batch) */
  def batchesOf(handle: Array[Long], types: Array[DataType]): Iterator[ColumnarBatch] = {
    val e = DeviceShuffleStore.get(handle(0).toInt)
    Native.tableRetain(e.table)
    Iterator.single(new DeviceBatch(e.table, types): ColumnarBatch)
  }
}

object BroadcastExchangeLikeThreads {
  lazy val pool: java.util.concurrent.ExecutorService = java.util.concurrent.Executors.newCachedThreadPool()
}
