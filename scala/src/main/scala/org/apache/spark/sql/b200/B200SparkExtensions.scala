/*
 * Scala host side of the drop-in: enabled with
 *   --conf spark.sql.extensions=org.apache.spark.sql.b200.B200SparkExtensions
 *   --conf spark.plugins=org.apache.spark.sql.b200.B200Plugin
 *   --conf spark.executor.resource.gpu.amount=1 --conf spark.task.resource.gpu.amount=1
 * Catalyst, the DataFrame API and the SparkPlan operator surface are untouched: the rule only swaps physical
 * operators for GPU ones that keep the reference operator's output / partitioning / ordering contracts.
 * Written against the reference at /root/reference (5.0.0-SNAPSHOT); not compiled in this image (no JDK).
 */
package org.apache.spark.sql.b200

import org.apache.spark.rdd.RDD
import org.apache.spark.sql.{SparkSession, SparkSessionExtensions}
import org.apache.spark.sql.catalyst.InternalRow
import org.apache.spark.sql.catalyst.expressions._
import org.apache.spark.sql.catalyst.expressions.aggregate._
import org.apache.spark.sql.catalyst.plans.physical._
import org.apache.spark.sql.catalyst.rules.Rule
import org.apache.spark.sql.execution._
import org.apache.spark.sql.execution.aggregate.HashAggregateExec
import org.apache.spark.sql.execution.exchange.{ShuffleExchangeExec, ShuffleExchangeLike, ShuffleOrigin}
import org.apache.spark.sql.execution.joins.{BroadcastHashJoinExec, SortMergeJoinExec}
import org.apache.spark.sql.execution.metric.{SQLMetric, SQLMetrics}
import org.apache.spark.sql.vectorized.{ColumnarBatch, ColumnVector}

/** spark.sql.extensions entry point (SparkSessionExtensions.scala:116, injectColumnar :168). */
class B200SparkExtensions extends (SparkSessionExtensions => Unit) {
  override def apply(ext: SparkSessionExtensions): Unit = ext.injectColumnar(_ => B200ColumnarRule)
}

/** ColumnarRule (Columnar.scala:47-50).  pre: swap CPU operators for GPU ones and fuse Filter/Project into the
 *  consumer; post: replace RowToColumnarExec / ColumnarToRowExec with host<->HBM copies. */
object B200ColumnarRule extends ColumnarRule {
  override def preColumnarTransitions: Rule[SparkPlan] = new Rule[SparkPlan] {
    def apply(plan: SparkPlan): SparkPlan = plan.transformUp {
      case agg: HashAggregateExec if GpuSupport.supports(agg) =>
        val (cond, projected, source) = GpuSupport.collapse(agg.child)   // Filter/Project chain under the aggregate
        GpuHashAggregateExec(agg, cond, projected, source)
      case s: SortExec if GpuSupport.supports(s) => GpuSortExec(s.sortOrder, s.global, s.child)
      case t: TakeOrderedAndProjectExec => GpuTakeOrderedAndProjectExec(t.limit, t.sortOrder, t.projectList, t.child)
      case j: BroadcastHashJoinExec if GpuSupport.supports(j) => GpuHashJoinExec(j.leftKeys, j.rightKeys, j.joinType, j.left, j.right)
      case j: SortMergeJoinExec if GpuSupport.supports(j) =>
        // same multiset; outputOrdering is dropped, so EnsureRequirements re-inserts a sort only if a parent needs it
        GpuHashJoinExec(j.leftKeys, j.rightKeys, j.joinType, j.left, j.right)
      case e: ShuffleExchangeExec if GpuSupport.supports(e) => GpuShuffleExchangeExec(e.outputPartitioning, e.child, e.shuffleOrigin)
      case f: FilterExec if GpuSupport.supports(f) => GpuFilterProjectExec(Some(f.condition), f.output, f.child)
      case p: ProjectExec if GpuSupport.supports(p) => GpuFilterProjectExec(None, p.projectList, p.child)
    }
  }
  override def postColumnarTransitions: Rule[SparkPlan] = new Rule[SparkPlan] {
    def apply(plan: SparkPlan): SparkPlan = plan.transformUp {
      case RowToColumnarExec(child) => HostToDeviceExec(child)     // Columnar.scala:503-546 -> sb_table_import_host
      case ColumnarToRowExec(child: GpuExec) => ColumnarToRowExec(DeviceToHostExec(child))
    }
  }
}

/** Common contract of the GPU operators (SparkPlan.scala:92, 232, 359). */
trait GpuExec extends SparkPlan {
  override def supportsColumnar: Boolean = true
  override protected def doExecute(): RDD[InternalRow] = throw new IllegalStateException("columnar only: no CPU fallback")
  protected def withStream[T](f: Long => T): T = {
    val s = Native.streamCreate()
    try f(s) finally Native.streamDestroy(s)
  }
}

/** ColumnVector holding one column of an HBM-resident sb_table; close() == sb_table_release
 *  (ColumnarBatch creator closes it, SparkPlan.scala:355-358). */
final class DeviceBatch(val table: Long, schemaTypes: Array[org.apache.spark.sql.types.DataType])
  extends ColumnarBatch(schemaTypes.indices.map(i => new DeviceColumnVector(table, i, schemaTypes(i)): ColumnVector).toArray,
    Native.tableNumRows(table).toInt) {
  override def close(): Unit = Native.tableRelease(table)
}

case class GpuHashAggregateExec(cpu: HashAggregateExec, condition: Option[Expression], inputs: Seq[NamedExpression], child: SparkPlan)
  extends UnaryExecNode with GpuExec {
  override def output: Seq[Attribute] = cpu.output
  override def outputPartitioning: Partitioning = cpu.outputPartitioning
  override def requiredChildDistribution: Seq[Distribution] = cpu.requiredChildDistribution   // Partial -> Exchange -> Final unchanged
  override lazy val metrics: Map[String, SQLMetric] = Map(
    "numOutputRows" -> SQLMetrics.createMetric(sparkContext, "number of output rows"),
    "aggTime" -> SQLMetrics.createTimingMetric(sparkContext, "time in aggregation build"))   // HashAggregateExec.scala:70-86
  override protected def doExecuteColumnar(): RDD[ColumnarBatch] = {
    val mode = GpuSupport.mode(cpu)           // SB_AGG_MODE_PARTIAL / FINAL / COMPLETE (AggUtils.scala:131-208)
    child.executeColumnar().mapPartitions { batches =>
      val in = GpuSupport.concatToDevice(batches)                       // one partition == one device table
      val plan = GpuSupport.compileAgg(cpu, condition, inputs, in)      // key columns, SB_AGG_* codes, sb_expr handles
      val out = withStream(s => Native.hashAggregate(in.table, mode, plan.keyCols, plan.funcs, plan.inputExprs, plan.filterExpr, 0L, s))
      in.close()
      Iterator.single(new DeviceBatch(out, cpu.schema.fields.map(_.dataType)))
    }
  }
  override protected def withNewChildInternal(c: SparkPlan): SparkPlan = copy(child = c)
}

/** Must stay a ShuffleExchangeLike or AQE refuses the plan (AdaptiveSparkPlanExec.scala:977-988). */
case class GpuShuffleExchangeExec(override val outputPartitioning: Partitioning, child: SparkPlan, shuffleOrigin: ShuffleOrigin)
  extends ShuffleExchangeLike with GpuExec {
  // numMappers / numPartitions / mapOutputStatisticsFuture / runtimeStatistics are synthesised from the bucket byte counts the
  // all-to-all already exchanges (ShuffleExchangeExec.scala:51-152); getShuffleRDD returns RDD[ColumnarBatch] for AQEShuffleReadExec.
  override protected def doExecuteColumnar(): RDD[ColumnarBatch] = child.executeColumnar().barrier().mapPartitions { batches =>
    // barrier stage: all P executors run together, which the NCCL all-to-all requires (SURVEY.md 7, hard parts)
    val in = GpuSupport.concatToDevice(batches)
    val n = outputPartitioning.numPartitions
    val offs = new Array[Long](n + 1)
    val out = withStream { s =>
      val parted = outputPartitioning match {
        case HashPartitioning(keys, _) => Native.hashPartition(in.table, GpuSupport.ordinals(keys, child.output), n, s, offs)
        case RoundRobinPartitioning(_) => Native.roundRobinPartition(in.table, GpuSupport.roundRobinStart(n), n, s, offs)
      }
      val recvOffs = new Array[Long](n + 1)
      try Native.allToAll(parted, offs, n, s, recvOffs) finally Native.tableRelease(parted)
    }
    in.close()
    Iterator.single(new DeviceBatch(out, child.schema.fields.map(_.dataType)))
  }
  override protected def withNewChildInternal(c: SparkPlan): SparkPlan = copy(child = c)
}
// GpuSortExec, GpuTakeOrderedAndProjectExec, GpuHashJoinExec, GpuFilterProjectExec, HostToDeviceExec, DeviceToHostExec,
// DeviceColumnVector, GpuSupport (type/expr support checks, ExprCompiler lowering Catalyst expressions to sb_expr) and
// B200Plugin (SparkPlugin: executor init -> Native.init(gpu ordinal from TaskContext.resources()("gpu")); driver ->
// Native.commGetUniqueId broadcast through PluginContext.send) follow the same shape: one sb_* call per partition.
