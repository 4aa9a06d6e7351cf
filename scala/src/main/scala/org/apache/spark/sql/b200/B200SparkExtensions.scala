/*
 * Scala host side of the drop-in: enabled with
 *   --conf spark.sql.extensions=org.apache.spark.sql.b200.B200SparkExtensions
 *   --conf spark.plugins=org.apache.spark.sql.b200.B200Plugin
 *   --conf spark.executor.resource.gpu.amount=1 --conf spark.task.resource.gpu.amount=1
 * Catalyst, the DataFrame API and the SparkPlan operator surface are untouched: the rule only swaps physical
 * operators for GPU ones that keep the reference operator's output / partitioning / ordering contracts.
 *
 * Written against the reference at /root/reference (5.0.0-SNAPSHOT).  This image has no JDK / scalac, so the sources are
 * NOT compiled here; tests/test_capi_cpu.py checks what can be checked without a compiler (every class the rule refers to is
 * defined in this directory, every Native.* call has a declaration in Native.java and a JNI function in the C shim).
 *
 * Files: B200SparkExtensions.scala (this: extension entry point + ColumnarRule), GpuExec.scala (operator contract, DeviceBatch,
 * DeviceColumnVector), GpuOperators.scala (aggregate / sort / top-n / join / filter-project / transitions),
 * GpuExchanges.scala (ShuffleExchangeLike / BroadcastExchangeLike implementations), GpuSupport.scala (support checks and the
 * Catalyst -> sb_expr lowering), B200Plugin.scala (SparkPlugin: GPU binding and NCCL rendezvous).
 */
package org.apache.spark.sql.b200

import org.apache.spark.sql.SparkSessionExtensions
import org.apache.spark.sql.catalyst.expressions.Attribute
import org.apache.spark.sql.catalyst.optimizer.BuildRight
import org.apache.spark.sql.catalyst.plans.{Inner, LeftSemi}
import org.apache.spark.sql.catalyst.rules.Rule
import org.apache.spark.sql.execution._
import org.apache.spark.sql.execution.aggregate.{HashAggregateExec, SortAggregateExec}
import org.apache.spark.sql.execution.window.WindowExec
import org.apache.spark.sql.execution.exchange.{BroadcastExchangeExec, ShuffleExchangeExec}
import org.apache.spark.sql.execution.joins.{BroadcastHashJoinExec, ShuffledHashJoinExec, SortMergeJoinExec}

/** spark.sql.extensions entry point (SparkSessionExtensions.scala:116, injectColumnar :168). */
class B200SparkExtensions extends (SparkSessionExtensions => Unit) {
  override def apply(ext: SparkSessionExtensions): Unit = ext.injectColumnar(_ => B200ColumnarRule)
}

/**
 * ColumnarRule (Columnar.scala:47-50).
 *  pre:  swap CPU operators for GPU ones; Filter / Project chains under an aggregate are folded into it (what
 *        CollapseCodegenStages + WholeStageCodegenExec do for the CPU path).  A node that the GPU path cannot run
 *        (GpuSupport says why) is left alone: ApplyColumnarRulesAndInsertTransitions then puts the row <-> column
 *        transitions around the GPU islands.
 *  post: the transitions around GPU islands become host <-> HBM copies.
 */
object B200ColumnarRule extends ColumnarRule {
  override def preColumnarTransitions: Rule[SparkPlan] = new Rule[SparkPlan] {
    def apply(plan: SparkPlan): SparkPlan = injectRuntimeFilters(replaceOperators(plan))
    private def replaceOperators(plan: SparkPlan): SparkPlan = plan.transformUp {
      case agg: HashAggregateExec if GpuSupport.supports(agg) =>
        val collapsed = GpuSupport.collapse(agg)          // (condition, aggregate inputs over source attributes, source plan)
        GpuHashAggregateExec(agg, collapsed.condition, collapsed.inputs, collapsed.source)
      case agg: SortAggregateExec if GpuSupport.supports(agg) =>           // same answer as the hash aggregate: one device operator
        val collapsed = GpuSupport.collapse(agg)
        GpuHashAggregateExec(agg, collapsed.condition, collapsed.inputs, collapsed.source)
      case e: ExpandExec if GpuSupport.supports(e) => GpuExpandExec(e.projections, e.output, e.child)
      case w: WindowExec if GpuSupport.supports(w) => GpuWindowExec(w.windowExpression, w.partitionSpec, w.orderSpec, w.child)
      case s: SortExec if GpuSupport.supports(s) => GpuSortExec(s.sortOrder, s.global, s.child)
      case t: TakeOrderedAndProjectExec if GpuSupport.supports(t) =>
        GpuTakeOrderedAndProjectExec(t.limit, t.sortOrder, t.projectList, t.child)
      case j: BroadcastHashJoinExec if GpuSupport.supports(j) =>
        GpuHashJoinExec(j.leftKeys, j.rightKeys, j.joinType, j.buildSide, j.condition, j.left, j.right, j.isNullAwareAntiJoin,
          broadcast = true)
      case j: ShuffledHashJoinExec if GpuSupport.supports(j) =>
        GpuHashJoinExec(j.leftKeys, j.rightKeys, j.joinType, j.buildSide, j.condition, j.left, j.right, false, broadcast = false)
      case j: SortMergeJoinExec if GpuSupport.supports(j) =>
        // same multiset; outputOrdering is dropped, so EnsureRequirements re-inserts a sort only if a parent needs one
        GpuHashJoinExec(j.leftKeys, j.rightKeys, j.joinType, GpuSupport.buildSideFor(j), j.condition, j.left, j.right, false,
          broadcast = false)
      case e: ShuffleExchangeExec if GpuSupport.supports(e) =>
        GpuShuffleExchangeExec(e.outputPartitioning, e.child, e.shuffleOrigin, e.advisoryPartitionSize)
      case b: BroadcastExchangeExec if GpuSupport.supports(b) => GpuBroadcastExchangeExec(b.mode, b.child)
      case f: FilterExec if GpuSupport.supports(f) => GpuFilterProjectExec(Some(f.condition), f.output, f.child)
      case p: ProjectExec if GpuSupport.supports(p) => GpuFilterProjectExec(None, p.projectList, p.child)
    }
  }

  /**
   * Runtime filters on the physical plan (the rule the reference runs on the logical plan: InjectRuntimeFilter.scala:196-260, 411-460).
   * An inner broadcast join J2 whose streamed side is (filters / projections of bare attributes over) another inner or left-semi
   * join J1, and whose build side sits under a selective predicate: rows of J1's STREAMED input whose J2 key cannot be a key of J2's
   * build side never reach the output, so J1 tests that key against the relation's prefilter (GpuRuntimeFilter).  At most two per join.
   */
  def injectRuntimeFilters(plan: SparkPlan): SparkPlan = plan.transformUp {
    case j2: GpuHashJoinExec if j2.broadcast && j2.joinType == Inner && j2.condition.isEmpty =>
      val (build, streamed, buildKeys, streamKeys) =
        if (j2.buildSide == BuildRight) (j2.right, j2.left, j2.rightKeys, j2.leftKeys) else (j2.left, j2.right, j2.leftKeys, j2.rightKeys)
      def selective(p: SparkPlan): Boolean = p.isInstanceOf[FilterExec] ||
        (p match { case f: GpuFilterProjectExec => f.condition.isDefined; case _ => false }) || p.children.exists(selective)
      def passThrough(p: SparkPlan): Option[GpuHashJoinExec] = p match {
        case j1: GpuHashJoinExec => Some(j1)
        case f: GpuFilterProjectExec if streamKeys.forall(k => f.projectList.exists(n => n.isInstanceOf[Attribute] && n.semanticEquals(k))) =>
          passThrough(f.child)
        case _ => None
      }
      passThrough(streamed) match {
        case Some(j1) if selective(build) && j1.condition.isEmpty && (j1.joinType == Inner || j1.joinType == LeftSemi) =>
          val j1Streamed = if (j1.buildSide == BuildRight) j1.left else j1.right
          val filters = streamKeys.zip(buildKeys).collect {
            case (s: Attribute, b: Attribute) if j1Streamed.outputSet.contains(s) => GpuRuntimeFilter(s, b, build)
          }.take(2 - j1.runtimeFilters.length)
          if (filters.isEmpty) j2
          else {
            val j1f = j1.copy(runtimeFilters = j1.runtimeFilters ++ filters)
            val rewritten = streamed.transformDown { case x if x eq j1 => j1f }
            if (j2.buildSide == BuildRight) j2.copy(left = rewritten) else j2.copy(right = rewritten)
          }
        case _ => j2
      }
  }

  override def postColumnarTransitions: Rule[SparkPlan] = new Rule[SparkPlan] {
    def apply(plan: SparkPlan): SparkPlan = plan.transformUp {
      case RowToColumnarExec(child) if !child.isInstanceOf[GpuExec] => HostToDeviceExec(child)   // Columnar.scala:503-546
      case ColumnarToRowExec(child: GpuExec) => ColumnarToRowExec(DeviceToHostExec(child))       // Columnar.scala:67-214
    }
  }
}
