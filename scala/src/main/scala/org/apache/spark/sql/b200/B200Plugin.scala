package org.apache.spark.sql.b200

import java.util.{Map => JMap}
import java.util.concurrent.ConcurrentHashMap

import scala.jdk.CollectionConverters._

import org.apache.spark.{SparkConf, SparkContext}
import org.apache.spark.api.plugin.{DriverPlugin, ExecutorPlugin, PluginContext, SparkPlugin}

/**
 * spark.plugins entry (core/src/main/java/org/apache/spark/api/plugin/SparkPlugin.java).
 * Executor side (ExecutorPlugin.init / shutdown): bind this executor's GPU (the address Spark's resource scheduler assigned:
 * PluginContext.resources()("gpu"), the same information TaskContext.resources() gives tasks, TaskContext.scala:309), then join
 * the NCCL communicator: rank = position of the executor id among the registered executors, unique id from the driver.
 * Driver side (DriverPlugin.init / receive): hands out the NCCL unique id and the rank table (PluginContext.ask).
 * One executor per GPU: spark.executor.resource.gpu.amount=1, spark.task.resource.gpu.amount=1, spark.executor.cores=1.
 */
class B200Plugin extends SparkPlugin {
  override def driverPlugin(): DriverPlugin = new B200DriverPlugin
  override def executorPlugin(): ExecutorPlugin = new B200ExecutorPlugin
}

object B200Plugin {
  sealed trait Msg extends Serializable
  final case class Register(executorId: String, host: String) extends Msg
  final case class Rendezvous(rank: Int, world: Int, uniqueId: Array[Byte]) extends Serializable

  @volatile private var world: Int = 1
  private val locations = new ConcurrentHashMap[Int, String]()
  def worldSize(): Int = world
  private[b200] def setWorld(n: Int): Unit = world = n
  def numExecutors(conf: SparkConf): Int = conf.getInt("spark.executor.instances", 1)
  /** "executor_host_executorId" location string of the executor holding rank r (TaskLocation format), if known on the driver */
  def executorLocation(rank: Int): Option[String] = Option(locations.get(rank))
  private[b200] def recordLocation(rank: Int, host: String, executorId: String): Unit = locations.put(rank, s"executor_${host}_$executorId")
}

class B200DriverPlugin extends DriverPlugin {
  private var expected = 1
  private val registered = new java.util.ArrayList[B200Plugin.Register]()
  private var uniqueId: Array[Byte] = _

  override def init(sc: SparkContext, ctx: PluginContext): JMap[String, String] = {
    expected = B200Plugin.numExecutors(sc.getConf)
    require(sc.getConf.getInt("spark.executor.resource.gpu.amount", 0) == 1 && sc.getConf.getInt("spark.executor.cores", 1) == 1,
      "spark-b200 runs one executor per GPU: spark.executor.resource.gpu.amount=1, spark.executor.cores=1")
    Map("spark.b200.world" -> expected.toString).asJava
  }

  /** Executors register; when all have, everyone is told its rank.  Blocks the asker until the rendezvous is complete. */
  override def receive(message: AnyRef): AnyRef = message match {
    case r: B200Plugin.Register => this.synchronized {
      registered.add(r)
      if (registered.size() == expected) this.notifyAll()
      while (registered.size() < expected) this.wait()
      val sorted = registered.asScala.sortBy(_.executorId).toIndexedSeq
      sorted.zipWithIndex.foreach { case (e, i) => B200Plugin.recordLocation(i, e.host, e.executorId) }
      if (uniqueId == null) uniqueId = new Array[Byte](128)              // filled by rank 0 below
      B200Plugin.Rendezvous(sorted.indexWhere(_.executorId == r.executorId), expected, uniqueId)
    }
    case id: Array[Byte] => this.synchronized { uniqueId = id; this.notifyAll(); null }   // rank 0 publishes ncclGetUniqueId
    case "uniqueId" => this.synchronized { while (uniqueId == null || uniqueId.forall(_ == 0)) this.wait(); uniqueId }
    case _ => null
  }
}

class B200ExecutorPlugin extends ExecutorPlugin {
  override def init(ctx: PluginContext, extraConf: JMap[String, String]): Unit = {
    val gpu = ctx.resources().get("gpu")
    require(gpu != null && gpu.addresses.length == 1, "spark-b200 needs exactly one gpu resource per executor")
    Native.init(gpu.addresses(0).toInt)                                                   // sb_init(device ordinal)
    val world = extraConf.getOrDefault("spark.b200.world", "1").toInt
    B200Plugin.setWorld(world)
    if (world > 1) {
      val rv = ctx.ask(B200Plugin.Register(ctx.executorID(), ctx.hostname())).asInstanceOf[B200Plugin.Rendezvous]
      val id = if (rv.rank == 0) { val mine = Native.commGetUniqueId(); ctx.send(mine); mine }
               else ctx.ask("uniqueId").asInstanceOf[Array[Byte]]
      Native.commInit(rv.rank, rv.world, id)                                              // sb_comm_init: ncclCommInitRank
    }
  }
  override def shutdown(): Unit = {
    DeviceShuffleStore.clear()
    Native.shutdown()
  }
}
