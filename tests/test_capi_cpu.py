"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/spark_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os

import numpy as np
import pytest


def test_library_exports_every_declared_symbol():
    from spark_b200 import _capi as capi
    lib = capi.load()
    syms = capi.declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert set(capi._SIGNATURES) == set(syms), set(capi._SIGNATURES) ^ set(syms)
    assert lib.sb_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from spark_b200 import _capi as capi
    lib = capi.load()
    assert lib.sb_init(0) != 0
    assert b"no CPU fallback" in lib.sb_last_error()
    h = C.c_void_p()
    assert lib.sb_stream_create(C.byref(h)) == 6          # SB_ERR_NOT_INITIALIZED
    cols = (capi.sb_column * 1)()
    assert lib.sb_table_import_host(cols, 0, None, C.byref(h)) == 6


def test_exchange_plan_contiguous_ownership():
    """Host-only planner of the all-to-all: rank r owns partitions [ceil(r*n/R), ceil((r+1)*n/R))."""
    from spark_b200 import _capi as capi
    lib = capi.load()
    rng = np.random.default_rng(0)
    for nparts, nranks in [(200, 8), (2048, 8), (5, 2), (3, 4), (1, 2)]:
        counts = rng.integers(0, 100, nparts)
        offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        send = np.zeros(nranks, np.int64)
        rc = lib.sb_exchange_plan(offs.ctypes.data_as(C.POINTER(C.c_int64)), nparts, nranks,
                                  send.ctypes.data_as(C.POINTER(C.c_int64)))
        assert rc == 0
        assert send.sum() == counts.sum()
        lo = [-(-r * nparts // nranks) for r in range(nranks + 1)]
        for r in range(nranks):
            assert send[r] == counts[lo[r]:lo[r + 1]].sum()


def test_expression_lowering_and_fusion_rule():
    from spark_b200 import _capi as capi
    from spark_b200.expressions import CompiledExpr, Literal, Schema, Sum, col
    from spark_b200.execution import B200ColumnarRule, FilterExec, HashAggregateExec, ProjectExec, SparkPlan
    schema = Schema(["p", "d", "s"], [capi.SB_FLOAT64, capi.SB_FLOAT64, capi.SB_DATE32])
    e = col("p") * (Literal(1) - col("d"))
    ce = CompiledExpr(e, schema)
    ops = [ce.arr[i].op for i in range(ce.c.n)]
    assert ops == [capi.SB_OP["COL"], capi.SB_OP["LIT_F64"], capi.SB_OP["COL"], capi.SB_OP["SUB"], capi.SB_OP["MUL"]]
    assert ce.arr[1].lit.d == 1.0 and ce.c.out_type == capi.SB_FLOAT64
    assert e.sexpr() == ("mul", ("col", "p"), ("sub", ("lit", 1, np.int32), ("col", "d")))

    class Leaf(SparkPlan):
        pass
    leaf = Leaf()
    plan = HashAggregateExec(["k"], [(Sum(col("x")), "sx")],
                             ProjectExec([("k", col("k")), ("x", col("p") * col("d"))], FilterExec(col("s") <= Literal(10), leaf)),
                             mode="partial")
    fused = B200ColumnarRule().preColumnarTransitions(plan)
    assert isinstance(fused, HashAggregateExec) and fused.child is leaf
    assert fused.condition.sexpr()[0] == "le"
    assert fused.aggregateExpressions[0][0].child.sexpr() == ("mul", ("col", "p"), ("col", "d"))


def test_jni_shim_typechecks_against_the_header():
    """scala/src/main/native/sparkb200_jni.c cannot be built here (no JDK), but it must at least agree with include/spark_b200.h:
    gcc -fsyntax-only with a stand-in jni.h catches signature drift between the C ABI and the JNI face, and every `native`
    method of Native.java must have its Java_..._Native_<name> definition."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = os.path.join(root, "scala/src/main/native/sparkb200_jni.c")
    r = subprocess.run(["/usr/bin/gcc", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(root, "scala/src/main/native/jni_stub"),
                        "-I" + os.path.join(root, "include"), shim], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    java = open(os.path.join(root, "scala/src/main/java/org/apache/spark/sql/b200/Native.java")).read()
    declared = set(re.findall(r"public static native [\w\[\]]+ (\w+)\(", java))
    src = open(shim).read()
    defined = set(re.findall(r"Java_org_apache_spark_sql_b200_Native_(\w+)\(", src)) | set(re.findall(r"NATIVE\(\w+, (\w+)\)", src))
    assert declared and declared <= defined, sorted(declared - defined)


def test_scala_plugin_is_self_consistent():
    """No scalac in this image, so the plugin sources are checked as text for the defects a compiler would report first (and that
    round 1 had): every class / object the sources refer to with the plugin's naming is defined in the directory, every Native.*
    call has a `native` declaration, every abstract member of ShuffleExchangeLike / BroadcastExchangeLike (reference:
    ShuffleExchangeExec.scala:55-151, BroadcastExchangeExec.scala:45-90) is implemented, braces and parentheses balance."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = sorted(glob.glob(os.path.join(root, "scala/src/main/scala/org/apache/spark/sql/b200/*.scala")))
    assert len(files) >= 6
    text = {f: open(f).read() for f in files}
    code = {f: re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", t, flags=re.S)) for f, t in text.items()}
    allcode = "\n".join(code.values())
    defined = set(re.findall(r"\b(?:class|object|trait)\s+(\w+)", allcode))
    used = set(re.findall(r"\b((?:Gpu|B200|Device|HostTo|DeviceTo|ExprCompiler|BroadcastExchangeLikeThreads)\w*)\b", allcode))
    missing = {u for u in used if u not in defined and u not in ("B200Exception",)}      # B200Exception is the Java class next door
    assert not missing, "referenced but not defined: %s" % sorted(missing)
    assert os.path.exists(os.path.join(root, "scala/src/main/java/org/apache/spark/sql/b200/B200Exception.java"))
    java = open(os.path.join(root, "scala/src/main/java/org/apache/spark/sql/b200/Native.java")).read()
    declared = set(re.findall(r"public static native [\w\[\]]+ (\w+)\(", java))
    called = set(re.findall(r"\bNative\.(\w+)\(", allcode))
    assert called <= declared, "Native methods used but not declared: %s" % sorted(called - declared)
    for f, c in code.items():
        c = re.sub(r'"(?:[^"\\]|\\.)*"', '""', c)        # string literals may hold brackets
        for a, b in ("{}", "()", "[]"):
            assert c.count(a) == c.count(b), "%s: unbalanced %s%s (%d vs %d)" % (os.path.basename(f), a, b, c.count(a), c.count(b))
    ex = code[[f for f in files if f.endswith("GpuExchanges.scala")][0]]
    shuffle = ex[ex.index("case class GpuShuffleExchangeExec"):ex.index("class DeviceShuffleReadRDD")]
    for member in ("numMappers", "numPartitions", "advisoryPartitionSize", "shuffleOrigin", "mapOutputStatisticsFuture", "getShuffleRDD",
                   "runtimeStatistics", "shuffleId", "doExecuteColumnar", "withNewChildInternal"):
        assert re.search(r"\b%s\b" % member, shuffle), "GpuShuffleExchangeExec lacks %s" % member
    bcast = ex[ex.index("case class GpuBroadcastExchangeExec"):ex.index("object GpuBroadcastExchangeExec")]
    for member in ("runId", "relationFuture", "completionFuture", "runtimeStatistics", "doExecuteBroadcast", "doPrepare"):
        assert re.search(r"\b%s\b" % member, bcast), "GpuBroadcastExchangeExec lacks %s" % member


def test_runtime_filter_injection_rule():
    """B200ColumnarRule.injectRuntimeFilters: inner join over an inner join, creation side with a FilterExec -> the lower join gets a
    RuntimeFilter on its streamed column and both consumers share one ReusedExchangeExec; no filter -> plan untouched; key coming from
    the lower join's BUILD side -> untouched."""
    from spark_b200.expressions import Literal, col
    from spark_b200.execution import (B200ColumnarRule, BroadcastHashJoinExec, FilterExec, ProjectExec, ReusedExchangeExec, SparkPlan)

    class Leaf(SparkPlan):
        def __init__(self, names):
            self.output_names = names
    lineitem, orders, supplier = Leaf(["l_orderkey", "l_suppkey", "l_price"]), Leaf(["o_orderkey", "o_nation"]), Leaf(["s_suppkey", "s_region"])
    sup_f = ProjectExec(["s_suppkey"], FilterExec(col("s_region").eq(Literal(2)), supplier))
    j1 = BroadcastHashJoinExec(["l_orderkey"], ["o_orderkey"], "inner", "right", lineitem, orders)
    j2 = BroadcastHashJoinExec(["l_suppkey"], ["s_suppkey"], "inner", "right", ProjectExec(["l_suppkey", "o_nation", "l_price"], j1), sup_f)
    out = B200ColumnarRule().injectRuntimeFilters(j2)
    assert out is j2 and isinstance(j2.right, ReusedExchangeExec) and j2.right.uses == 2 and j2.right.child is sup_f
    assert len(j1.runtimeFilters) == 1
    rf = j1.runtimeFilters[0]
    assert (rf.applicationKey, rf.creationKey) == ("l_suppkey", "s_suppkey") and rf.creationPlan is j2.right
    # no selective predicate on the creation side: nothing happens
    j1b = BroadcastHashJoinExec(["l_orderkey"], ["o_orderkey"], "inner", "right", lineitem, orders)
    j2b = BroadcastHashJoinExec(["l_suppkey"], ["s_suppkey"], "inner", "right", j1b, supplier)
    B200ColumnarRule().injectRuntimeFilters(j2b)
    assert j1b.runtimeFilters == [] and j2b.right is supplier
    # the key is produced by the lower join's build side: a filter on the streamed input would test the wrong table
    j1c = BroadcastHashJoinExec(["l_orderkey"], ["o_orderkey"], "inner", "right", lineitem, orders)
    j2c = BroadcastHashJoinExec(["o_nation"], ["s_suppkey"], "inner", "right", j1c, sup_f)
    B200ColumnarRule().injectRuntimeFilters(j2c)
    assert j1c.runtimeFilters == []
    # an outer join below: its streamed rows must all survive
    j1d = BroadcastHashJoinExec(["l_orderkey"], ["o_orderkey"], "left_outer", "right", lineitem, orders)
    j2d = BroadcastHashJoinExec(["l_suppkey"], ["s_suppkey"], "inner", "right", j1d, sup_f)
    B200ColumnarRule().injectRuntimeFilters(j2d)
    assert j1d.runtimeFilters == []


def test_column_pruning_hints_reach_the_join():
    """ColumnPruning's effect on a join (Optimizer.scala, object ColumnPruning): the ProjectExec / HashAggregateExec directly above a
    join tells it which attributes are read, so only those are materialised (BroadcastHashJoinExec.requiredOutput)."""
    from spark_b200.expressions import Literal, Sum, col
    from spark_b200.execution import BroadcastHashJoinExec, HashAggregateExec, ProjectExec, SparkPlan

    seen = {}

    class Leaf(SparkPlan):
        pass

    class Probe(BroadcastHashJoinExec):
        def executeColumnar(self, stream=None):
            seen["need"] = self.requiredOutput
            raise StopIteration          # nothing to run on a CPU-only machine: the hint is what is under test

    j = Probe(["a"], ["b"], "inner", "right", Leaf(), Leaf())
    for plan in (ProjectExec(["x", ("y", col("p") * (Literal(1) - col("d")))], j),
                 HashAggregateExec(["k"], [(Sum(col("p") * col("d")), "s")], j, mode="partial", condition=col("z") > Literal(3))):
        seen.clear()
        try:
            plan.executeColumnar(None)
        except StopIteration:
            pass
        assert seen["need"] == ({"x", "p", "d"} if isinstance(plan, ProjectExec) else {"k", "p", "d", "z"}), seen


def test_reused_exchange_runs_its_child_once_per_round():
    """ReusedExchangeExec: the child runs for the first consumer, the batch is shared with the others and released after the last --
    and the next execution of the plan (the next step of the bench) starts a new round."""
    from spark_b200.execution import ReusedExchangeExec, SparkPlan
    log = []

    class Batch:
        names = ["a"]

        def __init__(self, tag):
            self.tag = tag

        def rename(self, names):
            log.append(("share", self.tag))
            return Batch(self.tag)

        def close(self):
            log.append(("close", self.tag))

    class Child(SparkPlan):
        runs = 0

        def executeColumnar(self, stream=None):
            Child.runs += 1
            return Batch(Child.runs)

    ex = ReusedExchangeExec(Child(), uses=2)
    a, b = ex.executeColumnar(), ex.executeColumnar()
    assert Child.runs == 1 and (a.tag, b.tag) == (1, 1)
    assert log == [("share", 1), ("share", 1), ("close", 1)]
    c = ex.executeColumnar()
    assert Child.runs == 2 and c.tag == 2
