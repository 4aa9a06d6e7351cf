"""CPU tests of B200ColumnarRule.preColumnarTransitions (the CollapseCodegenStages analogue): Project / Filter chains under a
HashAggregateExec are folded into it with substitutions composed top-down (Agg(Filter(Project)) and
Agg(Project(Project)) were collapsed wrongly)."""
from spark_b200.execution import B200ColumnarRule, FilterExec, HashAggregateExec, ProjectExec, SparkPlan
from spark_b200.expressions import Literal, Sum, col


class Leaf(SparkPlan):
    pass


def _rule(plan):
    return B200ColumnarRule().preColumnarTransitions(plan)


def test_filter_above_project_is_rewritten_through_it():
    leaf = Leaf()
    plan = HashAggregateExec(["k"], [(Sum(col("x")), "s")],
                             FilterExec(col("x") > Literal(10), ProjectExec([("k", col("k")), ("x", col("x") * Literal(2))], leaf)))
    r = _rule(plan)
    assert r.child is leaf
    assert r.condition.sexpr() == (col("x") * Literal(2) > Literal(10)).sexpr()          # the filter sees x*2, not the source x
    assert r.aggregateExpressions[0][0].child.sexpr() == (col("x") * Literal(2)).sexpr()


def test_stacked_projects_compose_top_down():
    leaf = Leaf()
    plan = HashAggregateExec(["k"], [(Sum(col("rev")), "s")],
                             ProjectExec([("k", col("k")), ("rev", col("y") * Literal(3))],
                                         ProjectExec([("k", col("k")), ("y", col("x") + Literal(1))], leaf)))
    r = _rule(plan)
    assert r.child is leaf and r.condition is None
    assert r.aggregateExpressions[0][0].child.sexpr() == ((col("x") + Literal(1)) * Literal(3)).sexpr()


def test_filters_on_several_levels_become_one_conjunction_over_source_attributes():
    leaf = Leaf()
    plan = HashAggregateExec(["k"], [(Sum(col("rev")), "s")],
                             FilterExec(col("rev") > Literal(0),
                                        ProjectExec([("k", col("k")), ("rev", col("y") * Literal(3))],
                                                    FilterExec(col("y") < Literal(5),
                                                               ProjectExec([("k", col("k")), ("y", col("x") + Literal(1))], leaf)))))
    r = _rule(plan)
    y = col("x") + Literal(1)
    assert r.child is leaf
    assert r.condition.sexpr() == ((y * Literal(3) > Literal(0)) & (y < Literal(5))).sexpr()


def test_renamed_or_computed_group_keys_are_not_collapsed():
    leaf = Leaf()
    plan = HashAggregateExec(["g"], [(Sum(col("x")), "s")], ProjectExec([("g", col("k")), ("x", col("x"))], leaf))
    assert isinstance(_rule(plan).child, ProjectExec)
    plan = HashAggregateExec(["k"], [(Sum(col("x")), "s")], ProjectExec([("k", col("k") + Literal(1)), ("x", col("x"))], leaf))
    assert isinstance(_rule(plan).child, ProjectExec)
