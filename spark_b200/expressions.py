"""Catalyst-style expression nodes for the GPU operators.

Names and semantics mirror sql/catalyst/src/main/scala/org/apache/spark/sql/catalyst/expressions of the
reference (arithmetic.scala, predicates.scala, nullExpressions.scala, Cast.scala, aggregate/*.scala); only the
subset the shuffle / sort / aggregate / join path needs.  `compile()` lowers a tree to the postfix sb_expr
program of include/spark_b200.h; `sexpr()` gives the neutral nested-tuple form the test oracle consumes.
"""
from __future__ import annotations

import ctypes as C
import datetime

from . import _capi as capi

VT_BOOL, VT_I32, VT_I64, VT_F64 = capi.SB_VT_BOOL, capi.SB_VT_I32, capi.SB_VT_I64, capi.SB_VT_F64

_VT_OF_SB = {capi.SB_BOOL: VT_BOOL, capi.SB_INT8: VT_I32, capi.SB_INT16: VT_I32, capi.SB_INT32: VT_I32,
             capi.SB_DATE32: VT_I32, capi.SB_INT64: VT_I64, capi.SB_TIMESTAMP: VT_I64, capi.SB_DECIMAL64: VT_I64,
             capi.SB_FLOAT32: VT_F64, capi.SB_FLOAT64: VT_F64}
_OUT_OF_VT = {VT_BOOL: capi.SB_BOOL, VT_I32: capi.SB_INT32, VT_I64: capi.SB_INT64, VT_F64: capi.SB_FLOAT64}


class Schema:
    """names + physical sb types of an operator's input (the `output: Seq[Attribute]` of a SparkPlan)."""

    def __init__(self, names, types):
        self.names = list(names)
        self.types = list(types)

    def index(self, name):
        try:
            return self.names.index(name)
        except ValueError:
            raise KeyError("unresolved attribute %r among %r" % (name, self.names))


class Expression:
    children = ()

    def vtype(self, schema) -> int:
        raise NotImplementedError

    def emit(self, schema, out):
        raise NotImplementedError

    def sexpr(self):
        raise NotImplementedError

    def references(self):
        r = set()
        for c in self.children:
            r |= c.references()
        return r

    # operator sugar
    def __add__(self, o): return Add(self, _wrap(o))
    def __radd__(self, o): return Add(_wrap(o), self)
    def __sub__(self, o): return Subtract(self, _wrap(o))
    def __rsub__(self, o): return Subtract(_wrap(o), self)
    def __mul__(self, o): return Multiply(self, _wrap(o))
    def __rmul__(self, o): return Multiply(_wrap(o), self)
    def __truediv__(self, o): return Divide(self, _wrap(o))
    def __neg__(self): return UnaryMinus(self)
    def __lt__(self, o): return LessThan(self, _wrap(o))
    def __le__(self, o): return LessThanOrEqual(self, _wrap(o))
    def __gt__(self, o): return GreaterThan(self, _wrap(o))
    def __ge__(self, o): return GreaterThanOrEqual(self, _wrap(o))
    def eq(self, o): return EqualTo(self, _wrap(o))
    def ne(self, o): return Not(EqualTo(self, _wrap(o)))
    def __and__(self, o): return And(self, _wrap(o))
    def __or__(self, o): return Or(self, _wrap(o))
    def __invert__(self): return Not(self)
    def is_null(self): return IsNull(self)
    def is_not_null(self): return IsNotNull(self)


def _wrap(x):
    return x if isinstance(x, Expression) else Literal(x)


def _node(op, vtype, arg=0, lit_i=None, lit_d=None):
    n = capi.sb_expr_node()
    n.op = capi.SB_OP[op]
    n.vtype = vtype
    n.arg = arg
    if lit_d is not None:
        n.lit.d = lit_d
    elif lit_i is not None:
        n.lit.i = lit_i
    return n


class AttributeReference(Expression):
    def __init__(self, name):
        self.name = name

    def vtype(self, schema):
        t = schema.types[schema.index(self.name)]
        if t == capi.SB_STRING:
            return 0
        return _VT_OF_SB[t]

    def emit(self, schema, out):
        out.append(_node("COL", self.vtype(schema), schema.index(self.name)))

    def sexpr(self):
        return ("col", self.name)

    def references(self):
        return {self.name}


col = AttributeReference


class Literal(Expression):
    def __init__(self, value, vt=None):
        if isinstance(value, datetime.date):
            value = (value - datetime.date(1970, 1, 1)).days
            vt = vt or VT_I32
        self.value = value
        if vt is None:
            if value is None:
                vt = VT_I64
            elif isinstance(value, bool):
                vt = VT_BOOL
            elif isinstance(value, float):
                vt = VT_F64
            else:
                vt = VT_I32 if -2 ** 31 <= value < 2 ** 31 else VT_I64
        self.vt = vt

    def vtype(self, schema):
        return self.vt

    def emit(self, schema, out):
        if self.value is None:
            out.append(_node("LIT_NULL", self.vt))
        elif self.vt == VT_F64:
            out.append(_node("LIT_F64", VT_F64, lit_d=float(self.value)))
        else:
            out.append(_node("LIT_I64", self.vt, lit_i=int(self.value)))

    def sexpr(self):
        import numpy as np
        dt = {VT_BOOL: bool, VT_I32: np.int32, VT_I64: np.int64, VT_F64: np.float64}[self.vt]
        return ("lit", self.value, dt)


lit = Literal


def _promote(a, b):
    """Binary numeric promotion (TypeCoercion): int32 < int64 < double."""
    return max(a, b)


class _Cast(Expression):
    def __init__(self, child, to_vt):
        self.child = child
        self.children = (child,)
        self.to = to_vt

    def vtype(self, schema):
        return self.to

    def emit(self, schema, out):
        self.child.emit(schema, out)
        src = self.child.vtype(schema)
        if src == self.to:
            return
        op = {VT_F64: "CAST_F64", VT_I64: "CAST_I64", VT_I32: "CAST_I32"}[self.to]
        out.append(_node(op, self.to, src))

    def sexpr(self):
        return ({VT_F64: "cast_f64", VT_I64: "cast_i64", VT_I32: "cast_i32"}[self.to], self.child.sexpr())


def Cast(child, to):
    vt = {"double": VT_F64, "long": VT_I64, "int": VT_I32}.get(to, to)
    return _Cast(child, vt)


def _coerce(e, schema, to_vt):
    """Implicit cast of a child to the operator's type; literals are re-typed instead of cast."""
    vt = e.vtype(schema)
    if vt == to_vt:
        return e
    if isinstance(e, Literal) and e.value is not None:
        return Literal(float(e.value) if to_vt == VT_F64 else e.value, to_vt)
    return _Cast(e, to_vt)


class _BinaryArith(Expression):
    op = None
    name = None

    def __init__(self, left, right):
        self.left, self.right = left, right
        self.children = (left, right)

    def vtype(self, schema):
        return _promote(self.left.vtype(schema), self.right.vtype(schema))

    def emit(self, schema, out):
        vt = self.vtype(schema)
        _coerce(self.left, schema, vt).emit(schema, out)
        _coerce(self.right, schema, vt).emit(schema, out)
        out.append(_node(self.op, vt))

    def sexpr(self):
        return (self.name, self.left.sexpr(), self.right.sexpr())


class Add(_BinaryArith):
    op, name = "ADD", "add"


class Subtract(_BinaryArith):
    op, name = "SUB", "sub"


class Multiply(_BinaryArith):
    op, name = "MUL", "mul"


class Divide(_BinaryArith):
    """Spark's `/` on non-decimals is double division; NULL when the divisor is 0 (non-ANSI)."""
    op, name = "DIV", "div"

    def vtype(self, schema):
        return VT_F64


class UnaryMinus(Expression):
    def __init__(self, child):
        self.child = child
        self.children = (child,)

    def vtype(self, schema):
        return self.child.vtype(schema)

    def emit(self, schema, out):
        self.child.emit(schema, out)
        out.append(_node("NEG", self.vtype(schema)))

    def sexpr(self):
        return ("neg", self.child.sexpr())


class _Comparison(Expression):
    op = None
    name = None

    def __init__(self, left, right):
        self.left, self.right = left, right
        self.children = (left, right)

    def vtype(self, schema):
        return VT_BOOL

    def emit(self, schema, out):
        vt = _promote(self.left.vtype(schema), self.right.vtype(schema))
        _coerce(self.left, schema, vt).emit(schema, out)
        _coerce(self.right, schema, vt).emit(schema, out)
        out.append(_node(self.op, VT_BOOL, vt))   # arg = operand class

    def sexpr(self):
        return (self.name, self.left.sexpr(), self.right.sexpr())


class EqualTo(_Comparison):
    op, name = "EQ", "eq"


class LessThan(_Comparison):
    op, name = "LT", "lt"


class LessThanOrEqual(_Comparison):
    op, name = "LE", "le"


class GreaterThan(_Comparison):
    op, name = "GT", "gt"


class GreaterThanOrEqual(_Comparison):
    op, name = "GE", "ge"


class _Logical(Expression):
    op = None
    name = None

    def __init__(self, left, right):
        self.left, self.right = left, right
        self.children = (left, right)

    def vtype(self, schema):
        return VT_BOOL

    def emit(self, schema, out):
        self.left.emit(schema, out)
        self.right.emit(schema, out)
        out.append(_node(self.op, VT_BOOL))

    def sexpr(self):
        return (self.name, self.left.sexpr(), self.right.sexpr())


class And(_Logical):
    op, name = "AND", "and"


class Or(_Logical):
    op, name = "OR", "or"


class _Unary(Expression):
    op = None
    name = None

    def __init__(self, child):
        self.child = child
        self.children = (child,)

    def vtype(self, schema):
        return VT_BOOL

    def emit(self, schema, out):
        self.child.emit(schema, out)
        out.append(_node(self.op, VT_BOOL))

    def sexpr(self):
        return (self.name, self.child.sexpr())


class Not(_Unary):
    op, name = "NOT", "not"


class IsNull(_Unary):
    op, name = "ISNULL", "isnull"


class IsNotNull(_Unary):
    op, name = "ISNOTNULL", "isnotnull"


class CompiledExpr:
    """Owns the ctypes arrays behind one sb_expr."""

    def __init__(self, expr: Expression, schema: Schema):
        nodes = []
        expr.emit(schema, nodes)
        self.arr = (capi.sb_expr_node * len(nodes))(*nodes)
        self.c = capi.sb_expr()
        self.c.nodes = C.cast(self.arr, C.POINTER(capi.sb_expr_node))
        self.c.n = len(nodes)
        if isinstance(expr, AttributeReference):
            self.c.out_type = schema.types[schema.index(expr.name)]
        else:
            self.c.out_type = _OUT_OF_VT[expr.vtype(schema)]


# ---- aggregate functions (expressions/aggregate/{Sum,Average,Count,Min,Max}.scala) -----------------
class AggregateFunction:
    func = None

    def __init__(self, child=None):
        self.child = _wrap(child) if child is not None else None

    def sexpr_input(self):
        return None if self.child is None else self.child.sexpr()


class Sum(AggregateFunction):
    func = "sum"


class Average(AggregateFunction):
    func = "avg"


class Count(AggregateFunction):
    """Count(expr); Count() / Count(Literal(1)) is count(*)."""
    func = "count"

    def __init__(self, child=None):
        if child is None or (isinstance(child, Literal)) or isinstance(child, int):
            self.child = None
            self.func = "count_star"
        else:
            self.child = _wrap(child)


class Min(AggregateFunction):
    func = "min"


class Max(AggregateFunction):
    func = "max"


class SortOrder:
    """SortOrder(child, direction, nullOrdering) (SortOrder.scala:63); default null ordering: ASC -> NULLS
    FIRST, DESC -> NULLS LAST."""

    def __init__(self, child, ascending=True, nulls_first=None):
        self.child = child if isinstance(child, str) else child.name
        self.ascending = ascending
        self.nulls_first = ascending if nulls_first is None else nulls_first

    def as_tuple(self):
        return (self.child, self.ascending, self.nulls_first)
