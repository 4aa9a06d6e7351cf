"""Comparison helpers shared by the parity tests (the reference's checkAnswer rules:
sql/core/src/test/scala/org/apache/spark/sql/RowComparisonUtils.scala:87-112 -- rows compared as a
multiset unless the query orders them; integers/keys/counts exact; floating SUM/AVG within 1e-6 relative,
the tolerance BASELINE.json:north_star states)."""
import math

import numpy as np
import pyarrow as pa

FLOAT_RTOL = 1e-6


def _rows(table: pa.Table):
    cols = [table.column(i).to_pylist() for i in range(table.num_columns)]
    return list(zip(*cols)) if cols else []


def _sort_key(row, key_idx):
    out = []
    for i in key_idx:
        x = row[i]
        if x is None:
            out.append((0, 0))
        elif isinstance(x, float) and math.isnan(x):
            out.append((2, 0))
        else:
            out.append((1, x))
    return tuple(out)


def assert_tables_equal(got: pa.Table, want: pa.Table, ordered=False, key_cols=None, rtol=FLOAT_RTOL):
    assert got.num_columns == want.num_columns, (got.schema, want.schema)
    assert got.num_rows == want.num_rows, "row count %d != %d" % (got.num_rows, want.num_rows)
    g, w = _rows(got), _rows(want)
    if not ordered:
        if key_cols is None:
            # sort on every non-float column (exact), floats compared after alignment
            key_idx = [i for i in range(want.num_columns) if not pa.types.is_floating(want.schema.field(i).type)]
            if not key_idx:
                key_idx = list(range(want.num_columns))
        else:
            key_idx = [want.column_names.index(k) for k in key_cols]
        g = sorted(g, key=lambda r: _sort_key(r, key_idx))
        w = sorted(w, key=lambda r: _sort_key(r, key_idx))
    for ri, (a, b) in enumerate(zip(g, w)):
        for ci, (x, y) in enumerate(zip(a, b)):
            if x is None or y is None:
                assert x is None and y is None, "row %d col %s: %r vs %r" % (ri, want.column_names[ci], x, y)
            elif isinstance(y, float) or isinstance(x, float):
                if math.isnan(y):
                    assert math.isnan(x), "row %d col %s: %r vs NaN" % (ri, want.column_names[ci], x)
                else:
                    assert math.isclose(x, y, rel_tol=rtol, abs_tol=1e-12), \
                        "row %d col %s: %r vs %r" % (ri, want.column_names[ci], x, y)
            else:
                assert x == y, "row %d col %s: %r vs %r" % (ri, want.column_names[ci], x, y)


def random_nullable(rng, values, null_frac):
    mask = rng.random(len(values)) < null_frac
    return pa.array(values, mask=mask)
