"""GPU checks of the structural table ops the operators are built from: slice and concat keep values,
string offsets and validity bits at arbitrary (non multiple-of-8) row offsets."""
import pyarrow as pa
import pytest

from test_gpu_partition import _mixed_table
from util import assert_tables_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cuts", [[0, 1, 13, 13, 4097, 10000], [0, 10000], [0, 3, 5, 9, 64, 65, 9999, 10000]])
def test_slice_then_concat_is_identity(gpu, stream, cuts):
    from spark_b200.columnar import ColumnarBatch
    t = _mixed_table(10000, seed=5)
    b = ColumnarBatch.from_arrow(t, stream)
    pieces = [b.slice(lo, hi, stream) for lo, hi in zip(cuts[:-1], cuts[1:])]
    for p, lo, hi in zip(pieces, cuts[:-1], cuts[1:]):
        assert_tables_equal(p.to_arrow(stream), t.slice(lo, hi - lo), ordered=True)
    whole = ColumnarBatch.concat(pieces, stream)
    assert_tables_equal(whole.to_arrow(stream), t, ordered=True)


def test_concat_mixes_columns_with_and_without_validity(gpu, stream):
    from spark_b200.columnar import ColumnarBatch
    a = pa.table({"x": pa.array([1, 2, 3], type=pa.int64())})
    b = pa.table({"x": pa.array([None, 5, None, 7, 8], type=pa.int64())})
    got = ColumnarBatch.concat([ColumnarBatch.from_arrow(a, stream), ColumnarBatch.from_arrow(b, stream),
                                ColumnarBatch.from_arrow(a, stream)], stream).to_arrow(stream)
    assert got.column("x").to_pylist() == [1, 2, 3, None, 5, None, 7, 8, 1, 2, 3]
