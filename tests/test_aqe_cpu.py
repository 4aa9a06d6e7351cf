"""CPU tests: ShufflePartitionsUtil.coalescePartitions -- the oracle's restatement and the product's host-only
sb_coalesce_partitions -- against the literal vectors of ShufflePartitionsUtilSuite.scala:54-300."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O

# (bytesByPartitionId per shuffle, kwargs, expected specs per shuffle: [(start, end, dataSize)] or [] = Nil)
VECTORS = [
    ([[10, 0, 20, 0, 0]], dict(target=100), [[(0, 5, 30)]]),
    ([[10, 0, 90, 20, 0]], dict(target=100), [[(0, 3, 100), (3, 5, 20)]]),
    ([[110, 10, 100, 110, 0]], dict(target=100), [[(0, 1, 110), (1, 2, 10), (2, 3, 100), (3, 4, 110)]]),
    ([[100, 110, 100, 110, 110]], dict(target=100), []),
    ([[30, 30, 0, 40, 110]], dict(target=100), [[(0, 4, 100), (4, 5, 110)]]),
    ([[0, 10, 0, 20, 0], [30, 0, 20, 0, 20]], dict(target=100), [[(0, 5, 30)], [(0, 5, 70)]]),
    ([[0, 10, 0, 20, 0], [30, 0, 70, 0, 30]], dict(target=100), [[(0, 2, 10), (2, 4, 20), (4, 5, 0)], [(0, 2, 30), (2, 4, 70), (4, 5, 30)]]),
    ([[0, 99, 0, 20, 0], [30, 0, 70, 0, 30]], dict(target=100),
     [[(0, 1, 0), (1, 2, 99), (2, 4, 20), (4, 5, 0)], [(0, 1, 30), (1, 2, 0), (2, 4, 70), (4, 5, 30)]]),
    ([[0, 100, 0, 30, 0], [30, 0, 70, 0, 30]], dict(target=100),
     [[(0, 1, 0), (1, 2, 100), (2, 4, 30), (4, 5, 0)], [(0, 1, 30), (1, 2, 0), (2, 4, 70), (4, 5, 30)]]),
    ([[0, 100, 40, 30, 0], [30, 0, 60, 0, 110]], dict(target=100), []),
    ([[100, 100, 40, 30, 0], [30, 0, 60, 70, 110]], dict(target=100), []),
    ([[1] * 10], dict(target=100), [[(0, 10, 10)]]),
    ([[1] * 10], dict(target=100, maxr=4), [[(0, 4, 4), (4, 8, 4), (8, 10, 2)]]),
    ([[1] * 3], dict(target=100, maxr=1), []),
    ([[1, 0, 0, 0, 0, 0, 0, 0, 1]], dict(target=100, maxr=4), [[(0, 4, 1), (8, 9, 1)]]),
    ([[30, 30, 1, 1, 100]], dict(target=60, min_size=10, maxr=3), [[(0, 2, 60), (2, 5, 102)]]),
    ([[60, 1, 1, 1]], dict(target=60, min_size=10, maxr=3), [[(0, 1, 60), (1, 4, 3)]]),
]


@pytest.mark.parametrize("case", range(len(VECTORS)))
def test_oracle_coalesce_partitions_reference_vectors(case):
    stats, kw, want = VECTORS[case]
    got = O.coalesce_partitions(stats, kw["target"], kw.get("min_parts", 1), kw.get("min_size", 0), kw.get("maxr", 2 ** 31 - 1))
    assert got == want


@pytest.mark.parametrize("case", range(len(VECTORS)))
def test_product_coalesce_partitions_reference_vectors(case):
    from spark_b200 import _capi as capi
    lib = capi.load()
    stats, kw, want = VECTORS[case]
    ns, npart = len(stats), len(stats[0])
    arrs = [np.array(s, np.int64) for s in stats]
    ptrs = (C.POINTER(C.c_int64) * ns)(*[a.ctypes.data_as(C.POINTER(C.c_int64)) for a in arrs])
    st, en = (C.c_int32 * npart)(), (C.c_int32 * npart)()
    sizes = (C.c_int64 * (ns * npart))()
    n = C.c_int32()
    capi.check(lib.sb_coalesce_partitions(ptrs, ns, npart, kw["target"], kw.get("min_parts", 1), kw.get("min_size", 0), kw.get("maxr", 2 ** 31 - 1),
                                          st, en, sizes, C.byref(n)))
    got = [[(st[k], en[k], sizes[s * n.value + k]) for k in range(n.value)] for s in range(ns)] if n.value else []
    assert got == want


def test_mismatched_partition_counts_are_not_coalesced():
    assert O.coalesce_partitions([[0] * 5, [0] * 6], 100) == []        # ShufflePartitionsUtilSuite.scala:103-114
