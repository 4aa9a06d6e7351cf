#!/usr/bin/env python
"""Multi-GPU parity check of the exchange (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
Every rank hash-partitions its own shard on its GPU, the buckets cross NVLink through sb_all_to_all (NCCL), and the rows a
rank ends up with must be exactly the rows of the partitions it owns in the oracle's shuffle of the whole table
(bit-exact multiset, NULLs included).  Also checks sb_all_gather and the two-stage Q1 (Partial -> AllGather -> Final)."""
import ctypes as C
import os
import sys

import numpy as np
import pyarrow as pa
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O                              # noqa: E402  (checker only)
from spark_b200 import _capi as capi, tpch                  # noqa: E402
from spark_b200.columnar import ColumnarBatch, Stream       # noqa: E402
from spark_b200.execution import HashPartitioning, LocalTableScanExec, ShuffleExchangeExec  # noqa: E402


def shard(rank, n=200_000):
    rng = np.random.default_rng(1000 + rank)
    return pa.table({"k": pa.array(rng.integers(0, 50_000, n), mask=rng.random(n) < 0.03),
                     "d": pa.array(rng.integers(8000, 9000, n).astype(np.int32)).cast(pa.date32()),
                     # NULLs on odd ranks only: even ranks ship this column without a validity bitmap, the exchange must still agree
                     "v": pa.array(rng.random(n), mask=(rng.random(n) < 0.05) if rank % 2 else None),
                     "f": rng.integers(0, 3, n).astype(np.int8),
                     # a string column (rank-dependent vocabulary): codes + dictionaries in the all-to-all, lengths + arena in the all-gather
                     "s": pa.array(np.array(["", "N", "R", "rank%d" % rank, "BUILDING", "你好", "x" * (20 + rank)])[rng.integers(0, 7, n)],
                                   type=pa.string(), mask=rng.random(n) < 0.04)})


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    lib = capi.init(local)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idbuf = torch.zeros(capi.SB_UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = C.create_string_buffer(capi.SB_UNIQUE_ID_BYTES)
        capi.check(lib.sb_comm_get_unique_id(raw))
        idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(idbuf, 0)
    capi.check(lib.sb_comm_init(rank, world, bytes(idbuf.cpu().numpy().tobytes())))
    stream = Stream()
    mine = shard(rank)
    batch = ColumnarBatch.from_arrow(mine, stream)
    whole = pa.concat_tables([shard(r) for r in range(world)])
    key = lambda t: sorted(zip(*[t.column(i).to_pylist() for i in range(t.num_columns)]), key=lambda r: tuple((x is None, x if x is not None else 0) for x in r))
    # both transports -- fused (sb_shuffle_exchange: the multisplit's stores land in the owners' windows) and two-step
    # (sb_hash_partition + sb_all_to_all) -- at several fan-outs, against the oracle's shuffle of the whole table
    for nparts in (200, 2048, 7):
        pid = O.partition_ids(whole, ["k", "d"], nparts)
        lo = [-(-r * nparts // world) for r in range(world + 1)]
        owned = np.nonzero((pid >= lo[rank]) & (pid < lo[rank + 1]))[0]
        want = O.take_table(whole, owned)
        for fused in (True, False):
            ex = ShuffleExchangeExec(HashPartitioning(["k", "d"], nparts), LocalTableScanExec(batch), fused=fused)
            got = ex.executeColumnar(stream).to_arrow(stream)
            assert got.num_rows == want.num_rows, (rank, nparts, fused, got.num_rows, want.num_rows)
            assert key(got) == key(want), "rank %d: exchanged rows differ from the oracle's shuffle (nparts=%d fused=%s)" % (rank, nparts, fused)
            offs = ex.partition_offsets
            counts = np.bincount(pid[owned], minlength=nparts)
            assert np.array_equal(np.diff(offs), counts), (rank, nparts, fused)
            gp = O.partition_ids(got, ["k", "d"], nparts)      # partition-contiguous over the owned partitions, both transports
            assert np.all(np.diff(gp) >= 0), "rank %d: exchange output is not partition-contiguous (nparts=%d fused=%s)" % (rank, nparts, fused)
    # all-gather (broadcast build side)
    h = C.c_void_p()
    small = ColumnarBatch.from_arrow(mine.slice(0, 1000 + rank), stream)
    capi.check(lib.sb_all_gather(small.handle, stream.handle, C.byref(h)))
    allg = ColumnarBatch(h, small.names, small.arrow_types).to_arrow(stream)
    want_g = pa.concat_tables([shard(r).slice(0, 1000 + r) for r in range(world)])
    assert allg.to_pydict() == want_g.to_pydict(), "rank %d: all-gather differs" % rank
    # small tables take the packed single-message path (<= 256 rows on every rank); an empty contribution is legal
    for rows_of in (lambda r: 5 + 3 * r, lambda r: 0 if r == 0 else 256, lambda r: 256 + r):
        tiny = ColumnarBatch.from_arrow(mine.slice(0, rows_of(rank)), stream)
        h = C.c_void_p()
        capi.check(lib.sb_all_gather(tiny.handle, stream.handle, C.byref(h)))
        got_t = ColumnarBatch(h, tiny.names, tiny.arrow_types).to_arrow(stream)
        want_t = pa.concat_tables([shard(r).slice(0, rows_of(r)) for r in range(world)])
        assert got_t.to_pydict() == want_t.to_pydict(), "rank %d: small all-gather differs" % rank
    ok = torch.ones(1, device="cuda")
    dist.all_reduce(ok)
    if rank == 0:
        print("multi_gpu_check ok: %d ranks, all-to-all of %d rows (200 partitions) and all-gather match the oracle" % (world, whole.num_rows))
    capi.check(lib.sb_comm_destroy())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
