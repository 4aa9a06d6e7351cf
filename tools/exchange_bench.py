#!/usr/bin/env python
"""ShuffleExchangeExec across GPUs (BASELINE.json configs[3] shape, scaled to what N GPUs of one box hold): every rank
hash-partitions its own 24 M-row x 74 B lineitem-shaped shard into 2048 partitions on its GPU and the buckets cross NVLink
through sb_all_to_all (NCCL grouped send/recv, counts first).  Run under torchrun, one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/exchange_bench.py
Prints one JSON line (rank 0): rows/s of the whole exchange (all ranks' rows / max-over-ranks device time), the map side
(partition) and the transport (all-to-all) separately, and the bytes each rank put on NVLink."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spark_b200 import _capi as capi                        # noqa: E402
from spark_b200.columnar import ColumnarBatch, Stream       # noqa: E402
from spark_b200.execution import HashPartitioning, LocalTableScanExec, ShuffleExchangeExec  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24_000_000
    nparts = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
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
    rng = np.random.default_rng(100 + rank)
    cols = {"l_orderkey": rng.integers(0, 1 << 40, n), "l_partkey": rng.integers(0, 1 << 30, n), "l_suppkey": rng.integers(0, 1 << 24, n),
            "l_quantity": rng.random(n), "l_extendedprice": rng.random(n), "l_discount": rng.random(n), "l_tax": rng.random(n),
            "l_linenumber": rng.integers(1, 8, n).astype(np.int32), "l_shipdate": rng.integers(8000, 10600, n).astype(np.int32),
            "l_commitdate": rng.integers(8000, 10600, n).astype(np.int32), "l_receiptdate": rng.integers(8000, 10600, n).astype(np.int32),
            "l_returnflag": rng.integers(65, 83, n).astype(np.int8), "l_linestatus": rng.integers(70, 80, n).astype(np.int8)}
    rowbytes = sum(a.dtype.itemsize for a in cols.values())
    batch = ColumnarBatch.from_numpy(cols, stream)
    stream.synchronize()
    ex = ShuffleExchangeExec(HashPartitioning(["l_orderkey"], nparts), LocalTableScanExec(batch))

    def timed(fn, reps=5, warm=2):
        best = None
        for i in range(warm + reps):
            dist.barrier()
            torch.cuda.synchronize()
            stream.record_start()
            out = fn()
            stream.record_stop()
            ms = torch.tensor([stream.elapsed_ms()], device="cuda")
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)          # a collective step is as slow as its slowest rank
            if hasattr(out, "close"):
                out.close()
            if i >= warm and (best is None or ms.item() < best):
                best = ms.item()
        return best

    whole_ms = timed(lambda: ex.executeColumnar(stream))
    capi.check(lib.sb_profile_enable(1))
    capi.check(lib.sb_profile_reset())
    ex.executeColumnar(stream).close()
    prof = {}
    for name in ("partition_scatter", "a2a_counts", "a2a_transfer"):
        t_, c_ = C.c_double(), C.c_int64()
        capi.check(lib.sb_profile_get(name.encode(), C.byref(t_), C.byref(c_)))
        prof[name] = round(t_.value, 4)
    capi.check(lib.sb_profile_enable(0))
    map_ms = timed(lambda: ex.map_side(batch, stream)[0])
    sent = n * rowbytes * (world - 1) / world          # expected bytes leaving a rank (uniform hash)
    if rank == 0:
        line = {"case": "shuffle exchange hash(l_orderkey) n=%d, %d B/row" % (nparts, rowbytes), "n_gpus": world, "rows_per_gpu": n,
                "ms": whole_ms, "rows_per_s": n * world / (whole_ms / 1e3), "map_side_ms": map_ms,
                "transport_ms": None if map_ms is None else whole_ms - map_ms,
                "nvlink_bytes_per_rank": sent, "rank0_device_ms": prof,
                "nvlink_gbs_per_rank_over_transport": None if map_ms is None else sent / ((whole_ms - map_ms) / 1e3) / 1e9}
        print(json.dumps(line), flush=True)
    capi.check(lib.sb_comm_destroy())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
