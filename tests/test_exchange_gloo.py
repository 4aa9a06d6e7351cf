"""N>1 host logic on CPU: two ranks (gloo, world_size 2) run the exchange planning the way sb_all_to_all does --
per-partition counts all-gathered, contiguous partition ownership from sb_exchange_plan (the library's host-only
planner, no GPU needed), rows regrouped by owner -- and the union of what the ranks receive must equal the global
shuffle of the whole table.  The device data movement itself (NCCL) is covered by the GPU multi-rank bench."""
import ctypes as C
import os
import socket

import numpy as np
import pyarrow as pa
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

NPARTS = 13
WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _table(seed, n):
    rng = np.random.default_rng(seed)
    return pa.table({"k": rng.integers(0, 1000, n), "v": rng.integers(-10 ** 6, 10 ** 6, n)})


def _worker(rank, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as O
    from spark_b200 import _capi as capi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    lib = capi.load()
    shard = _table(100 + rank, 5000 + 700 * rank)
    parted, offs = O.hash_partition(shard, ["k"], NPARTS)                 # map side of this rank
    # host planner of the library: rows this rank sends to every destination
    send = np.zeros(WORLD, np.int64)
    offs64 = np.ascontiguousarray(offs, np.int64)
    assert lib.sb_exchange_plan(offs64.ctypes.data_as(C.POINTER(C.c_int64)), NPARTS, WORLD, send.ctypes.data_as(C.POINTER(C.c_int64))) == 0
    # counts matrix, exactly what sb_all_to_all all-gathers
    counts = torch.from_numpy(np.diff(offs64).copy())
    gathered = [torch.zeros_like(counts) for _ in range(WORLD)]
    dist.all_gather(gathered, counts)
    all_counts = torch.stack(gathered).numpy()
    lo = [-(-r * NPARTS // WORLD) for r in range(WORLD + 1)]
    for d in range(WORLD):
        assert send[d] == all_counts[rank, lo[d]:lo[d + 1]].sum()
    # move the rows: contiguous slice per destination (same slicing the NCCL path uses)
    k = np.asarray(parted.column("k")); v = np.asarray(parted.column("v"))
    recv_k, recv_v = [], []
    for src in range(WORLD):
        for dst in range(WORLD):
            n_rows = int(all_counts[src, lo[dst]:lo[dst + 1]].sum())
            if src == rank:
                b, e = int(offs64[lo[dst]]), int(offs64[lo[dst + 1]])
                assert e - b == n_rows
                payload = torch.from_numpy(np.stack([k[b:e], v[b:e]]).copy())
                if dst == rank:
                    recv_k.append(payload[0].numpy()); recv_v.append(payload[1].numpy())
                else:
                    dist.send(payload, dst)
            elif dst == rank:
                buf = torch.zeros((2, n_rows), dtype=torch.int64)
                dist.recv(buf, src)
                recv_k.append(buf[0].numpy()); recv_v.append(buf[1].numpy())
    rk = np.concatenate(recv_k); rv = np.concatenate(recv_v)
    np.save(os.path.join(out_dir, "k%d.npy" % rank), rk)
    np.save(os.path.join(out_dir, "v%d.npy" % rank), rv)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_equals_global_shuffle(tmp_path):
    from oracle import oracle as O
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    whole = pa.concat_tables([_table(100 + r, 5000 + 700 * r) for r in range(WORLD)])
    pid = O.partition_ids(whole, ["k"], NPARTS)
    lo = [-(-r * NPARTS // WORLD) for r in range(WORLD + 1)]
    wk = np.asarray(whole.column("k")); wv = np.asarray(whole.column("v"))
    total = 0
    for r in range(WORLD):
        rk = np.load(tmp_path / ("k%d.npy" % r)); rv = np.load(tmp_path / ("v%d.npy" % r))
        owned = (pid >= lo[r]) & (pid < lo[r + 1])
        want = sorted(zip(wk[owned].tolist(), wv[owned].tolist()))
        assert sorted(zip(rk.tolist(), rv.tolist())) == want          # rank r holds exactly its partitions' rows
        got_pid = O.partition_ids(pa.table({"k": rk}), ["k"], NPARTS)
        assert got_pid.min() >= lo[r] and got_pid.max() < lo[r + 1]
        total += len(rk)
    assert total == whole.num_rows
