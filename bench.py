#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Spark SQL hot path.

Workload (BASELINE.json configs[1]): TPC-H Q1 at SF10 per GPU -- scan + filter + hash aggregate over
lineitem (59,986,052 rows x 38 B/row = 2.28 GB of referenced columns, >> the 126 MB L2, so no L2 flush is
needed between timed iterations), then Final aggregate + sort of the 4 result rows.  Weak scaling: every
rank owns its own SF10 lineitem shard; Partial results meet through sb_all_gather (the SinglePartition
exchange of a 4-row table), every rank finishes Final + Sort.

One JSON line on stdout (rank 0):
  value      rows/s with the columns already resident in HBM (device-timed, max over ranks)
  e2e        rows/s through the public plan API with HOST (pinned) column buffers: H2D of the 7 columns,
             the plan, D2H of the result, all inside the timed region
  roofline   agg_update kernel: algorithmic bytes (38 B/row) / its CUDA-event duration vs the measured HBM peak
  cpu_baseline  the oracle's whole-stage restatement of the same stage timed on this host's cores

`--impl reference` times the reference's CPU path for the same workload.  The reference itself is JVM-only
and no JDK exists in this image, so this arm runs the oracle's C restatement of the
whole-stage-codegen loop (kind "port") on all host threads.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

Q1_ROWS_SF10 = 59_986_052


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"     # B200_PROFILING.md fallback


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host thread it can."""
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(ncpu)


def cpu_q1(cols, n, reps):
    """Times the oracle's whole-stage restatement of Q1 (Partial per thread + Final merge)."""
    _use_all_host_threads()
    from oracle import oracle as O
    from spark_b200 import tpch
    L = O.lib()
    k0 = np.zeros(16, np.int8); k1 = np.zeros(16, np.int8); sums = np.zeros(80); cnt = np.zeros(16, np.int64)
    args = [cols[c].ctypes.data for c in ("l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag",
                                          "l_linestatus", "l_shipdate")]
    def one_pass():
        t0 = time.perf_counter()
        L.so_q1_partial_final(*args, n, tpch.Q1_CUTOFF, 16, k0.ctypes.data, k1.ctypes.data, sums.ctypes.data, cnt.ctypes.data)
        return time.perf_counter() - t0
    # untimed calibration: the scan is DRAM-bound on the host, and on SMT boxes one thread per core is
    # sometimes faster than one per hardware thread -- keep whichever the CPU does better with
    hw = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    best = None
    for nt in sorted({hw, max(1, hw // 2)}, reverse=True):
        L.so_set_threads(nt)
        one_pass()
        t = min(one_pass(), one_pass())
        if best is None or t < best[0]:
            best = (t, nt)
    L.so_set_threads(best[1])
    times = [one_pass() for _ in range(reps)]
    return times, L.so_threads()


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (oracle port) on the same workload, host cores only."""
    if rank != 0:
        return
    from spark_b200 import tpch
    n = int(Q1_ROWS_SF10 * args.sf / 10)
    cols = tpch.lineitem_q1_columns(n, seed=42)
    times, threads = cpu_q1(cols, n, args.warmup + args.steps)
    t = times[args.warmup:]
    ms = 1000.0 * sum(t) / len(t)
    value = n / (ms / 1000.0)
    line = {"impl": "reference", "metric": "tpch_q1_rows_per_sec", "value": value, "unit": "rows/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "TPC-H Q1 SF%g (scan+filter+hash-agg on lineitem), CPU whole-stage restatement" % args.sf,
                       "rows": n, "bytes_per_row": tpch.Q1_BYTES_PER_ROW},
            "cpu_baseline": {"value": value, "unit": "rows/s", "cores": threads, "kind": "port",
                             "sample": "full SF%g lineitem (%d rows) per step" % (args.sf, n)},
            "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sf", type=float, default=10.0, help="TPC-H scale factor per GPU (default 10 = configs[1])")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from spark_b200 import _capi as capi, tpch
    from spark_b200.columnar import ColumnarBatch, HostColumn, PinnedArray, Stream
    from spark_b200.execution import LocalTableScanExec, SparkPlan

    lib = capi.init(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idbuf = torch.zeros(capi.SB_UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = C.create_string_buffer(capi.SB_UNIQUE_ID_BYTES)
            capi.check(lib.sb_comm_get_unique_id(raw))
            idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        capi.check(lib.sb_comm_init(rank, world, bytes(idbuf.cpu().numpy().tobytes())))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = Stream()
    n = int(Q1_ROWS_SF10 * args.sf / 10)
    # ---- synthetic lineitem (Q1 columns) generated straight into pinned host buffers ------------------
    pinned = {name: PinnedArray(n, dt) for name, dt in tpch.Q1_DTYPES.items()}
    cols = tpch.lineitem_q1_columns(n, seed=42 + rank, out={k: v.array for k, v in pinned.items()})
    names = list(tpch.Q1_DTYPES)
    sb_types = {"l_quantity": capi.SB_FLOAT64, "l_extendedprice": capi.SB_FLOAT64, "l_discount": capi.SB_FLOAT64,
                "l_tax": capi.SB_FLOAT64, "l_returnflag": capi.SB_INT8, "l_linestatus": capi.SB_INT8, "l_shipdate": capi.SB_DATE32}
    host_cols = [HostColumn(sb_types[c], cols[c]) for c in names]

    def import_batch():
        return ColumnarBatch.from_host_columns(names, host_cols, stream)

    class AllGatherExec(SparkPlan):
        """SinglePartition exchange of the (tiny) Partial output: every rank receives all partial rows."""

        def __init__(self, child):
            self.child = child

        def executeColumnar(self, stream=None):
            inp = self.child.executeColumnar(stream)
            try:
                h = C.c_void_p()
                capi.check(lib.sb_all_gather(inp.handle, stream.handle, C.byref(h)))
                return ColumnarBatch(h, inp.names, inp.arrow_types)
            finally:
                inp.close()

    class BatchSource(SparkPlan):
        """Leaf whose batch is swapped per step, so the operator tree (and its compiled plans) is built once."""

        def __init__(self):
            self.batch = None

        def executeColumnar(self, stream=None):
            return self.batch.rename(self.batch.names)

    source = BatchSource()
    partial_plan = tpch.q1_partial_plan(source, fused=True)
    if world > 1:
        partial_plan = AllGatherExec(partial_plan)
    q1_plan = tpch.q1_final_plan(partial_plan, sort=True)

    def q1(batch):
        source.batch = batch
        return q1_plan

    resident = import_batch()
    stream.synchronize()

    def step_resident():
        out = q1(resident).executeColumnar(stream)
        out.close()

    # ---- device-resident timing ----------------------------------------------------------------------------
    # clocks / throttle reasons are sampled every 20 ms from before the warm-up until the end of the end-to-end loop,
    # i.e. across both timed regions (the device-resident one alone lasts ~0.2 s)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    capi.check(lib.sb_profile_enable(1))
    capi.check(lib.sb_profile_reset())
    barrier()
    launches0 = capi.kernel_launch_count()
    stream.record_start()
    for _ in range(args.steps):
        step_resident()
    stream.record_stop()
    total_ms = stream.elapsed_ms()
    barrier()
    launches = capi.kernel_launch_count() - launches0
    kms, kcount = C.c_double(), C.c_int64()
    capi.check(lib.sb_profile_get(b"agg_update", C.byref(kms), C.byref(kcount)))
    capi.check(lib.sb_profile_enable(0))
    ms_per_step = total_ms / args.steps
    if world > 1:
        t = torch.tensor([ms_per_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item())
    value = world * n / (ms_per_step / 1000.0)

    # the dominant kernel: one agg_update launch over the shard per step (the Final aggregate's 4-row launch is
    # excluded by taking the per-step maximum share: the big launch is > 99.9% of the summed time)
    kernel_ms = kms.value / max(1, args.steps)
    peak, peak_src = _peaks()
    alg_bytes = n * tpch.Q1_BYTES_PER_ROW
    achieved = alg_bytes / (kernel_ms / 1000.0) / 1e9 if kernel_ms > 0 else 0.0

    # ---- end to end: host (pinned) columns -> H2D -> plan -> D2H of the result ----------------------------------
    result = None
    for _ in range(2):
        b = import_batch(); result = q1(b).collect(stream); b.close()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        b = import_batch()
        result = q1(b).collect(stream)
        b.close()
    stream.synchronize()
    e2e_ms = 1000.0 * (time.perf_counter() - t0) / args.e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * n / (e2e_ms / 1000.0)
    d2h = int(sum(result.column(i).nbytes for i in range(result.num_columns)))
    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline (rank 0, N=1 only): the oracle's whole-stage loop on this host's cores -------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        times, threads = cpu_q1(cols, n, 13)
        mean = sum(times[1:]) / len(times[1:])
        cpu = {"value": n / mean, "unit": "rows/s", "cores": threads, "kind": "port",
               "sample": "full SF%g lineitem (%d rows) x 12 passes after 1 warm-up, mean; %.1f core-seconds of CPU work"
                         % (args.sf, n, sum(times[1:]) * threads)}

    if rank == 0:
        line = {"metric": "tpch_q1_rows_per_sec", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "TPC-H Q1 SF%g per GPU (scan+filter+hash-agg on lineitem, Final agg + sort)" % args.sf,
                           "rows_per_gpu": n, "bytes_per_row": tpch.Q1_BYTES_PER_ROW, "plan": "HashAggregate(partial, fused Filter/Project)"
                           + (" -> AllGather" if world > 1 else "") + " -> HashAggregate(final) -> Sort",
                           "l2_policy": "inputs (%.2f GB) larger than L2, no flush" % (alg_bytes / 1e9)},
                "roofline": {"bound": "hbm", "kernel": "agg_update_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "frac_of_8tbs_nominal": achieved / 8000.0, "peak_source": peak_src,
                             "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes,
                             # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at SF10 (ncu --set full,
                             # profiles/r01_agg_update_final_ncu.txt): 2.2796 GB + 3.3 MB; reported only for the configuration it was captured on
                             "traffic": 2282970560 if n == Q1_ROWS_SF10 else None},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "rows/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": alg_bytes,
                        "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    resident.close()
    if world > 1:
        capi.check(lib.sb_comm_destroy())
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
