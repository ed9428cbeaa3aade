#!/usr/bin/env python
"""bench.py — rows/s and HBM GB/s (fraction of the measured roofline) for TPC-H Q6 (and Q1) over synthetic lineitem.

  python bench.py --gpus N --steps K --warmup W            our arm: the CUDA path through the C ABI
  python bench.py --impl reference --gpus N ...            the reference's CPU path (oracle port), host threads

A step = one pass of scan -> predicate -> MVCC -> aggregate over the batch: reset accumulators, one fused scan, finalize.
N=1 workload: BASELINE.json configs[1], TPC-H Q6 on synthetic lineitem SF10 (59 986 052 rows) resident in HBM.
N>1: every rank holds its own SF10-sized row-range shard (weak scaling); partial states merge over NCCL each step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line: NCCL's own banner ("NCCL version ...") goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "rust-llkv_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

from llkv_b200 import ffi, tpch  # noqa: E402
from llkv_b200.expr import DataType  # noqa: E402
from llkv_b200.table import HostColumn, HostTable  # noqa: E402

METRIC = "tpch_q6_lineitem_rows_per_sec"
UNIT = "rows/s"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def kernel_name(info) -> str:
    if info.used_jit_kernel:
        return "llkv_lean_jit (lean_kernel.cuh specialised on the plan shape by jit.cpp)"
    if info.used_fast_kernel:
        return "llkv::lean_scan_kernel<R> (lean_kernel.cuh, interpreted)"
    return "llkv::scan_kernel<WIDE,R> (general interpreter)"


def pinned_column(gpu, field_id, dtype, values: np.ndarray):
    """Copies `values` into page-locked host memory and returns (HostColumn view, raw ptr, nbytes)."""
    raw = np.ascontiguousarray(values)
    buf, ptr = gpu.pinned_empty(raw.nbytes)
    buf[:] = raw.view(np.uint8).reshape(-1)
    view = buf.view(raw.dtype).reshape(raw.shape)
    return HostColumn(field_id, dtype, view), ptr, raw.nbytes


def build_lineitem_pinned(gpu, n, seed, with_q1):
    from llkv_b200.table import decimal_from_i64
    a = tpch.lineitem_arrays(n, seed, with_q1)
    cols, ptrs = {}, []

    def add(fid, dtype, vals):
        c, p, _ = pinned_column(gpu, fid, dtype, vals)
        cols[fid] = c
        ptrs.append(p)

    add(tpch.L_QUANTITY, tpch.DEC_15_2, decimal_from_i64(a["quantity"]))
    add(tpch.L_EXTENDEDPRICE, tpch.DEC_15_2, decimal_from_i64(a["extendedprice"]))
    add(tpch.L_DISCOUNT, tpch.DEC_15_2, decimal_from_i64(a["discount"]))
    add(tpch.L_SHIPDATE, DataType.Date32, a["shipdate"])
    extra = {}
    if with_q1:
        add(tpch.L_TAX, tpch.DEC_15_2, decimal_from_i64(a["tax"]))
        extra["returnflag"] = a["returnflag"]
        extra["linestatus"] = a["linestatus"]
    return cols, ptrs, extra


def upload_column(dc, col: HostColumn, chunk_bytes: int):
    """ColumnStore::append shape: one append per chunk of ~chunk_bytes straight from the pinned buffer."""
    n = col.n_rows
    width = col.values.dtype.itemsize * (2 if col.dtype.type == ffi.PT_DECIMAL128 else 1)
    rows = max(1, chunk_bytes // width)
    base = col.values.ctypes.data
    for lo in range(0, n, rows):
        m = min(rows, n - lo)
        dc.append_raw(base + lo * width, m, lo)


def run_ours(args):
    import torch
    from llkv_b200 import gpu

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    n = tpch.lineitem_rows(args.sf)
    if args.rows:
        n = args.rows
    ctx = gpu.Context(local, n_streams=4, pinned_bytes=64 << 20)
    if world > 1:
        ids = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], world, rank)
    ctx.set_timing(True)

    with_q1 = not args.no_q1
    cols, ptrs, extra = build_lineitem_pinned(gpu, n, seed=6 + rank, with_q1=with_q1)
    table = gpu.DeviceTable(ctx, 1)
    chunk_bytes = 1 << 20
    dcs = {}
    for fid, c in cols.items():
        dc = gpu.DeviceColumn(ctx, gpu.logical_field_id(1, fid), c)
        dc.reserve(n)
        upload_column(dc, c, chunk_bytes)
        dcs[fid] = dc
        table.columns[fid] = dc
    table.n_rows = n
    snap = None
    if with_q1:
        from llkv_b200.tpch import _utf8_single_char
        for fid, key in ((tpch.L_RETURNFLAG, "returnflag"), (tpch.L_LINESTATUS, "linestatus")):
            table.add_column(_utf8_single_char(fid, extra[key]), chunk_rows=1 << 20)
        c, d, snap = tpch.mvcc_arrays(n, seed=6 + rank)
        table.add_mvcc(HostColumn(0xFFFFFFFF, DataType.UInt64, c), HostColumn(0xFFFFFFFE, DataType.UInt64, d), chunk_rows=1 << 17)
    table.seal()
    ctx.synchronize()

    def barrier():
        torch.cuda.synchronize()
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def bench_query(filter_expr, specs, group_by=(), snapshot=None, hint=0, steps=args.steps, warmup=args.warmup):
        prog = gpu.Program(ctx, filter_expr)
        table.set_snapshot(snapshot)
        agg = gpu.Aggregation(table, specs, group_by, cardinality_hint=hint)
        kernel_ms, launches = [], 0
        result = None

        def step(record):
            nonlocal launches, result
            agg.reset()
            agg.run(prog, snapshot is not None, 0, n)
            if world > 1 and not os.environ.get("LLKV_BENCH_NO_MERGE"):  # (diagnostic switch: cost of the merge alone)
                agg.merge()
            result = agg.finalize(64 if group_by else 1)
            if record:
                info = agg.run_info()
                kernel_ms.append(info.last_kernel_ms)
                launches += info.kernel_launches + 1  # + the accumulator-init kernel of reset()
                if world > 1:
                    launches += world + 1  # table re-init + one merge kernel per rank

        for _ in range(warmup):
            step(False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step(True)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        info = agg.run_info()
        agg.destroy()
        prog.destroy()
        return {"seconds": dt, "kernel_ms": kernel_ms, "launches": launches, "info": info, "result": result}

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    q6 = bench_query(tpch.q6_filter(), tpch.q6_aggregates())
    q1 = None
    if with_q1:
        q1 = bench_query(tpch.q1_filter(), tpch.q1_aggregates(), tpch.Q1_GROUP_BY, snap, hint=4)  # BASELINE.json: "4-group GROUP BY"

    # ---- end to end: host (pinned) buffers -> chunk appends -> fused scan -> result on the host, every step
    e2e_cols = [tpch.L_QUANTITY, tpch.L_EXTENDEDPRICE, tpch.L_DISCOUNT, tpch.L_SHIPDATE]
    h2d = sum(cols[f].values.nbytes for f in e2e_cols)
    prog = gpu.Program(ctx, tpch.q6_filter())
    table.set_snapshot(None)
    agg = gpu.Aggregation(table, tpch.q6_aggregates())

    def e2e_step():
        for f in e2e_cols:
            dcs[f].clear()
            upload_column(dcs[f], cols[f], chunk_bytes)
        for f in e2e_cols:
            dcs[f].seal()
        agg.reset()
        agg.run(prog, False, 0, n)
        if world > 1:
            agg.merge()
        return agg.finalize(1)

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_result = e2e_step()
    barrier()
    e2e_dt = max_over_ranks(time.perf_counter() - t0)
    d2h = 4 + 3 * 6 * 8  # status word + the ungrouped state row (6 words) and its two spare rows, read back by finalize
    agg.destroy()
    prog.destroy()
    # ---- BASELINE.json configs[3] at a bounded size (one GPU only): high-cardinality GROUP BY, 10 M distinct Int64 keys,
    # SUM + COUNT.  Kernel time only (CUDA events around the scan + apply launches): reading 10 M groups back is not part
    # of the hot path.  Both forms: partitioned (default once the group table exceeds L2) and the per-row global table.
    highcard = None
    if world == 1 and not args.no_highcard:
        hc_rows, hc_keys = (1 << 26, 10_000_000) if n >= 20_000_000 else (1 << 22, 2_000_000)
        rng = np.random.default_rng(4)
        hk = rng.integers(0, hc_keys, hc_rows, dtype=np.int64)
        ht = HostTable(2).add(HostColumn(tpch.K_FIELD, DataType.Int64, hk)).add(
            HostColumn(tpch.V_FIELD, DataType.Int64, rng.integers(0, 1001, hc_rows, dtype=np.int64)))
        n_unique = int(np.unique(hk).size)
        hdt = gpu.DeviceTable.from_host(ctx, ht, chunk_rows=1 << 20)
        del hk, ht
        highcard = {"workload": f"GROUP BY over {hc_keys} distinct Int64 keys, SUM + COUNT, {hc_rows} rows (BASELINE.json configs[3] at a bounded size)",
                    "rows": hc_rows, "keys": hc_keys, "unit": UNIT}
        for name, mode in (("partitioned", 2), ("per_row", 0)):
            ctx.set_partitioning(mode)
            hagg = gpu.Aggregation(hdt, tpch.highcard_aggregates(), (tpch.K_FIELD,), cardinality_hint=hc_keys)
            ms = []
            for i in range(4):
                hagg.reset()
                hagg.run(None, False, 0, hc_rows)
                groups = hagg.group_count()
                if i:
                    ms.append(hagg.run_info().last_kernel_ms)
            hinfo = hagg.run_info()
            hagg.destroy()
            kms_hc = statistics.mean(ms)
            # HBM bytes the form has to move per row: the two Int64 columns once; partitioned also writes and re-reads
            # one (key, row id, operand) tuple.  (The per-row form is bound by random sector traffic on the group table,
            # the partitioned one by L2 request rate: DESIGN.md 3.1b.)
            bpr = 16 + (48 if mode else 0)
            highcard[name] = {"kernel_ms": kms_hc, "value": hc_rows / (kms_hc * 1e-3), "groups": groups,
                              "groups_expected": n_unique, "partitions": hinfo.partitions, "launches_per_run": hinfo.kernel_launches,
                              "streamed_bytes_per_row": bpr, "streamed_gbs": bpr * hc_rows / (kms_hc * 1e-3) / 1e9}
        ctx.set_partitioning(1)
        hdt.destroy()
    clocks = sampler.stop() if rank == 0 else None  # sampled across the timed regions (Q6, Q1, end to end, high cardinality)

    if rank == 0:
        peak, peak_src = measured_peak()
        total_rows = n * world
        kms = statistics.mean(q6["kernel_ms"]) if q6["kernel_ms"] else float("nan")
        # bytes per row: `arrow` = Arrow-layout value buffers (Decimal128 = 16 B), `resident` = what the kernel reads from HBM
        # (Decimal128 columns whose values fit i64 are kept as 8 B per row, DESIGN.md "data layout").  The roofline uses the
        # resident bytes, so the fraction can never exceed what the memory system delivered.
        arrow_bpr = q6["info"].algorithmic_bytes_per_row
        alg_bpr = q6["info"].physical_bytes_per_row
        achieved = alg_bpr * n / (kms * 1e-3) / 1e9 if kms == kms and kms > 0 else None
        achieved_arrow = arrow_bpr * n / (kms * 1e-3) / 1e9 if kms == kms and kms > 0 else None
        traffic = None  # dram__bytes_read + dram__bytes_write of the Q6 kernel from the committed ncu capture, scaled per row
        prof = os.path.join(ROOT, "profiles", "r01_q6_traffic.json")
        if os.path.exists(prof):
            try:
                t = json.load(open(prof))
                traffic = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["rows"] * n
            except Exception:
                pass
        line = {
            "metric": METRIC, "value": total_rows * args.steps / q6["seconds"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": q6["seconds"] / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "i128/i64 (Decimal128, Date32)", "data": "synthetic",
            "config": {"workload": f"TPC-H Q6 (filter + SUM(l_extendedprice*l_discount)) on synthetic lineitem SF{args.sf:g}, "
                                   f"{n} rows per GPU, resident in HBM", "rows_per_gpu": n, "sharding": "row-range per rank" if world > 1 else "none",
                       "l2_policy": "inputs larger than L2 (%.2f GB resident per pass vs 126 MB)" % (alg_bpr * n / 1e9), "chunk_bytes": chunk_bytes,
                       "kernel": {"grid": q6["info"].grid, "block": q6["info"].block, "rows_per_tile": q6["info"].rows_per_tile,
                                  "stages": q6["info"].stages, "smem_bytes": q6["info"].smem_bytes, "wide": q6["info"].used_wide_path,
                                  "specialised": q6["info"].used_jit_kernel}},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                         "traffic": traffic, "peak_source": peak_src, "kernel": kernel_name(q6["info"]), "kernel_ms": kms,
                         "basis": "resident bytes: what the kernel reads from HBM (Decimal128(15,2) columns are kept as i64 after seal, "
                                  "DESIGN.md section 2); SURVEY.md section 8(d) counts Arrow-layout bytes, given beside it",
                         "resident_bytes_per_row": alg_bpr, "algorithmic_bytes_per_row": arrow_bpr,
                         "algorithmic_gbs": achieved_arrow},
            "e2e": {"value": total_rows * e2e_steps / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_dt / e2e_steps * 1e3, "steps": e2e_steps},
            "gpu_launches": q6["launches"],
            "clocks": clocks,
            "result": {"q6_revenue_raw": q6["result"][0][1][0].value, "e2e_revenue_raw": e2e_result[0][1][0].value},
        }
        if q1 is not None:
            k1 = statistics.mean(q1["kernel_ms"])
            b1 = q1["info"].physical_bytes_per_row
            a1 = b1 * n / (k1 * 1e-3) / 1e9
            line["q1"] = {"workload": f"TPC-H Q1 (4-group GROUP BY, Decimal SUM/AVG/COUNT) with MVCC on the same lineitem, {n} rows per GPU",
                          "value": total_rows * args.steps / q1["seconds"], "unit": UNIT, "ms_per_step": q1["seconds"] / args.steps * 1e3,
                          "roofline": {"bound": "hbm", "achieved": a1, "peak": peak, "unit": "GB/s", "frac": a1 / peak, "kernel_ms": k1,
                                       "resident_bytes_per_row": b1, "algorithmic_bytes_per_row": q1["info"].algorithmic_bytes_per_row,
                                       "algorithmic_gbs": q1["info"].algorithmic_bytes_per_row * n / (k1 * 1e-3) / 1e9},
                          "groups": len(q1["result"]), "kernel_name": kernel_name(q1["info"]), "kernel": {"grid": q1["info"].grid, "block": q1["info"].block,
                                                                  "rows_per_tile": q1["info"].rows_per_tile, "stages": q1["info"].stages,
                                                                  "smem_bytes": q1["info"].smem_bytes, "fast_groups": q1["info"].fast_groups,
                                                                  "wide": q1["info"].used_wide_path}}
        if highcard is not None:
            line["highcard"] = highcard
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(seed=6)
        emit(line)

    for p in ptrs:
        gpu.pinned_free(p)
    if dist is not None:
        dist.barrier()
        ctx.comm_destroy()
        dist.destroy_process_group()
    ctx.close()


def cpu_baseline(seed: int, target_seconds: float = 6.0):
    """The reference-shaped CPU path (oracle port of the Rust code: one scan per predicate leaf -> bitmaps -> gather in
    64 K windows -> arrow-style temporaries -> scalar accumulators), leaf scans on all host threads, on a bounded sample."""
    from oracle import oracle
    cores = os.cpu_count() or 1
    probe = 1_000_000
    t, _ = tpch.lineitem_table(probe, seed=seed, with_q1=False)
    t0 = time.perf_counter()
    oracle.aggregate(t, tpch.q6_filter(), tpch.q6_aggregates(), n_threads=cores)
    rate = probe / (time.perf_counter() - t0)
    n = int(min(max(rate * target_seconds, probe), 24_000_000))
    t, _ = tpch.lineitem_table(n, seed=seed, with_q1=False)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        oracle.aggregate(t, tpch.q6_filter(), tpch.q6_aggregates(), n_threads=cores)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": n / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"TPC-H Q6 over the first {n} rows of the same synthetic lineitem (seed {seed}); best of 2; oracle/llkv_oracle.c, "
                      f"leaf scans on {cores} threads, the rest single-threaded like the reference"}


def run_reference(args):
    """--impl reference: the reference's own CPU path cannot be built here (pure Rust, no cargo/rustc in the image), so
    this arm times the oracle port of it on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    cores = os.cpu_count() or 1
    n = args.rows or 4_000_000
    t, _ = tpch.lineitem_table(n, seed=6, with_q1=False)
    for _ in range(args.warmup):
        oracle.aggregate(t, tpch.q6_filter(), tpch.q6_aggregates(), n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.aggregate(t, tpch.q6_filter(), tpch.q6_aggregates(), n_threads=cores)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = f"each step = TPC-H Q6 over a {n}-row sample of the synthetic lineitem (seed 6)"
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i128/i64 (Decimal128, Date32)", "data": "synthetic",
        "config": {"workload": f"TPC-H Q6 on synthetic lineitem SF{args.sf:g} — bounded CPU sample: {sample}"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# stdout carries exactly one JSON line.  Native libraries print there too (NCCL's "NCCL version ..." banner is a plain
# printf when NCCL_DEBUG=VERSION), so file descriptor 1 points at stderr while the benchmark runs and the JSON line goes to
# the real stdout.
_REAL_STDOUT = None


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sf", type=float, default=10.0)
    ap.add_argument("--rows", type=int, default=0, help="override the row count per GPU (smoke runs)")
    ap.add_argument("--no-q1", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-highcard", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
