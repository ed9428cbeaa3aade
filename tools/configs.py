# BASELINE.json configs 1 and 4 on one GPU (resident): kernel time, rows/s, GB/s.  Results are checked against numpy.
import sys, time
sys.path[:0] = ['rust-llkv_b200', '.']
import numpy as np
from llkv_b200 import gpu, tpch
from llkv_b200.expr import AggregateKind, AggregateSpec, DataType
from llkv_b200.table import HostColumn, HostTable

n4 = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
keys4 = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
ctx = gpu.Context(0)
ctx.set_timing(True)


def timed(dt, expr, specs, snap=None, group_by=(), hint=0, reps=4, cap=None):
    prog = gpu.Program(ctx, expr) if expr is not None else None
    dt.set_snapshot(snap)
    agg = gpu.Aggregation(dt, specs, group_by, cardinality_hint=hint)
    ms, res = [], None
    for i in range(reps):
        agg.reset()
        agg.run(prog, snap is not None)
        t0 = time.perf_counter()
        n_groups = agg.group_count() if group_by else len(agg.finalize(1))
        ms.append((agg.run_info().last_kernel_ms, (time.perf_counter() - t0) * 1e3))
    info = agg.run_info()
    if cap:
        res = agg.finalize(cap)
    agg.destroy()
    if prog:
        prog.destroy()
    return ms, info, res, n_groups


# ---- config 1: SELECT SUM(x) FROM t WHERE x BETWEEN a AND b, single Int64 column, 10 M rows, MVCC columns present
only4 = len(sys.argv) > 3 and sys.argv[3] == 'only4'
for n in (() if only4 else (10_000_000, 1_000_000_000 // 4)):
    t, snap = tpch.int64_table(n, seed=1)
    x = t.columns[tpch.X_FIELD].values
    a, b = np.percentile(x[:1_000_000], [25, 75]).astype(np.int64)
    dt = gpu.DeviceTable.from_host(ctx, t, chunk_rows=1 << 20)
    for with_mvcc in (False, True):
        ms, info, res, _ = timed(dt, tpch.between_filter(tpch.X_FIELD, int(a), int(b)), tpch.sum_int64(tpch.X_FIELD), snap if with_mvcc else None, cap=1)
        want = int(x[(x >= a) & (x <= b)].sum())
        assert res[0][1][0].value == want, (res, want)
        k = min(m[0] for m in ms[1:])
        bpr = info.physical_bytes_per_row
        print(f"config1 n={n} mvcc={with_mvcc}: kernel {k:.4f} ms  {n / k / 1e6:.1f} Grows/s  {bpr} B/row -> {bpr * n / k / 1e6:.0f} GB/s  jit={info.used_jit_kernel} grid={info.grid}", flush=True)
    dt.destroy()
    del t, x

# ---- config 4: high-cardinality GROUP BY (keys4 distinct Int64 keys, SUM + COUNT) over n4 rows
# (the table is generated and uploaded in pieces of 2^26 rows: the 1 B-row configuration is 16 GB of input)
rng = np.random.default_rng(4)
piece = 1 << 26
dt = gpu.DeviceTable(ctx, 1)
cols = {}
seen = np.zeros(keys4, dtype=bool)
for lo in range(0, n4, piece):
    m = min(piece, n4 - lo)
    k = rng.integers(0, keys4, m, dtype=np.int64)
    v = rng.integers(0, 1001, m, dtype=np.int64)
    seen[k] = True
    for fid, arr in ((tpch.K_FIELD, k), (tpch.V_FIELD, v)):
        hc = HostColumn(fid, DataType.Int64, arr)
        if fid not in cols:
            cols[fid] = gpu.DeviceColumn(ctx, gpu.logical_field_id(1, fid), hc)
            cols[fid].reserve(n4)
            dt.columns[fid] = cols[fid]
        cols[fid].append(hc, 1 << 20, row_id_base=lo)
dt.n_rows = n4
dt.seal()
n_unique = int(seen.sum())
del k, v, seen
for part_mode in ((0, 1) if len(sys.argv) > 4 and sys.argv[4] == 'both' else (1,)):
    ctx.set_partitioning(part_mode)
    ms, info, _, n_groups = timed(dt, None, tpch.highcard_aggregates(), group_by=(tpch.K_FIELD,), hint=keys4, reps=3)
    kk = min(m[0] for m in ms[1:])
    print(f"config4 n={n4} keys={keys4} partitioning={part_mode}: groups={n_groups} kernel {kk:.3f} ms  {n4 / kk / 1e6:.2f} Grows/s  stream {16 * n4 / kk / 1e6:.0f} GB/s  fast={info.used_fast_kernel} jit={info.used_jit_kernel} launches={info.kernel_launches} partitions={info.partitions} all_ms={[round(m[0], 3) for m in ms]}", flush=True)
    assert n_groups == n_unique
