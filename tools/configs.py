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
rng = np.random.default_rng(4)
k = rng.integers(0, keys4, n4, dtype=np.int64)
v = rng.integers(0, 1001, n4, dtype=np.int64)
t = HostTable(1).add(HostColumn(tpch.K_FIELD, DataType.Int64, k)).add(HostColumn(tpch.V_FIELD, DataType.Int64, v))
dt = gpu.DeviceTable.from_host(ctx, t, chunk_rows=1 << 20)
ms, info, _, n_groups = timed(dt, None, tpch.highcard_aggregates(), group_by=(tpch.K_FIELD,), hint=keys4, reps=3)
kk = min(m[0] for m in ms[1:])
print(f"config4 n={n4} keys={keys4}: groups={n_groups} kernel {kk:.3f} ms  {n4 / kk / 1e6:.2f} Grows/s  stream {16 * n4 / kk / 1e6:.0f} GB/s  fast={info.used_fast_kernel} jit={info.used_jit_kernel} launches={info.kernel_launches} partitions={info.partitions} all_ms={[round(m[0], 3) for m in ms]}", flush=True)
assert n_groups == len(np.unique(k))
