# Q6 on a lineitem clustered by l_shipdate (rows arrive in date order): kernel time with and without zone-map pruning.
import sys
sys.path[:0] = ['rust-llkv_b200', '.']
import numpy as np
from llkv_b200 import gpu, tpch
from llkv_b200.expr import DataType
from llkv_b200.table import HostColumn, HostTable, decimal_from_i64

n = int(sys.argv[1]) if len(sys.argv) > 1 else 59_986_052
jitter = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # days of disorder: shipdate = sorted date + U[0, jitter)
ctx = gpu.Context(0)
ctx.set_timing(True)
a = tpch.lineitem_arrays(n, 6, False)
key = a["shipdate"].astype(np.int64)
if jitter:
    key = key - np.random.default_rng(1).integers(0, jitter, n)
order = np.argsort(key, kind="stable")
t = HostTable(1)
t.add(HostColumn(tpch.L_QUANTITY, tpch.DEC_15_2, decimal_from_i64(a["quantity"][order])))
t.add(HostColumn(tpch.L_EXTENDEDPRICE, tpch.DEC_15_2, decimal_from_i64(a["extendedprice"][order])))
t.add(HostColumn(tpch.L_DISCOUNT, tpch.DEC_15_2, decimal_from_i64(a["discount"][order])))
t.add(HostColumn(tpch.L_SHIPDATE, DataType.Date32, a["shipdate"][order]))
dt = gpu.DeviceTable.from_host(ctx, t, chunk_rows=1 << 20)
prog = gpu.Program(ctx, tpch.q6_filter())
res = {}
for mode in (0, 1):
    ctx.set_pruning(mode)
    agg = gpu.Aggregation(dt, tpch.q6_aggregates())
    ms = []
    for i in range(8):
        agg.reset()
        agg.run(prog, False, 0, n)
        r = agg.finalize(1)
        ms.append(agg.run_info().last_kernel_ms)
    info = agg.run_info()
    res[mode] = r[0][1][0].value
    tiles = (n + info.rows_per_tile - 1) // info.rows_per_tile
    print(f"pruning={mode}: kernel {min(ms[3:]):.4f} ms  tiles {tiles} pruned {info.tiles_pruned} ({100.0 * info.tiles_pruned / tiles:.1f} %)  "
          f"launches {info.kernel_launches} jit {info.used_jit_kernel}  {n / min(ms[3:]) / 1e6:.1f} G rows/s", flush=True)
    agg.destroy()
assert res[0] == res[1], res
print("results identical:", res[0])
