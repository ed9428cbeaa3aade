import os, sys
sys.path[:0] = ['rust-llkv_b200', '.']
import numpy as np
from llkv_b200 import gpu, tpch
from llkv_b200.expr import DataType
from llkv_b200.table import HostColumn, HostTable
n = int(sys.argv[1]); keys = int(sys.argv[2])
ctx = gpu.Context(0); ctx.set_timing(True); ctx.set_jit(2); ctx.set_partitioning(2)
rng = np.random.default_rng(4)
k = rng.integers(0, keys, n, dtype=np.int64); v = rng.integers(0, 1001, n, dtype=np.int64)
t = HostTable(1).add(HostColumn(tpch.K_FIELD, DataType.Int64, k)).add(HostColumn(tpch.V_FIELD, DataType.Int64, v))
dt = gpu.DeviceTable.from_host(ctx, t, chunk_rows=1 << 20)
for var in sys.argv[3].split(","):
    os.environ["LLKV_GPU_PART_VARIANT"] = var
    agg = gpu.Aggregation(dt, tpch.highcard_aggregates(), (tpch.K_FIELD,), cardinality_hint=keys)
    ms = []
    for i in range(3):
        agg.reset(); agg.run(None, False)
        g = agg.group_count()
        ms.append(round(agg.run_info().last_kernel_ms, 3))
    print(f"keys={keys} variant={var} groups={g} ms={ms}", flush=True)
    agg.destroy()
