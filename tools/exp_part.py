# experiments on the partitioned GROUP BY: kernel time against key count, partitioning mode and slice size
import os, sys
sys.path[:0] = ['rust-llkv_b200', '.']
import numpy as np
from llkv_b200 import gpu, tpch
from llkv_b200.expr import DataType
from llkv_b200.table import HostColumn, HostTable

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
ctx = gpu.Context(0)
ctx.set_timing(True)
ctx.set_jit(2)
for keys in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "250000,2000000,10000000").split(",")]:
    rng = np.random.default_rng(4)
    k = rng.integers(0, keys, n, dtype=np.int64)
    v = rng.integers(0, 1001, n, dtype=np.int64)
    t = HostTable(1).add(HostColumn(tpch.K_FIELD, DataType.Int64, k)).add(HostColumn(tpch.V_FIELD, DataType.Int64, v))
    dt = gpu.DeviceTable.from_host(ctx, t, chunk_rows=1 << 20)
    del k, v, t
    for mode, slice_mb in ((0, 24), (2, 24), (2, 6), (2, 96)):
        os.environ["LLKV_GPU_PART_SLICE_MB"] = str(slice_mb)
        ctx.set_partitioning(mode)
        agg = gpu.Aggregation(dt, tpch.highcard_aggregates(), (tpch.K_FIELD,), cardinality_hint=keys)
        ms = []
        for i in range(3):
            agg.reset()
            agg.run(None, False)
            g = agg.group_count()
            ms.append(round(agg.run_info().last_kernel_ms, 3))
        info = agg.run_info()
        print(f"keys={keys} mode={mode} slice_mb={slice_mb}: partitions={info.partitions} groups={g} ms={ms} -> {n / min(ms) / 1e6:.1f} Grows/s", flush=True)
        agg.destroy()
    dt.destroy()
