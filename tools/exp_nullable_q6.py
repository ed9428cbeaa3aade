"""Q6 at SF10 with 20 % NULLs in l_discount against the dense table: kernel times (within 1.3x)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "rust-llkv_b200"), ROOT]
import numpy as np

from llkv_b200 import gpu, tpch
from llkv_b200.table import pack_validity

n = int(sys.argv[1]) if len(sys.argv) > 1 else tpch.lineitem_rows(10.0)
ctx = gpu.Context(0)
ctx.set_timing(True)
t, _ = tpch.lineitem_table(n, seed=6, with_q1=False)
for label, frac in (("dense", 0.0), ("20% NULL l_discount", 0.2)):
    if frac:
        rng = np.random.default_rng(3)
        t.columns[tpch.L_DISCOUNT].validity = pack_validity(rng.random(n) >= frac)
    dt = gpu.DeviceTable.from_host(ctx, t, chunk_rows=1 << 20)
    prog = gpu.Program(ctx, tpch.q6_filter())
    agg = gpu.Aggregation(dt, tpch.q6_aggregates())
    ms = []
    for i in range(6):
        agg.execute(prog)
        res = agg.finalize(1)
        ms.append(agg.run_info().last_kernel_ms)
    info = agg.run_info()
    print(f"{label}: kernel ms {[round(m, 4) for m in ms[2:]]} lean={info.used_fast_kernel} jit={info.used_jit_kernel} bytes/row={info.physical_bytes_per_row} "
          f"result={res[0][1][0].value}", flush=True)
    agg.destroy()
    prog.destroy()
    dt.destroy()
