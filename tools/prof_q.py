# profiling driver: runs Q6 then Q1 (resident) a few times; used under ncu (see profiles/*.md for the command lines)
import sys
sys.path[:0] = ['rust-llkv_b200', '.']
from llkv_b200 import gpu, tpch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tune = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 and sys.argv[3] != '-' else None  # block,R,stages,ctas
hint1 = int(sys.argv[4]) if len(sys.argv) > 4 else 6
ctx = gpu.Context(0)
ctx.set_timing(True)
import os
ctx.set_jit(int(os.environ.get('LLKV_JIT', '2')))
if tune:
    ctx.set_tuning(block_threads=tune[0], rows_per_thread=tune[1], stages=tune[2], ctas_per_sm=tune[3])
t, snap = tpch.lineitem_table(n, seed=6, with_q1=True, with_mvcc=True)
dt = gpu.DeviceTable.from_host(ctx, t, chunk_rows=1 << 20)
for name, f, specs, keys, sn, hint in [("q6", tpch.q6_filter(), tpch.q6_aggregates(), (), None, 0),
                                       ("q1", tpch.q1_filter(), tpch.q1_aggregates(), tpch.Q1_GROUP_BY, snap, hint1)]:
    prog = gpu.Program(ctx, f)
    dt.set_snapshot(sn)
    agg = gpu.Aggregation(dt, specs, keys, cardinality_hint=hint)
    for i in range(reps):
        agg.reset()
        agg.run(prog, sn is not None)
        agg.finalize(16)
        info = agg.run_info()
        print(name, i, "%.3f ms" % info.last_kernel_ms, "grid", info.grid, "block", info.block, "tile", info.rows_per_tile, "stages", info.stages,
              "smem", info.smem_bytes, "fast", info.used_fast_kernel, "jit", info.used_jit_kernel, flush=True)
    agg.destroy()
    prog.destroy()
