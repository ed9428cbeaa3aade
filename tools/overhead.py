# host-side overhead of one step (reset -> run -> finalize) around a tiny scan: where the microseconds go
import sys, time
sys.path[:0] = ['rust-llkv_b200', '.']
from llkv_b200 import gpu, tpch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = gpu.Context(0)
ctx.set_timing(True)
t, snap = tpch.lineitem_table(n, seed=6, with_q1=True, with_mvcc=True)
dt = gpu.DeviceTable.from_host(ctx, t, chunk_rows=1 << 20)
for name, f, specs, keys, sn, hint, cap in [("q6", tpch.q6_filter(), tpch.q6_aggregates(), (), None, 0, 1),
                                            ("q1", tpch.q1_filter(), tpch.q1_aggregates(), tpch.Q1_GROUP_BY, snap, 6, 16)]:
    prog = gpu.Program(ctx, f)
    dt.set_snapshot(sn)
    agg = gpu.Aggregation(dt, specs, keys, cardinality_hint=hint)
    acc = [0.0, 0.0, 0.0]
    reps = 200
    for i in range(reps + 5):
        t0 = time.perf_counter(); agg.reset()
        t1 = time.perf_counter(); agg.run(prog, sn is not None)
        t2 = time.perf_counter(); agg.finalize(cap)
        t3 = time.perf_counter()
        if i >= 5:
            acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2
    info = agg.run_info()
    print(f"{name}: reset {acc[0]/reps*1e6:.1f} us  run {acc[1]/reps*1e6:.1f} us  finalize {acc[2]/reps*1e6:.1f} us  (kernel {info.last_kernel_ms*1e3:.1f} us, jit={info.used_jit_kernel})", flush=True)
    agg.destroy(); prog.destroy()
