# correctness + time of Q1 / Q6 under explicit geometries (llkv_gpu_ctx_set_tuning): block,R,stages,ctas per argument
import sys
sys.path[:0] = ['rust-llkv_b200', '.', 'tests']
import util
from llkv_b200 import gpu, tpch
from oracle import oracle
n = int(sys.argv[1])
ctx = gpu.Context(0)
ctx.set_timing(True)
ctx.set_jit(2)
t, snap = tpch.lineitem_table(n, seed=6, with_q1=True, with_mvcc=True)
dt = gpu.DeviceTable.from_host(ctx, t, chunk_rows=1 << 20)
want = {"q1": oracle.aggregate(t, tpch.q1_filter(), tpch.q1_aggregates(), snap, tpch.Q1_GROUP_BY, group_capacity=16),
        "q6": oracle.aggregate(t, tpch.q6_filter(), tpch.q6_aggregates())} if n <= 2_000_000 else None
for arg in sys.argv[2:]:
    b, r, s, c = [int(x) for x in arg.split(',')]
    ctx.set_tuning(block_threads=b, rows_per_thread=r, stages=s, ctas_per_sm=c)
    for name, f, specs, keys, sn, hint in [("q6", tpch.q6_filter(), tpch.q6_aggregates(), (), None, 0),
                                           ("q1", tpch.q1_filter(), tpch.q1_aggregates(), tpch.Q1_GROUP_BY, snap, 4)]:
        prog = gpu.Program(ctx, f)
        dt.set_snapshot(sn)
        agg = gpu.Aggregation(dt, specs, keys, cardinality_hint=hint)
        ms = []
        for i in range(3):
            agg.reset()
            agg.run(prog, sn is not None)
            got = agg.finalize(16)
            ms.append(round(agg.run_info().last_kernel_ms, 4))
        info = agg.run_info()
        ok = "-"
        if want:
            try:
                util.assert_same_result(got, want[name], 1e-12)
                ok = "ok"
            except AssertionError as e:
                ok = "WRONG " + str(e)[:80]
        print(f"{arg} {name}: ms={ms} grid={info.grid} block={info.block} tile={info.rows_per_tile} stages={info.stages} smem={info.smem_bytes} "
              f"launches={info.kernel_launches} jit={info.used_jit_kernel} fg={info.fast_groups} result={ok}", flush=True)
        agg.destroy()
        prog.destroy()
