# Q6 (and optionally Q1) geometry sweep at SF10 on the final resident layout (16 / 38 B per row): kernel ms per configuration
import os, sys
sys.path[:0] = ['rust-llkv_b200', '.']
from llkv_b200 import gpu, tpch
n = int(sys.argv[1]) if len(sys.argv) > 1 else tpch.lineitem_rows(10)
with_q1 = os.environ.get("Q1", "0") == "1"
ctx = gpu.Context(0)
ctx.set_timing(True)
ctx.set_jit(2)
t, snap = tpch.lineitem_table(n, seed=6, with_q1=with_q1, with_mvcc=with_q1)
dt = gpu.DeviceTable.from_host(ctx, t, chunk_rows=1 << 20)


def run(filter_expr, specs, keys=(), snapshot=None, hint=0, reps=12):
    prog = gpu.Program(ctx, filter_expr)
    dt.set_snapshot(snapshot)
    agg = gpu.Aggregation(dt, specs, keys, cardinality_hint=hint)
    ms = []
    for i in range(reps + 3):
        agg.reset()
        agg.run(prog, snapshot is not None)
        agg.finalize(16)
        if i >= 3:
            ms.append(agg.run_info().last_kernel_ms)
    info = agg.run_info()
    agg.destroy()
    prog.destroy()
    ms.sort()
    return ms[0], ms[len(ms) // 2], info


cfgs = [tuple(int(x) for x in c.split(',')) for c in os.environ['CFGS'].split()] if os.environ.get('CFGS') else [(0, 0, 0, 0), (128, 4, 3, 4), (128, 4, 2, 4), (128, 4, 4, 4), (128, 4, 3, 5), (128, 4, 2, 6), (128, 8, 2, 4), (128, 8, 3, 3), (128, 8, 2, 3), (128, 2, 4, 6), (128, 2, 3, 8),
        (256, 4, 3, 2), (256, 4, 2, 3), (256, 2, 3, 4), (256, 8, 2, 2), (64, 4, 3, 8), (64, 8, 3, 6), (64, 8, 2, 8), (192, 4, 3, 3), (96, 4, 3, 5), (96, 8, 3, 4), (512, 4, 2, 1), (512, 2, 3, 2)]
for nt, r, st, ct in cfgs:
    ctx.set_tuning(ctas_per_sm=ct, block_threads=nt, stages=st, rows_per_thread=r)
    try:
        best, med, i6 = run(tpch.q6_filter(), tpch.q6_aggregates())
        line = f"NT={nt} R={r} stages={st} ctas={ct}: Q6 best {best:.4f} median {med:.4f} ms ({16 * n / med / 1e6:.0f} GB/s) grid={i6.grid} block={i6.block} smem={i6.smem_bytes}"
        if with_q1:
            b1, m1, i1 = run(tpch.q1_filter(), tpch.q1_aggregates(), tpch.Q1_GROUP_BY, snap, 4)
            line += f" | Q1 best {b1:.4f} median {m1:.4f} ms grid={i1.grid} smem={i1.smem_bytes} stages={i1.stages} tile={i1.rows_per_tile}"
        print(line, flush=True)
    except Exception as e:
        print(nt, r, st, ct, "ERR", e, flush=True)
