# tuning sweep for the lean kernel on Q6 / Q1 (resident), prints kernel ms per configuration
import os, sys, time, itertools
sys.path[:0]=['rust-llkv_b200','.']
import numpy as np
from llkv_b200 import gpu, tpch
from llkv_b200.table import HostColumn
from llkv_b200.expr import DataType
n=int(sys.argv[1]) if len(sys.argv)>1 else 20_000_000
ctx=gpu.Context(0); ctx.set_timing(True)
import os
ctx.set_jit(int(os.environ.get('LLKV_JIT','2')))
t,snap=tpch.lineitem_table(n, seed=6, with_q1=True, with_mvcc=True)
dt=gpu.DeviceTable.from_host(ctx,t,chunk_rows=1<<20)
def run(filter_expr,specs,keys=(),snapshot=None,hint=0,reps=5):
    prog=gpu.Program(ctx,filter_expr); dt.set_snapshot(snapshot)
    agg=gpu.Aggregation(dt,specs,keys,cardinality_hint=hint)
    ms=[]
    for i in range(reps+2):
        agg.reset(); agg.run(prog, snapshot is not None); agg.finalize(16)
        if i>=2: ms.append(agg.run_info().last_kernel_ms)
    info=agg.run_info(); agg.destroy(); prog.destroy()
    return min(ms), info
for nt,r,st,ct in [(0,0,0,0),(128,4,2,2),(128,8,2,1),(128,8,3,1),(128,4,3,1),(128,4,4,1),(64,4,2,4),(64,4,3,3),(64,4,4,3),(64,8,2,3),(64,8,2,2),(64,8,3,2),(64,4,2,3),(128,1,4,4),(128,1,4,3),(128,1,6,3),(256,4,2,1),(256,1,4,2),(256,1,6,2),(256,1,3,2),(96,4,2,3),(96,4,3,2),(96,8,2,2),(32,8,3,6),(32,8,4,4),(32,4,4,6),(192,4,2,1),(192,1,4,2),(160,4,2,2)]:
    ctx.set_tuning(ctas_per_sm=ct, block_threads=nt, stages=st, rows_per_thread=r)
    try:
        m6,i6=run(tpch.q6_filter(),tpch.q6_aggregates())
        m1,i1=run(tpch.q1_filter(),tpch.q1_aggregates(),tpch.Q1_GROUP_BY,snap,6)
        print(f"NT={nt} R={r} stages={st} ctas={ct}: Q6 {m6:.3f} ms ({52*n/m6/1e6:.0f} GB/s) grid={i6.grid} smem={i6.smem_bytes} jit={i6.used_jit_kernel} | Q1 {m1:.3f} ms ({94*n/m1/1e6:.0f} GB/s) grid={i1.grid} smem={i1.smem_bytes} st={i1.stages} tile={i1.rows_per_tile} fg={i1.fast_groups} jit={i1.used_jit_kernel}", flush=True)
    except Exception as e:
        print(nt,r,st,ct,"ERR",e, flush=True)
