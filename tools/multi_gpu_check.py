"""Run under torchrun on N GPUs: every rank scans its row-range shard, partial states merge over NCCL
(llkv_gpu_agg_merge), and every rank's merged result must equal the oracle over the whole table."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "rust-llkv_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import util
from llkv_b200 import gpu, tpch
from llkv_b200.expr import AggregateKind, AggregateSpec, DataType, Operator, pred
from llkv_b200.table import HostColumn, HostTable
from oracle import oracle


def shard_table(t: HostTable, lo: int, hi: int) -> HostTable:
    s = HostTable(t.table_id)
    for c in t.columns.values():
        if c.dtype.type == 12:  # Utf8: single-char columns of the generator
            s.add(HostColumn(c.field_id, c.dtype, np.arange(hi - lo + 1, dtype=np.int32), aux=c.aux[lo:hi].copy()))
        else:
            s.add(HostColumn(c.field_id, c.dtype, c.values[lo:hi].copy()))
    if t.created_by is not None:
        s.add_mvcc(t.created_by.values[lo:hi].copy(), t.deleted_by.values[lo:hi].copy())
    return s


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = gpu.Context(local)
    ids = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0], world, rank)
    n = 400_000
    full, snap = tpch.lineitem_table(n, seed=6, with_q1=True, with_mvcc=True)
    lo, hi = tpch.shard_range(n, world, rank, align=4096)
    dt = from_host_at(ctx, shard_table(full, lo, hi), lo)  # the shard keeps its rows' ids: first-appearance order is global
    dt.set_snapshot(snap)
    hc = tpch.highcard_table(300_000, 40_000, seed=4)
    hlo, hhi = tpch.shard_range(300_000, world, rank, align=4096)
    hdt = from_host_at(ctx, _retable(shard_table(hc, hlo, hhi), 2), hlo)
    cases = [
        (dt, full, tpch.q6_filter(), tpch.q6_aggregates(), (), snap, 0, True),
        (dt, full, tpch.q1_filter(), tpch.q1_aggregates(), tpch.Q1_GROUP_BY, snap, 6, True),
        (hdt, hc, None, tpch.highcard_aggregates(), (tpch.K_FIELD,), None, 40_000, False),
    ]
    cases.append(cases[2] + ("partitioned",))  # the same aggregate through the partitioned (two-pass) form
    cases.append((dt, full, tpch.q6_filter(), tpch.q6_aggregates(), (), snap, 0, True, "pruned"))  # zone-map tile lists on
    for table, host, flt, specs, keys, sn, hint, ordered, *mode in cases:
        ctx.set_partitioning(2 if "partitioned" in mode else 1)
        ctx.set_pruning(2 if "pruned" in mode else 1)
        ctx.set_jit(2 if mode else 1)
        prog = gpu.Program(ctx, flt) if flt is not None else None
        agg = gpu.Aggregation(table, specs, keys, cardinality_hint=hint)
        agg.run(prog, sn is not None)
        agg.merge()
        got = agg.finalize(1 << 17)
        if "partitioned" in mode:
            assert agg.run_info().partitions >= 2
        want = oracle.aggregate(host, flt, specs, sn, keys, group_capacity=1 << 17)
        util.assert_same_result(got, want, 1e-12, ordered=ordered)
        agg.destroy()
        if prog:
            prog.destroy()
    # one call per step (llkv_gpu_agg_execute): the merge is one kernel over the NVLink peer mailboxes for the ungrouped state
    # and for Q1's small group table alike, and from the fourth unchanged step on the whole step replays as one CUDA graph
    ctx.set_partitioning(1)
    ctx.set_pruning(1)
    ctx.set_jit(1)
    for flt, specs, keys, hint in ((tpch.q6_filter(), tpch.q6_aggregates(), (), 0), (tpch.q1_filter(), tpch.q1_aggregates(), tpch.Q1_GROUP_BY, 4)):
        want = oracle.aggregate(full, flt, specs, snap, keys, group_capacity=64)
        prog = gpu.Program(ctx, flt)
        agg = gpu.Aggregation(dt, specs, keys, cardinality_hint=hint)
        for step in range(8):
            agg.execute(prog, True, merge=True)
            util.assert_same_result(agg.finalize(64), want, 1e-12)
        info = agg.run_info()
        if not os.environ.get("LLKV_GPU_NO_P2P_MERGE"):
            assert info.merged_p2p == 1, "the merge left the peer-mailbox path"
        assert info.graph_replays >= 3, info.graph_replays
        agg.destroy()
        prog.destroy()
    # a hint that is too low: the rank tables grow on their own and the fold of the union outgrows a rank's table
    small = gpu.Aggregation(hdt, tpch.highcard_aggregates(), (tpch.K_FIELD,), cardinality_hint=20)
    small.execute(None, False, merge=True)
    got = small.finalize(1 << 17)
    want = oracle.aggregate(hc, None, tpch.highcard_aggregates(), None, (tpch.K_FIELD,), group_capacity=1 << 17)
    util.assert_same_result(got, want, 1e-12, ordered=False)
    small.destroy()
    # GROUP BY over a Utf8 column with strings longer than 7 bytes: every rank interns its own shard's strings, rank 1 sees short
    # ones only (a packed column until the ranks agree); the ranks exchange their dictionaries before the plan is compiled
    long_words = ["DELIVER IN PERSON", "TAKE BACK RETURN", "COLLECT COD", "NONE", "AIR", "x"]
    rng = np.random.default_rng(21)
    m = 60_000
    pick = rng.integers(0, len(long_words), m)
    slo, shi = tpch.shard_range(m, world, rank, align=4096)
    if world > 1:
        r1lo, r1hi = tpch.shard_range(m, world, 1, align=4096)
        pick[r1lo:r1hi] = 3 + pick[r1lo:r1hi] % 3  # rank 1: "NONE", "AIR", "x"
    strs = [long_words[i] for i in pick]
    vals = rng.integers(-100, 100, m, dtype=np.int64)
    whole = HostTable(3).add(HostColumn.utf8(1, strs)).add(HostColumn(2, DataType.Int64, vals))
    part = HostTable(3).add(HostColumn.utf8(1, strs[slo:shi])).add(HostColumn(2, DataType.Int64, vals[slo:shi].copy()))
    sdt = from_host_at(ctx, part, slo)
    sspecs = [AggregateSpec("c", AggregateKind.CountStar()), AggregateSpec("s", AggregateKind.Sum(2, DataType.Int64))]
    sagg = gpu.Aggregation(sdt, sspecs, (1,), cardinality_hint=8)
    flt = pred(1, Operator.GreaterThanOrEquals("COLLECT COD"))
    sprog = gpu.Program(ctx, flt)
    want = oracle.aggregate(whole, flt, sspecs, None, (1,))
    for step in range(5):
        sagg.execute(sprog, False, merge=True)
        util.assert_same_result(sagg.finalize(16), want, 1e-12)
    sagg.destroy()
    sprog.destroy()
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU merge ok on {world} ranks: Q6 (also with tile lists), Q1 and a 40k-group hash aggregate (per-row and partitioned) match the oracle; long-string group keys merge through agreed dictionaries; "
              f"execute() steps merge over peer mailboxes and replay as CUDA graphs", flush=True)
    ctx.comm_destroy()
    dist.destroy_process_group()


def from_host_at(ctx, t: HostTable, row_id_base: int):
    """DeviceTable.from_host with the shard's row ids starting at row_id_base."""
    dt = gpu.DeviceTable(ctx, t.table_id)
    for col in t.columns.values():
        dc = gpu.DeviceColumn(ctx, gpu.logical_field_id(t.table_id, col.field_id), col)
        dc.reserve(col.n_rows)
        dc.append(col, row_id_base=row_id_base)
        dt.columns[col.field_id] = dc
        dt.n_rows = col.n_rows
    if t.created_by is not None:
        dt.created_by = gpu.DeviceColumn(ctx, gpu.logical_field_id(t.table_id, 0xFFFFFFFF, gpu.NS_TXN_CREATED_BY), t.created_by)
        dt.created_by.append(t.created_by, row_id_base=row_id_base)
        dt.deleted_by = gpu.DeviceColumn(ctx, gpu.logical_field_id(t.table_id, 0xFFFFFFFE, gpu.NS_TXN_DELETED_BY), t.deleted_by)
        dt.deleted_by.append(t.deleted_by, row_id_base=row_id_base)
    dt.seal()
    return dt


def _retable(t: HostTable, table_id: int) -> HostTable:
    t.table_id = table_id
    return t


if __name__ == "__main__":
    main()
