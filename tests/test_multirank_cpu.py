"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path — chunk-aligned row-range shards that cover the
table exactly once, and the merge of per-rank partial aggregate states (what llkv_gpu_agg_merge does on the device),
checked against the oracle over the whole table."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    for p in (os.path.join(ROOT, "rust-llkv_b200"), ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from llkv_b200 import gpu, tpch
    from llkv_b200.expr import AggregateKind, AggregateSpec
    from oracle import oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        n = 300_001
        t, snap = tpch.lineitem_table(n, seed=3, with_q1=True, with_mvcc=True)
        lo, hi = gpu.shard_rows(n, world, rank, align=4096)
        ranges = [None] * world
        dist.all_gather_object(ranges, (lo, hi))
        specs = [AggregateSpec("n", AggregateKind.CountStar()), AggregateSpec("s", AggregateKind.Sum(tpch.L_EXTENDEDPRICE, tpch.DEC_15_2)),
                 AggregateSpec("lo", AggregateKind.Min(tpch.L_QUANTITY, tpch.DEC_15_2)), AggregateSpec("hi", AggregateKind.Max(tpch.L_DISCOUNT, tpch.DEC_15_2))]
        rules = ["sum", "sum", "min", "max"]
        part = oracle.aggregate(t, tpch.q1_filter(), specs, snap, group_by=tpch.Q1_GROUP_BY, row_begin=lo, row_end=hi, group_capacity=16)
        parts = [None] * world
        dist.all_gather_object(parts, part)
        merged = gpu.merge_partial_results(parts, rules)
        whole = oracle.aggregate(t, tpch.q1_filter(), specs, snap, group_by=tpch.Q1_GROUP_BY, group_capacity=16)
        out.put((rank, ranges, sorted((k, [v.value for v in vals]) for k, vals in merged),
                 sorted((k, [v.value for v in vals]) for k, vals in whole), n))
    finally:
        dist.destroy_process_group()


def test_two_rank_shards_cover_the_table_and_partials_merge_to_the_whole():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ranges, merged, whole, n in results:
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
            assert a1 == b0 and a0 % 4096 == 0 and b0 % 4096 == 0  # contiguous, chunk aligned
        assert merged == whole


def test_shard_rows_edges():
    sys.path.insert(0, os.path.join(ROOT, "rust-llkv_b200"))
    from llkv_b200 import gpu
    assert gpu.shard_rows(0, 4, 2) == (0, 0)
    assert gpu.shard_rows(10, 1, 0) == (0, 10)
    spans = [gpu.shard_rows(1_000_000, 8, r) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == 1_000_000
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert all(s[0] % 131072 == 0 for s in spans)
    with pytest.raises(ValueError):
        gpu.shard_rows(10, 2, 2)
