"""GROUP BY keys that do not pack into 64 bits ("What's missing" 6; GroupKeyValue tuples of any width,
llkv-executor/src/lib.rs:99-106,9362-9456).  The device groups by a 64-bit hash of the keys' images, proves after the run
that no two different keys met in one group, and reads the key values back from the columns at each group's first row.
CPU: the plan lowers (and specialises) with the hashed-key form.  GPU: against the oracle — interpreted, specialised,
partitioned; nullable keys; strings among the keys; HAVING / ORDER BY over the keys."""
import numpy as np
import pytest

import util
from llkv_b200 import gpu
from llkv_b200.expr import AggregateKind, AggregateSpec, CompareOp, DataType, Operator, pred
from llkv_b200.table import HostColumn, HostTable, pack_validity
from oracle import oracle

SPECS = [AggregateSpec("c", AggregateKind.CountStar()), AggregateSpec("s", AggregateKind.Sum(5, DataType.Int64)),
         AggregateSpec("mn", AggregateKind.Min(5, DataType.Int64))]


def wide_table(n, seed, nulls=False, groups=300, tid=71):
    """Keys drawn from a small pool of (a, b, c, d) tuples whose columns span the full 64-bit range."""
    rng = np.random.default_rng(seed)
    pool_a = rng.integers(-(1 << 62), 1 << 62, groups, dtype=np.int64)
    pool_b = rng.integers(0, (1 << 64) - 1, groups, dtype=np.uint64)
    pool_a[:20] = pool_a[0]  # many tuples share a prefix: the second key decides
    pick = rng.integers(0, groups, n)
    words = ["", "A", "xy", "seven77", "DELIVER IN PERSON", "TAKE BACK RETURN"]
    t = HostTable(tid)
    cols = [HostColumn(1, DataType.Int64, pool_a[pick]), HostColumn(2, DataType.UInt64, pool_b[pick]),
            HostColumn(3, DataType.Int32, (pick % 7 - 3).astype(np.int32)),
            HostColumn.utf8(4, [words[i % len(words)] for i in pick]),
            HostColumn(5, DataType.Int64, rng.integers(-1000, 1000, n, dtype=np.int64))]
    if nulls:
        for c in cols[:3]:
            c.validity = pack_validity(rng.random(n) > 0.1)
    for c in cols:
        t.add(c)
    return t


def test_wide_keys_lower_to_the_hashed_form_of_the_lean_kernel():
    t = wide_table(2000, 1)
    text = gpu.debug_plan(t, pred(5, Operator.GreaterThan(-500)), SPECS, group_by=(1, 2, 3), cardinality_hint=512, jit=True)
    assert text.startswith("lean plan") and "specialised cubin" in text
    assert "3 keys" in text


@pytest.mark.gpu
@pytest.mark.parametrize("nulls", [False, True], ids=["dense", "nullable"])
def test_gpu_wide_keys_match_the_oracle(gpu_ctx, nulls):
    t = wide_table(80_000, 3, nulls=nulls)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        for keys in [(1, 2), (2, 1, 3), (1, 2, 4), (4, 2, 1, 3)]:
            for flt in (None, pred(5, Operator.LessThan(300))):
                want = oracle.aggregate(t, flt, SPECS, None, keys, group_capacity=1 << 12)
                for mode, part in ((0, 1), (2, 1), (2, 2)):
                    gpu_ctx.set_jit(mode)
                    gpu_ctx.set_partitioning(part)
                    got = dt.aggregate(flt, SPECS, None, keys, cardinality_hint=1024)
                    util.assert_same_result(got, want)
    finally:
        gpu_ctx.set_jit(1)
        gpu_ctx.set_partitioning(1)
        dt.destroy()


@pytest.mark.gpu
def test_gpu_wide_keys_having_order_by_and_repeated_steps(gpu_ctx):
    t = wide_table(50_000, 5, groups=64, tid=72)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        agg = gpu.Aggregation(dt, SPECS, (1, 2), cardinality_hint=128)
        agg.set_output(having=[("key", 0, CompareOp.Gt, 0)], order_by=[("key", 1, True, False)], limit=10)
        want = [r for r in oracle.aggregate(t, None, SPECS, None, (1, 2)) if r[0][0] > 0]
        want = sorted(want, key=lambda r: r[0][1], reverse=True)[:10]
        for step in range(5):  # execute(): the later steps replay a CUDA graph; the keys come from the columns every time
            agg.execute(None)
            util.assert_same_result(agg.finalize(16), want)
        assert agg.run_info().graph_replays >= 1
        agg.destroy()
    finally:
        dt.destroy()


@pytest.mark.gpu
def test_gpu_wide_keys_survive_table_growth(gpu_ctx):
    """5000 wide-key groups against a hint of 16: the table fills, grows and is rehashed by the (hashed) key; first-row words
    — where the key values are read from — move with their groups."""
    t = wide_table(60_000, 9, groups=5000, tid=73)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        want = oracle.aggregate(t, None, SPECS, None, (2, 1), group_capacity=1 << 14)
        for mode in (0, 2):
            gpu_ctx.set_jit(mode)
            util.assert_same_result(dt.aggregate(None, SPECS, None, (2, 1), cardinality_hint=16), want)
    finally:
        gpu_ctx.set_jit(1)
        dt.destroy()
