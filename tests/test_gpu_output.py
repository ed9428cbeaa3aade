"""GPU: what comes after the accumulators on the same path — ORDER BY on group keys and
aggregates, HAVING, OFFSET / LIMIT over the finalized rows (llkv-executor/src/lib.rs:5306-5348) and DISTINCT aggregates
(llkv-aggregate/src/lib.rs:103-204; known answer: llkv-sql/tests/aggregate_distinct_tests.rs)."""
import numpy as np
import pytest

import util
from llkv_b200 import ffi, tpch
from llkv_b200.expr import AggregateKind, AggregateSpec, DataType
from llkv_b200.table import HostColumn, HostTable, pack_validity
from oracle import oracle

pytestmark = pytest.mark.gpu


def test_q1_comes_out_in_its_order_by(gpu_ctx):
    """TPC-H Q1 ends with ORDER BY l_returnflag, l_linestatus."""
    from llkv_b200 import gpu
    t, snap = tpch.lineitem_table(200_000, seed=5, with_q1=True, with_mvcc=True)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        want = oracle.aggregate(t, tpch.q1_filter(), tpch.q1_aggregates(), snap, tpch.Q1_GROUP_BY, group_capacity=16)
        prog = gpu.Program(gpu_ctx, tpch.q1_filter())
        dt.set_snapshot(snap)
        agg = gpu.Aggregation(dt, tpch.q1_aggregates(), tpch.Q1_GROUP_BY, cardinality_hint=4)
        agg.run(prog, True)
        util.assert_same_result(agg.finalize(16), want)  # first-appearance order
        agg.set_output(order_by=[("key", 0, False, False), ("key", 1, False, False)])
        util.assert_same_result(agg.finalize(16), sorted(want, key=lambda r: r[0]))
        agg.set_output(order_by=[("key", 1, True, False), ("agg", 7, False, False)])  # linestatus DESC, count_order ASC
        util.assert_same_result(agg.finalize(16), sorted(want, key=lambda r: (-ord(r[0][1]), r[1][7].value)))
        # HAVING count_order > median AND l_returnflag <> 'N', ORDER BY sum_qty DESC LIMIT 1
        counts = sorted(r[1][7].value for r in want)
        agg.set_output(having=[("agg", 7, ffi.CMP_GT, counts[1]), ("key", 0, ffi.CMP_NE, "N")], order_by=[("agg", 0, True, False)], limit=1)
        keep = [r for r in want if r[1][7].value > counts[1] and r[0][0] != "N"]
        util.assert_same_result(agg.finalize(16), sorted(keep, key=lambda r: -r[1][0].value)[:1])
        agg.set_output(order_by=[("key", 0, False, False), ("key", 1, False, False)], offset=1, limit=2)
        util.assert_same_result(agg.finalize(2), sorted(want, key=lambda r: r[0])[1:3])
        agg.set_output()
        util.assert_same_result(agg.finalize(16), want)
        agg.destroy()
        prog.destroy()
    finally:
        dt.destroy()


def test_order_by_puts_nulls_where_asked(gpu_ctx):
    from llkv_b200 import gpu
    rng = np.random.default_rng(8)
    n = 50_000
    k = rng.integers(-20, 20, n, dtype=np.int64)
    v = rng.integers(-1000, 1000, n, dtype=np.int64)
    t = HostTable(1).add(HostColumn(1, DataType.Int64, k, pack_validity(rng.random(n) > 0.05)))
    t.add(HostColumn(2, DataType.Int64, v, pack_validity((k % 7 != 0) | (rng.random(n) > 0.999))))
    specs = [AggregateSpec("s", AggregateKind.Sum(2, DataType.Int64)), AggregateSpec("c", AggregateKind.CountStar())]
    want = oracle.aggregate(t, None, specs, None, (1,), group_capacity=64)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        agg = gpu.Aggregation(dt, specs, (1,), cardinality_hint=64)
        agg.run(None)
        for desc in (False, True):
            for nulls_first in (False, True):
                agg.set_output(order_by=[("key", 0, desc, nulls_first)])
                nn = sorted([r for r in want if r[0][0] is not None], key=lambda r: r[0][0], reverse=desc)
                nulls = [r for r in want if r[0][0] is None]
                util.assert_same_result(agg.finalize(64), nulls + nn if nulls_first else nn + nulls)
        agg.set_output(order_by=[("agg", 0, False, True), ("key", 0, True, False)])
        got = agg.finalize(64)
        sums = [r[1][0].value for r in got]
        assert sums == sorted(sums, key=lambda x: (x is not None, x if x is not None else 0))
        assert sorted(map(repr, got)) == sorted(map(repr, want))
        agg.destroy()
    finally:
        dt.destroy()


def test_distinct_aggregates(gpu_ctx):
    from llkv_b200 import gpu
    # llkv-sql/tests/aggregate_distinct_tests.rs:8-49: INSERT (1), (1), (2); SUM(DISTINCT val) = 3
    t = HostTable(1).add(HostColumn(1, DataType.Int64, np.array([1, 1, 2], dtype=np.int64)))
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    specs = [AggregateSpec("s", AggregateKind.Sum(1, DataType.Int64, distinct=True)), AggregateSpec("c", AggregateKind.Count(1, distinct=True)),
             AggregateSpec("a", AggregateKind.Avg(1, DataType.Int64, distinct=True)), AggregateSpec("t", AggregateKind.Total(1, DataType.Int64, distinct=True))]
    try:
        got = dt.aggregate(None, specs)[0][1]
        assert [v.value for v in got] == [3, 2, 1.5, 3.0]
    finally:
        dt.destroy()
    # many values, NULLs, a filter: against numpy
    rng = np.random.default_rng(2)
    n = 400_000
    x = rng.integers(-50_000, 50_000, n, dtype=np.int64)
    valid = rng.random(n) > 0.1
    y = rng.integers(0, 100, n, dtype=np.int64)
    t = HostTable(1).add(HostColumn(1, DataType.Int64, x, pack_validity(valid))).add(HostColumn(2, DataType.Int64, y))
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        for flt, mask in ((None, valid), (tpch.between_filter(2, 10, 19), valid & (y >= 10) & (y <= 19))):
            u = np.unique(x[mask])
            got = dt.aggregate(flt, specs, cardinality_hint=100_000)[0][1]
            assert got[0].value == int(u.sum()) and got[1].value == u.size
            assert abs(got[2].value - u.sum() / u.size) <= 1e-12 * abs(u.sum() / u.size) and got[3].value == float(u.sum())
        got = dt.aggregate(tpch.between_filter(2, 1000, 2000), specs)[0][1]  # nothing selected
        assert [v.value for v in got] == [None, 0, None, 0.0]
    finally:
        dt.destroy()
