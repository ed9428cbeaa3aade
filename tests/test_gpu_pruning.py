"""GPU: zone-map tile skipping (llkv_gpu_ctx_set_pruning) — the device-side form of the reference's chunk pruning
(llkv-column-map/src/store/pruning.rs:104-258).  Whatever is skipped, results equal the oracle's and the unpruned scan's."""
import numpy as np
import pytest

import util
from llkv_b200 import ffi, tpch
from llkv_b200.expr import AggregateKind, AggregateSpec, Bound, DataType, Expr, Literal, Operator, pred
from llkv_b200.table import HostColumn, HostTable, decimal_from_i64
from oracle import oracle

pytestmark = pytest.mark.gpu
REL = 1e-12


def clustered_lineitem(n, seed, by="shipdate", with_mvcc=True):
    """lineitem sorted by one column (real tables arrive roughly in date order): its zones become selective."""
    a = tpch.lineitem_arrays(n, seed, True)
    order = np.argsort(a[by], kind="stable")
    a = {k: v[order] for k, v in a.items()}
    t = HostTable(1)
    t.add(HostColumn(tpch.L_QUANTITY, tpch.DEC_15_2, decimal_from_i64(a["quantity"])))
    t.add(HostColumn(tpch.L_EXTENDEDPRICE, tpch.DEC_15_2, decimal_from_i64(a["extendedprice"])))
    t.add(HostColumn(tpch.L_DISCOUNT, tpch.DEC_15_2, decimal_from_i64(a["discount"])))
    t.add(HostColumn(tpch.L_SHIPDATE, DataType.Date32, a["shipdate"]))
    t.add(HostColumn(tpch.L_TAX, tpch.DEC_15_2, decimal_from_i64(a["tax"])))
    t.add(tpch._utf8_single_char(tpch.L_RETURNFLAG, a["returnflag"]))
    t.add(tpch._utf8_single_char(tpch.L_LINESTATUS, a["linestatus"]))
    snap = None
    if with_mvcc:
        c, d, snap = tpch.mvcc_arrays(n, seed)
        t.add_mvcc(c, d)
    return t, snap


def run(gpu_ctx, dt, expr, specs, keys=(), hint=0, snap=None, lo=0, hi=None, prune=2, jit=2, part=1, cap=1 << 16):
    from llkv_b200 import gpu
    gpu_ctx.set_pruning(prune)
    gpu_ctx.set_jit(jit)
    gpu_ctx.set_partitioning(part)
    prog = gpu.Program(gpu_ctx, expr) if expr is not None else None
    dt.set_snapshot(snap)
    agg = gpu.Aggregation(dt, specs, keys, cardinality_hint=hint)
    try:
        agg.run(prog, snap is not None, lo, dt.n_rows if hi is None else hi)
        return agg.finalize(cap), agg.run_info()
    finally:
        agg.destroy()
        if prog:
            prog.destroy()
        gpu_ctx.set_pruning(1)
        gpu_ctx.set_jit(1)
        gpu_ctx.set_partitioning(1)


def test_q6_on_a_date_clustered_table_reads_one_year_of_tiles(gpu_ctx):
    from llkv_b200 import gpu
    n = 300_000
    t, _ = clustered_lineitem(n, seed=6, with_mvcc=False)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        want = oracle.aggregate(t, tpch.q6_filter(), tpch.q6_aggregates())
        for jit in (0, 2):
            got, info = run(gpu_ctx, dt, tpch.q6_filter(), tpch.q6_aggregates(), jit=jit)
            util.assert_same_result(got, want, REL)
            tiles = (n + info.rows_per_tile - 1) // info.rows_per_tile
            assert info.used_fast_kernel == 1 and 0.75 * tiles < info.tiles_pruned < 0.9 * tiles  # one year of seven survives
        dense, info0 = run(gpu_ctx, dt, tpch.q6_filter(), tpch.q6_aggregates(), prune=0)
        assert info0.tiles_pruned == 0
        util.assert_same_result(dense, want, REL)
        for lo, hi in ((1, n - 1), (40_000, 41_000), (70_001, 250_003), (5, 5)):  # ragged ranges across pruned and kept tiles
            want_r = oracle.aggregate(t, tpch.q6_filter(), tpch.q6_aggregates(), row_begin=lo, row_end=hi)
            got_r, _ = run(gpu_ctx, dt, tpch.q6_filter(), tpch.q6_aggregates(), lo=lo, hi=hi)
            util.assert_same_result(got_r, want_r, REL)
    finally:
        dt.destroy()


def test_default_mode_prunes_from_the_second_scan_of_unchanged_columns(gpu_ctx):
    from llkv_b200 import gpu
    n = 200_000
    t, _ = clustered_lineitem(n, seed=8, with_mvcc=False)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    prog = gpu.Program(gpu_ctx, tpch.q6_filter())
    agg = gpu.Aggregation(dt, tpch.q6_aggregates())
    try:
        want = oracle.aggregate(t, tpch.q6_filter(), tpch.q6_aggregates())
        pruned = []
        for _ in range(4):
            agg.reset()
            agg.run(prog, False, 0, n)
            util.assert_same_result(agg.finalize(1), want, REL)
            pruned.append(agg.run_info().tiles_pruned)
        assert pruned[0] == 0 and pruned[1] > 0 and pruned[1] == pruned[2] == pruned[3]
    finally:
        agg.destroy()
        prog.destroy()
        dt.destroy()


def test_grouped_scan_with_mvcc_on_a_clustered_table(gpu_ctx):
    from llkv_b200 import gpu
    n = 150_000
    t, snap = clustered_lineitem(n, seed=3)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        early = Expr.And([pred(tpch.L_SHIPDATE, Operator.LessThanOrEquals(Literal.Date32(tpch.date32(1993, 6, 30))))])
        want = oracle.aggregate(t, early, tpch.q1_aggregates(), snap, tpch.Q1_GROUP_BY, group_capacity=16)
        got, info = run(gpu_ctx, dt, early, tpch.q1_aggregates(), tpch.Q1_GROUP_BY, hint=4, snap=snap, cap=16)
        assert info.tiles_pruned > 0
        util.assert_same_result(got, want, REL)
    finally:
        dt.destroy()


def test_decimal_zones_and_ranges_that_match_nothing(gpu_ctx):
    """Decimal128 columns have zone maps too (the reference keeps no chunk statistics for them, store/core.rs:1029-1032)."""
    from llkv_b200 import gpu
    n = 120_000
    t, _ = clustered_lineitem(n, seed=5, by="discount", with_mvcc=False)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        between = Expr.And([pred(tpch.L_DISCOUNT, Operator.Range(Bound.Included(Literal.Decimal128(5, 2)), Bound.Included(Literal.Decimal128(7, 2))))])
        want = oracle.aggregate(t, between, tpch.q6_aggregates())
        got, info = run(gpu_ctx, dt, between, tpch.q6_aggregates())
        assert info.tiles_pruned > 0
        util.assert_same_result(got, want, REL)
        nothing = Expr.And([pred(tpch.L_DISCOUNT, Operator.GreaterThan(Literal.Decimal128(50, 2)))])
        want = oracle.aggregate(t, nothing, tpch.q6_aggregates() + [AggregateSpec("c", AggregateKind.CountStar())])
        got, info = run(gpu_ctx, dt, nothing, tpch.q6_aggregates() + [AggregateSpec("c", AggregateKind.CountStar())])
        assert info.kernel_launches == 0  # every tile dropped out
        util.assert_same_result(got, want, REL)
    finally:
        dt.destroy()


def test_unclustered_columns_and_changed_columns(gpu_ctx):
    from llkv_b200 import gpu
    rng = np.random.default_rng(12)
    n = 100_000
    x = rng.integers(0, 1 << 40, n, dtype=np.int64)
    u = np.sort(rng.integers(0, 1 << 64, n, dtype=np.uint64))  # above 2^63 too: the leaf compares unsigned
    t = HostTable(1).add(HostColumn(1, DataType.Int64, x)).add(HostColumn(2, DataType.UInt64, u))
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    specs = [AggregateSpec("s", AggregateKind.Sum(1, DataType.Int64)), AggregateSpec("c", AggregateKind.CountStar())]
    try:
        f = tpch.between_filter(1, 1 << 38, 1 << 39)
        got, info = run(gpu_ctx, dt, f, specs)
        assert info.tiles_pruned == 0  # every zone of a uniformly random column spans the range
        util.assert_same_result(got, oracle.aggregate(t, f, specs), REL)
        a, b = int(u[n // 2]), int(u[n // 2 + 9_000])
        fu = Expr.And([pred(2, Operator.Range(Bound.Included(Literal.UInt64(a) if hasattr(Literal, "UInt64") else Literal.Int128(a)),
                                               Bound.Included(Literal.UInt64(b) if hasattr(Literal, "UInt64") else Literal.Int128(b))))])
        got, info = run(gpu_ctx, dt, fu, specs)
        util.assert_same_result(got, oracle.aggregate(t, fu, specs), REL)
        # more rows arrive: the zone map is rebuilt for the new content
        x2 = rng.integers(0, 1 << 40, 30_000, dtype=np.int64)
        u2 = np.sort(rng.integers(0, 1 << 64, 30_000, dtype=np.uint64))
        dt.columns[1].append(HostColumn(1, DataType.Int64, x2))
        dt.columns[2].append(HostColumn(2, DataType.UInt64, u2))
        dt.n_rows = n + 30_000
        dt.seal()
        t2 = HostTable(1).add(HostColumn(1, DataType.Int64, np.concatenate([x, x2]))).add(HostColumn(2, DataType.UInt64, np.concatenate([u, u2])))
        got, _ = run(gpu_ctx, dt, fu, specs)
        util.assert_same_result(got, oracle.aggregate(t2, fu, specs), REL)
    finally:
        dt.destroy()


@pytest.mark.parametrize("batch_rows", [0, 9_000])
def test_pruned_scan_feeding_the_partitioned_group_by(gpu_ctx, monkeypatch, batch_rows):
    """Also in several launches (the tuple buffers bound a launch; here the bound is lowered): each launch walks its own
    run of the tile list."""
    from llkv_b200 import gpu
    if batch_rows:
        monkeypatch.setenv("LLKV_GPU_PART_BATCH_ROWS", str(batch_rows))
    rng = np.random.default_rng(31)
    n = 200_000
    k = np.sort(rng.integers(0, 60_000, n, dtype=np.int64))
    v = rng.integers(-500, 500, n, dtype=np.int64)
    t = HostTable(1).add(HostColumn(tpch.K_FIELD, DataType.Int64, k)).add(HostColumn(tpch.V_FIELD, DataType.Int64, v))
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        f = tpch.between_filter(tpch.K_FIELD, 10_000, 24_999)
        want = oracle.aggregate(t, f, tpch.highcard_aggregates(), None, (tpch.K_FIELD,), group_capacity=1 << 16)
        got, info = run(gpu_ctx, dt, f, tpch.highcard_aggregates(), (tpch.K_FIELD,), hint=60_000, part=2)
        assert info.tiles_pruned > 0 and info.partitions >= 2
        assert info.kernel_launches >= (8 if batch_rows else 2)
        util.assert_same_result(got, want, REL)
    finally:
        dt.destroy()


def test_an_in_list_over_a_sorted_unsigned_column_prunes_nothing(gpu_ctx):
    """An IN-list leaf is a set, not a range: its first two entries must not be read as the bounds of a zone test (they were
    for UInt64 columns, whose leaves and zones share the unsigned order).  Entries far apart in a sorted column: every tile
    between them holds matches of the other entries."""
    from llkv_b200 import gpu
    n = 200_000
    t = HostTable(1)
    t.add(HostColumn(1, DataType.UInt64, np.arange(n, dtype=np.uint64)))
    t.add(HostColumn(2, DataType.Int64, np.arange(n, dtype=np.int64) % 1000))
    wanted = [5, 7, 60_000, 120_001, 199_999]
    f = pred(1, Operator.In(wanted))
    specs = [AggregateSpec("n", AggregateKind.CountStar()), AggregateSpec("s", AggregateKind.Sum(2, DataType.Int64))]
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        want = oracle.aggregate(t, f, specs)
        for jit in (0, 2):
            for _ in range(2):  # (mode 1 would only prune from the second unchanged scan on; mode 2 always)
                got, info = run(gpu_ctx, dt, f, specs, jit=jit)
                util.assert_same_result(got, want, REL)
                assert info.used_fast_kernel == 1 and info.tiles_pruned == 0
        # the same values as ranges do prune
        rng_f = Expr.And([pred(1, Operator.GreaterThanOrEquals(60_000)), pred(1, Operator.LessThan(60_100))])
        got, info = run(gpu_ctx, dt, rng_f, specs)
        util.assert_same_result(got, oracle.aggregate(t, rng_f, specs), REL)
        assert info.tiles_pruned > 0
    finally:
        dt.destroy()
