"""GPU: the lean kernel as a build specialised on the plan shape (jit.cpp) against the oracle and against its own
interpreted build.  Same bar as test_gpu_parity: bit-exact for integer / decimal / count / min / max, 1e-12 for f64."""
import numpy as np
import pytest

import util
from llkv_b200 import tpch
from llkv_b200.expr import AggregateKind, AggregateSpec, CompareOp, DataType, Expr, ScalarExpr
from llkv_b200.table import HostColumn, HostTable, LlkvError, Snapshot
from oracle import oracle
from test_gpu_parity import PREDICATES, REL, all_aggs, device_table, mixed_table

pytestmark = pytest.mark.gpu
G = util.golden()


@pytest.fixture()
def jit_always(gpu_ctx):
    gpu_ctx.set_jit(2)
    yield gpu_ctx
    gpu_ctx.set_jit(1)
    gpu_ctx.set_tuning()


def run(ctx, dt, expr, specs, snapshot=None, group_by=(), hint=0, row_begin=0, row_end=None, cap=None):
    from llkv_b200 import gpu
    prog = gpu.Program(ctx, expr) if expr is not None else None
    dt.set_snapshot(snapshot)
    agg = gpu.Aggregation(dt, specs, group_by, cardinality_hint=hint)
    try:
        agg.run(prog, snapshot is not None, row_begin, row_end)
        got = agg.finalize(cap)
        return got, agg.run_info()
    finally:
        agg.destroy()
        if prog:
            prog.destroy()


def test_specialised_kernel_runs_q6_and_q1(jit_always):
    ctx = jit_always
    t, snap = tpch.lineitem_table(150_000, seed=6, with_q1=True, with_mvcc=True)
    dt = device_table(ctx, t)
    try:
        got, info = run(ctx, dt, tpch.q6_filter(), tpch.q6_aggregates())
        assert info.used_fast_kernel == 1 and info.used_jit_kernel == 1
        util.assert_same_result(got, oracle.aggregate(t, tpch.q6_filter(), tpch.q6_aggregates()), REL)
        got, info = run(ctx, dt, tpch.q1_filter(), tpch.q1_aggregates(), snap, tpch.Q1_GROUP_BY, hint=6, cap=16)
        assert info.used_jit_kernel == 1
        util.assert_same_result(got, oracle.aggregate(t, tpch.q1_filter(), tpch.q1_aggregates(), snap, group_by=tpch.Q1_GROUP_BY, group_capacity=16), REL)
        # ragged row ranges reuse the same specialised kernel: row range, literals and snapshot are run-time parameters
        for lo, hi in [(0, 1), (1, 1), (1023, 1025), (77_777, 150_000), (149_999, 150_000)]:
            got, info = run(ctx, dt, tpch.q6_filter(), tpch.q6_aggregates(), row_begin=lo, row_end=hi)
            assert info.used_jit_kernel == (1 if hi > lo else 0)  # an empty range launches nothing
            util.assert_same_result(got, oracle.aggregate(t, tpch.q6_filter(), tpch.q6_aggregates(), row_begin=lo, row_end=hi), REL)
    finally:
        dt.destroy()


def test_auto_mode_specialises_from_the_second_run(gpu_ctx):
    from llkv_b200 import gpu
    gpu_ctx.set_jit(1)
    t, _ = tpch.lineitem_table(40_000, seed=9, with_q1=False)
    dt = device_table(gpu_ctx, t)
    prog = gpu.Program(gpu_ctx, tpch.q6_filter())
    agg = gpu.Aggregation(dt, tpch.q6_aggregates() + [AggregateSpec("mx", AggregateKind.Max(tpch.L_QUANTITY, tpch.DEC_15_2))])
    try:
        want = oracle.aggregate(t, tpch.q6_filter(), tpch.q6_aggregates() + [AggregateSpec("mx", AggregateKind.Max(tpch.L_QUANTITY, tpch.DEC_15_2))])
        seen = []
        for _ in range(3):
            agg.reset()
            agg.run(prog, False)
            util.assert_same_result(agg.finalize(1), want, REL)
            seen.append(agg.run_info().used_jit_kernel)
        assert seen == [0, 1, 1]
    finally:
        agg.destroy()
        prog.destroy()
        dt.destroy()


@pytest.mark.parametrize("n", [1, 777, 20011])
def test_specialised_ungrouped_aggregates_match_oracle(jit_always, n):
    ctx = jit_always
    t = mixed_table(n, seed=5 + n)
    dt = device_table(ctx, t)
    try:
        for e in (None, PREDICATES[0], PREDICATES[5], Expr.Literal(False)):
            got, info = run(ctx, dt, e, all_aggs())
            util.assert_same_result(got, oracle.aggregate(t, e, all_aggs()), REL)
    finally:
        dt.destroy()


@pytest.mark.parametrize("nulls", [False, True], ids=["dense", "nulls"])
def test_specialised_group_by_matches_oracle(jit_always, nulls):
    """nulls: keys (1, 2) and arguments are nullable; a NULL key value is its own group (a null bit in the packed key)."""
    ctx = jit_always
    t = mixed_table(9000, seed=21, nulls=nulls, long_strings=False)
    d = DataType.Decimal128(15, 2)
    specs = [
        AggregateSpec("n", AggregateKind.CountStar()),
        AggregateSpec("s1", AggregateKind.Sum(1, DataType.Int64)),
        AggregateSpec("s5", AggregateKind.Sum(5, d)),
        AggregateSpec("a5", AggregateKind.Avg(5, d)),
        AggregateSpec("sx", AggregateKind.Sum(ScalarExpr.Column(5) * (1 - ScalarExpr.Column(5)), DataType.Decimal128(38, 4))),
        AggregateSpec("mn3", AggregateKind.Min(3, DataType.Float64)),
        AggregateSpec("s3", AggregateKind.Sum(3, DataType.Float64)),
        AggregateSpec("mx1", AggregateKind.Max(1, DataType.Int64)),
    ]
    dt = device_table(ctx, t)
    try:
        jitted = 0
        no_decimals = [sp for sp in specs if sp.alias not in ("s5", "a5", "sx")]
        for keys, hint, sel in [((10,), 6, specs), ((2,), 100, specs), ((9, 10), 0, specs), ((8,), 600, specs), ((1,), 2000, specs),
                                ((1,), 2000, no_decimals)]:
            for e in (None, PREDICATES[0]):
                try:
                    want = oracle.aggregate(t, e, sel, group_by=keys, group_capacity=1 << 14)
                except LlkvError as want_err:
                    # (nulls: a group whose Decimal128 argument is NULL on every row fails the reference's downcast; so do we)
                    with pytest.raises(LlkvError) as got_err:
                        run(ctx, dt, e, sel, group_by=keys, hint=hint, cap=1 << 14)
                    assert got_err.value.code == want_err.code
                    continue
                got, info = run(ctx, dt, e, sel, group_by=keys, hint=hint, cap=1 << 14)
                jitted += info.used_jit_kernel
                util.assert_same_result(got, want, REL)
        assert jitted >= 8
    finally:
        dt.destroy()


@pytest.mark.parametrize("mode", [0, 2], ids=["interpreted", "specialised"])
def test_mvcc_truth_table_through_the_lean_kernel(gpu_ctx, mode):
    """RowVersion::is_visible_for vectors (llkv-transaction/src/mvcc.rs:528-555 + one vector per rule) as COUNT(*) under
    a snapshot: all vectors in one table would share a snapshot, so each runs alone; every run has the same plan shape."""
    gpu_ctx.set_jit(mode)
    try:
        for v in G["mvcc"]["vectors"] + G["mvcc"]["rule_vectors"]:
            t = HostTable(1).add(HostColumn(1, DataType.Int64, np.array([5], dtype=np.int64)))
            t.add_mvcc(np.array([v["created_by"]], np.uint64), np.array([v["deleted_by"]], np.uint64))
            snap = Snapshot(v["txn_id"], v["snapshot_id"], tuple(v["noncommitted"]))
            dt = device_table(gpu_ctx, t)
            try:
                got, info = run(gpu_ctx, dt, None, [AggregateSpec("n", AggregateKind.CountStar())], snap)
                assert info.used_fast_kernel == 1 and info.used_jit_kernel == (1 if mode else 0)
            finally:
                dt.destroy()
            assert got[0][1][0].value == (1 if v["visible"] else 0), v["note"]
    finally:
        gpu_ctx.set_jit(1)


@pytest.mark.parametrize("tuning", [dict(block_threads=32, rows_per_thread=8, stages=4), dict(block_threads=256, rows_per_thread=1, stages=2, ctas_per_sm=2),
                                    dict(block_threads=64, rows_per_thread=4, ctas_per_sm=4), dict(block_threads=96, rows_per_thread=8)],
                         ids=["32x8", "256x1", "64x4", "96x8"])
def test_specialised_geometries_agree(jit_always, tuning):
    ctx = jit_always
    t, snap = tpch.lineitem_table(60_000, seed=3, with_q1=True, with_mvcc=True)
    ctx.set_tuning(**tuning)
    dt = device_table(ctx, t)
    try:
        got, info = run(ctx, dt, tpch.q6_filter(), tpch.q6_aggregates())
        assert info.used_jit_kernel == 1
        util.assert_same_result(got, oracle.aggregate(t, tpch.q6_filter(), tpch.q6_aggregates()), REL)
        got, info = run(ctx, dt, tpch.q1_filter(), tpch.q1_aggregates(), snap, tpch.Q1_GROUP_BY, hint=6, cap=16)
        assert info.used_jit_kernel == 1
        util.assert_same_result(got, oracle.aggregate(t, tpch.q1_filter(), tpch.q1_aggregates(), snap, group_by=tpch.Q1_GROUP_BY, group_capacity=16), REL)
    finally:
        dt.destroy()


def test_count_star_without_any_column(gpu_ctx):
    """COUNT(*) with no filter and no snapshot reads no column at all: the plan must not take the tile pipeline."""
    t, _ = tpch.lineitem_table(10_000, seed=2, with_q1=False)
    dt = device_table(gpu_ctx, t)
    try:
        for mode in (0, 2):
            gpu_ctx.set_jit(mode)
            got, info = run(gpu_ctx, dt, None, [AggregateSpec("n", AggregateKind.CountStar())])
            assert got[0][1][0].value == 10_000
            got, info = run(gpu_ctx, dt, None, [AggregateSpec("n", AggregateKind.CountStar())], row_begin=17, row_end=4242)
            assert got[0][1][0].value == 4242 - 17
    finally:
        gpu_ctx.set_jit(1)
        dt.destroy()


@pytest.mark.parametrize("mode", [0, 2], ids=["interpreted", "specialised"])
@pytest.mark.parametrize("hint", [1, 3, 4])
def test_four_slot_associative_group_table(gpu_ctx, mode, hint):
    """Cardinality hints <= 4 take the four-slot fully associative CTA table (Q1 has four groups); a hint below the real
    cardinality and a fifth, sixth ... group (column 10 has six distinct strings) overflow to the global table."""
    gpu_ctx.set_jit(mode)
    try:
        t, snap = tpch.lineitem_table(90_000, seed=11, with_q1=True, with_mvcc=True)
        dt = device_table(gpu_ctx, t)
        try:
            got, info = run(gpu_ctx, dt, tpch.q1_filter(), tpch.q1_aggregates(), snap, tpch.Q1_GROUP_BY, hint=hint, cap=16)
            assert info.used_fast_kernel == 1 and info.fast_groups == 4
            util.assert_same_result(got, oracle.aggregate(t, tpch.q1_filter(), tpch.q1_aggregates(), snap, group_by=tpch.Q1_GROUP_BY, group_capacity=16), REL)
        finally:
            dt.destroy()
        m = mixed_table(7000, seed=4, long_strings=False)
        dm = device_table(gpu_ctx, m)
        try:
            specs = [AggregateSpec("n", AggregateKind.CountStar()), AggregateSpec("s1", AggregateKind.Sum(1, DataType.Int64)),
                     AggregateSpec("mx5", AggregateKind.Max(5, DataType.Decimal128(15, 2)))]
            for keys in [(10,), (9,), (9, 10)]:
                got, info = run(gpu_ctx, dm, PREDICATES[0], specs, group_by=keys, hint=hint, cap=64)
                util.assert_same_result(got, oracle.aggregate(m, PREDICATES[0], specs, group_by=keys, group_capacity=64), REL)
        finally:
            dm.destroy()
    finally:
        gpu_ctx.set_jit(1)


@pytest.mark.parametrize("mode", [0, 2], ids=["interpreted", "specialised"])
def test_nullable_keys_in_the_four_slot_table(gpu_ctx, mode):
    """Three key values + NULL under a hint of 4: the four-slot table compares the packed keys, null bit included, in 32 bits;
    with two nullable keys (nine + groups) the surplus groups take the global-table copy of the program."""
    rng = np.random.default_rng(17)
    n = 50_000
    t = HostTable(1)
    for fid, hi in ((1, 3), (2, 2)):
        c = HostColumn(fid, DataType.Int32, rng.integers(0, hi, n).astype(np.int32))
        c.validity = np.packbits(rng.random(n) > 0.15, bitorder="little")
        t.add(c)
    v = HostColumn(3, DataType.Int64, rng.integers(-1000, 1000, n, dtype=np.int64))
    v.validity = np.packbits(rng.random(n) > 0.1, bitorder="little")
    t.add(v)
    specs = [AggregateSpec("n", AggregateKind.CountStar()), AggregateSpec("c", AggregateKind.Count(3)),
             AggregateSpec("s", AggregateKind.Sum(3, DataType.Int64)), AggregateSpec("mn", AggregateKind.Min(3, DataType.Int64))]
    gpu_ctx.set_jit(mode)
    dt = device_table(gpu_ctx, t)
    try:
        for keys in [(1,), (1, 2)]:
            for e in (None, Expr.Compare(ScalarExpr.Column(3) + ScalarExpr.Column(1), CompareOp.Gt, ScalarExpr.Column(2) * ScalarExpr.Column(2))):
                got, info = run(gpu_ctx, dt, e, specs, group_by=keys, hint=4, cap=64)
                assert info.used_fast_kernel == 1 and info.fast_groups == 4 and info.used_jit_kernel == (1 if mode else 0)
                util.assert_same_result(got, oracle.aggregate(t, e, specs, group_by=keys, group_capacity=64), REL)
    finally:
        gpu_ctx.set_jit(1)
        dt.destroy()


def test_a_low_cardinality_hint_is_corrected_after_the_first_finalize(gpu_ctx):
    """Seven groups under a hint of 3: the first run has four CTA-local slots, so three groups send their rows to the global
    table one atomic at a time.  finalize sees seven groups and the next run of the same aggregate gets eight slots."""
    from llkv_b200 import gpu
    rng = np.random.default_rng(5)
    n = 80_000
    t = HostTable(1).add(HostColumn(tpch.K_FIELD, DataType.Int64, rng.integers(0, 7, n, dtype=np.int64))).add(
        HostColumn(tpch.V_FIELD, DataType.Int64, rng.integers(-1000, 1000, n, dtype=np.int64)))
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    want = oracle.aggregate(t, None, tpch.highcard_aggregates(), None, (tpch.K_FIELD,), group_capacity=16)
    agg = gpu.Aggregation(dt, tpch.highcard_aggregates(), (tpch.K_FIELD,), cardinality_hint=3)
    try:
        slots = []
        for _ in range(3):
            agg.reset()
            agg.run(None, False)
            util.assert_same_result(agg.finalize(16), want, 1e-12)
            slots.append(agg.run_info().fast_groups)
        assert slots == [4, 8, 8]
    finally:
        agg.destroy()
        dt.destroy()
