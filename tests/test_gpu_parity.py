"""GPU: the CUDA path through the C ABI (libllkv_gpu.so) against the oracle on the same seeded inputs, and against the
reference's known answers.  Bit-exact for integer / decimal / count / min / max / bitmaps; 1e-12 relative for f64 sums."""
import math

import numpy as np
import pytest

import util
from llkv_b200 import ffi, tpch
from llkv_b200.expr import (AggregateKind, AggregateSpec, BinaryOp, Bound, CompareOp, DataType, Expr, Literal, Operator,
                            ScalarExpr, pred)
from llkv_b200.table import HostColumn, HostTable, LlkvError, Snapshot, decimal_array, decimal_from_i64, pack_validity
from oracle import oracle

pytestmark = pytest.mark.gpu
G = util.golden()
REL = 1e-12  # north_star: f64 SUM/AVG within 1e-12 relative (reassociation changes rounding)


def device_table(gpu_ctx, table, **kw):
    from llkv_b200 import gpu
    return gpu.DeviceTable.from_host(gpu_ctx, table, **kw)


def check_filter(gpu_ctx, table, expr, snapshot=None, row_begin=0, row_end=None, dt=None):
    own = dt is None
    dt = dt or device_table(gpu_ctx, table)
    try:
        want_w, want_c = oracle.filter_bitmap(table, expr, snapshot, row_begin, row_end)
        got_w, got_c = dt.filter_bitmap(expr, snapshot, row_begin, row_end)
        assert got_c == want_c
        assert np.array_equal(got_w, want_w)
    finally:
        if own:
            dt.destroy()


def check_agg(gpu_ctx, table, expr, specs, snapshot=None, group_by=(), dt=None, ordered=True, **kw):
    own = dt is None
    dt = dt or device_table(gpu_ctx, table)
    try:
        okw = {k: v for k, v in kw.items() if k != "cardinality_hint"}
        try:
            want = oracle.aggregate(table, expr, specs, snapshot, group_by, **okw)
        except LlkvError as want_err:  # the reference fails this query: the GPU path must fail the same way
            with pytest.raises(LlkvError) as got_err:
                dt.aggregate(expr, specs, snapshot, group_by, **kw)
            assert got_err.value.code == want_err.code, (got_err.value, want_err)
            return None
        got = dt.aggregate(expr, specs, snapshot, group_by, **kw)
        util.assert_same_result(got, want, REL, ordered)
        return got
    finally:
        if own:
            dt.destroy()


# ---------------------------------------------------------------- reference known answers through the GPU
@pytest.mark.parametrize("case", G["filter_cases"], ids=lambda c: c["name"])
def test_filter_known_answers(gpu_ctx, case):
    t = util.table_from_json(G["table_t4"])
    dt = device_table(gpu_ctx, t)
    try:
        words, count = dt.filter_bitmap(util.expr_from_json(case["filter"]))
    finally:
        dt.destroy()
    pos = util.selected_positions(words, t.n_rows)
    assert count == len(pos)
    assert util.host_values(t.columns[case["select"]], pos) == case["expect"]


@pytest.mark.parametrize("mode", [0, 2], ids=["interpreted", "specialised"])
def test_filter_known_answers_through_the_lean_kernel(gpu_ctx, mode):
    """The same reference-held filters (llkv-table/src/table.rs tests: ranges, IN lists, floats, AND / OR / NOT, the
    two-column `a + c > 220` comparison) as the selection of an aggregate: every one of them runs on the lean kernel, and
    COUNT(*) is the number of rows the reference's test expects."""
    from llkv_b200 import gpu
    t = util.table_from_json(G["table_t4"])
    dt = device_table(gpu_ctx, t)
    gpu_ctx.set_jit(mode)
    try:
        for case in G["filter_cases"]:
            prog = gpu.Program(gpu_ctx, util.expr_from_json(case["filter"]))
            agg = gpu.Aggregation(dt, [AggregateSpec("n", AggregateKind.CountStar()), AggregateSpec("c", AggregateKind.Count(case["select"]))])
            try:
                agg.run(prog, False)
                (_, vals), = agg.finalize(1)
                assert agg.run_info().used_fast_kernel == 1, case["name"]
                assert vals[0].value == len(case["expect"]), case["name"]
                assert vals[1].value == sum(v is not None for v in case["expect"]), case["name"]
            finally:
                agg.destroy()
                prog.destroy()
    finally:
        gpu_ctx.set_jit(1)
        dt.destroy()


def test_count_with_a_complex_expression_filter(gpu_ctx):
    """llkv-slt-tester/tests/slt/duckdb/constraints/primarykey/test_pk_append_many_duplicates.slt:27-39 through the GPU:
    `SELECT COUNT(*) FROM integers WHERE i+i = ${val}*2` is 1 for every val of 0..99 (an Expr::Compare of two scalar
    expressions; the lean kernel's FO_CMP, specialised once: the literal is a run-time parameter)."""
    from llkv_b200 import gpu
    t = HostTable(1).add(HostColumn(1, DataType.Int64, np.arange(100, dtype=np.int64)))
    dt = device_table(gpu_ctx, t)
    i = ScalarExpr.Column(1)
    try:
        agg = gpu.Aggregation(dt, [AggregateSpec("n", AggregateKind.CountStar())])
        try:
            for val in list(range(100)) + [100, -1]:
                prog = gpu.Program(gpu_ctx, Expr.Compare(i + i, CompareOp.Eq, ScalarExpr.Literal(val) * 2))
                try:
                    agg.reset()
                    agg.run(prog, False)
                    (_, vals), = agg.finalize(1)
                    assert vals[0].value == (1 if 0 <= val < 100 else 0), val
                    assert agg.run_info().used_fast_kernel == 1
                finally:
                    prog.destroy()
        finally:
            agg.destroy()
    finally:
        dt.destroy()


@pytest.mark.parametrize("case", G["aggregate_cases"], ids=lambda c: c["name"])
def test_aggregate_known_answers(gpu_ctx, case):
    col = util.column_from_json(1, case["column"])
    t = HostTable(1).add(col)
    arg = util.sexpr_from_json(case["expr"]) if "expr" in case else 1
    kind = {"avg": AggregateKind.Avg, "sum": AggregateKind.Sum}[case["agg"]](arg, col.dtype)
    dt = device_table(gpu_ctx, t)
    try:
        v = dt.aggregate(None, [AggregateSpec("a", kind)])[0][1][0]
    finally:
        dt.destroy()
    assert v.value == case["expect"]


def test_computed_projection_known_answer(gpu_ctx):
    case = G["computed_cases"][0]
    t = util.table_from_json(G["table_t4"])
    e = util.sexpr_from_json(case["expr"])
    dt = device_table(gpu_ctx, t)
    try:
        for i, want in enumerate(case["expect"]):
            v = dt.aggregate(None, [AggregateSpec("v", AggregateKind.Min(e, DataType.Float64))], row_begin=i, row_end=i + 1)[0][1][0]
            assert v.type == ffi.PT_FLOAT64 and v.value == want
    finally:
        dt.destroy()


def test_mvcc_truth_table(gpu_ctx):
    vs = G["mvcc"]["vectors"] + G["mvcc"]["rule_vectors"]
    for v in vs:  # one row per vector, its own snapshot
        t = HostTable(1).add(HostColumn(1, DataType.Int64, np.array([5], dtype=np.int64)))
        t.add_mvcc(np.array([v["created_by"]], np.uint64), np.array([v["deleted_by"]], np.uint64))
        snap = Snapshot(v["txn_id"], v["snapshot_id"], tuple(v["noncommitted"]))
        dt = device_table(gpu_ctx, t)
        try:
            _, count = dt.filter_bitmap(None, snap)
        finally:
            dt.destroy()
        assert bool(count) == v["visible"], v["note"]


# ---------------------------------------------------------------- seeded parity against the oracle
def mixed_table(n, seed, nulls=False, long_strings=True):
    rng = np.random.default_rng(seed)
    t = HostTable(1)
    cols = [
        HostColumn(1, DataType.Int64, rng.integers(-1000, 1000, n, dtype=np.int64)),
        HostColumn(2, DataType.Int32, rng.integers(-50, 50, n, dtype=np.int64).astype(np.int32)),
        HostColumn(3, DataType.Float64, np.round(rng.normal(0, 100, n), 3)),
        HostColumn(4, DataType.UInt64, rng.integers(0, 500, n, dtype=np.int64).astype(np.uint64)),
        HostColumn(5, DataType.Decimal128(15, 2), decimal_from_i64(rng.integers(-10**7, 10**7, n, dtype=np.int64))),
        HostColumn(6, DataType.Date32, rng.integers(8000, 11000, n, dtype=np.int64).astype(np.int32)),
        HostColumn(7, DataType.Float32, rng.integers(-20, 20, n, dtype=np.int64).astype(np.float32) / 4),
        HostColumn(8, DataType.Int16, rng.integers(-300, 300, n, dtype=np.int64).astype(np.int16)),
        HostColumn(9, DataType.Boolean, rng.integers(0, 2, n, dtype=np.int64).astype(np.uint8)),
    ]
    strs = ["A", "N", "R", "", "xy", "abcdefg" if long_strings else "abcde"]
    cols.append(HostColumn.utf8(10, [strs[i] for i in rng.integers(0, len(strs), n)]))
    if nulls:
        for c in cols[:6]:
            c.validity = pack_validity(rng.random(n) > 0.2)
    for c in cols:
        t.add(c)
    return t


PREDICATES = [
    pred(1, Operator.GreaterThan(0)),
    pred(1, Operator.Range(Bound.Included(-100), Bound.Excluded(250))),
    pred(2, Operator.In([1, 2, 3, -7])),
    pred(3, Operator.LessThanOrEquals(12.5)),
    pred(4, Operator.Equals(42)),
    pred(5, Operator.Range(Bound.Included(Literal.Decimal128(-12345, 2)), Bound.Included(Literal.Decimal128(9999999, 3)))),
    pred(5, Operator.In([Literal.Decimal128(100, 2), Literal.Int128(7)])),
    pred(6, Operator.Range(Bound.Included(Literal.Date32(9000)), Bound.Excluded(Literal.Date32(9500)))),
    pred(7, Operator.In([0.25, -1.0, 3.5])),
    pred(8, Operator.LessThan(-10)),
    pred(9, Operator.Equals(True)),
    pred(10, Operator.Equals("N")),
    pred(10, Operator.Range(Bound.Included("A"), Bound.Included("a"))),
    pred(1, Operator.IsNull),
    pred(1, Operator.IsNotNull),
    pred(2, Operator.Range(Bound.Unbounded, Bound.Unbounded)),
    Expr.And([pred(1, Operator.GreaterThanOrEquals(-500)), pred(1, Operator.LessThanOrEquals(500))]),  # FusedAnd
    Expr.And([pred(1, Operator.GreaterThan(-500)), pred(2, Operator.LessThan(25)), pred(3, Operator.GreaterThan(-50.0))]),
    Expr.Or([pred(1, Operator.LessThan(-900)), pred(4, Operator.GreaterThan(450)), pred(10, Operator.Equals("xy"))]),
    Expr.Not(pred(2, Operator.Equals(3))),
    Expr.Not(Expr.And([pred(1, Operator.GreaterThan(0)), pred(2, Operator.LessThan(10))])),
    Expr.Not(Expr.Or([pred(1, Operator.IsNull), pred(3, Operator.GreaterThan(0.0))])),
    Expr.Compare(ScalarExpr.Column(1) + ScalarExpr.Column(2), CompareOp.Gt, ScalarExpr.Literal(100)),
    Expr.Compare(ScalarExpr.Column(1) * 2, CompareOp.LtEq, ScalarExpr.Column(4)),
    Expr.Compare(ScalarExpr.Column(3), CompareOp.Lt, ScalarExpr.Column(1)),
    Expr.Compare(ScalarExpr.Column(5), CompareOp.GtEq, ScalarExpr.Literal(Literal.Decimal128(0, 2))),
    Expr.Compare(ScalarExpr.Column(5) * ScalarExpr.Column(5), CompareOp.Gt, ScalarExpr.Literal(Literal.Decimal128(10**9, 4))),
    Expr.Compare(ScalarExpr.Column(1) / ScalarExpr.Column(2), CompareOp.Eq, ScalarExpr.Literal(3)),
    Expr.Compare(ScalarExpr.Column(6), CompareOp.Lt, ScalarExpr.Literal(Literal.Date32(9100))),
    Expr.InList(ScalarExpr.Column(1), [1, 2, ScalarExpr.Column(2)]),
    Expr.InList(ScalarExpr.Column(2), [5, Literal.Null()], negated=True),
    Expr.InList(ScalarExpr.Column(3), [], negated=True),
    Expr.IsNull(ScalarExpr.Column(1) + ScalarExpr.Column(2)),
    Expr.IsNull(ScalarExpr.Column(3), negated=True),
    Expr.And([Expr.Literal(True), pred(1, Operator.GreaterThan(0))]),
    Expr.Or([Expr.Literal(False), Expr.Not(Expr.Compare(ScalarExpr.Column(1), CompareOp.NotEq, ScalarExpr.Column(2)))]),
]


@pytest.mark.parametrize("nulls", [False, True], ids=["dense", "nulls"])
def test_predicates_match_oracle(gpu_ctx, nulls):
    t = mixed_table(5000 + 37, seed=11, nulls=nulls)
    dt = device_table(gpu_ctx, t)
    try:
        for i, e in enumerate(PREDICATES):
            try:
                check_filter(gpu_ctx, t, e, dt=dt)
            except AssertionError as err:
                raise AssertionError(f"predicate #{i}: {err}") from err
        # ragged sub-ranges
        for rb, re in [(0, 0), (1, 2), (63, 65), (1000, 4999), (5036, 5037)]:
            check_filter(gpu_ctx, t, PREDICATES[1], row_begin=rb, row_end=re, dt=dt)
    finally:
        dt.destroy()


def all_aggs():
    d = DataType.Decimal128(15, 2)
    return [
        AggregateSpec("n", AggregateKind.CountStar()),
        AggregateSpec("c1", AggregateKind.Count(1)),
        AggregateSpec("cn", AggregateKind.CountNulls(1)),
        AggregateSpec("s1", AggregateKind.Sum(1, DataType.Int64)),
        AggregateSpec("a1", AggregateKind.Avg(1, DataType.Int64)),
        AggregateSpec("t1", AggregateKind.Total(1, DataType.Int64)),
        AggregateSpec("mn1", AggregateKind.Min(1, DataType.Int64)),
        AggregateSpec("mx1", AggregateKind.Max(1, DataType.Int64)),
        AggregateSpec("s3", AggregateKind.Sum(3, DataType.Float64)),
        AggregateSpec("a3", AggregateKind.Avg(3, DataType.Float64)),
        AggregateSpec("mn3", AggregateKind.Min(3, DataType.Float64)),
        AggregateSpec("mx3", AggregateKind.Max(3, DataType.Float64)),
        AggregateSpec("s5", AggregateKind.Sum(5, d)),
        AggregateSpec("a5", AggregateKind.Avg(5, d)),
        AggregateSpec("mn5", AggregateKind.Min(5, d)),
        AggregateSpec("mx5", AggregateKind.Max(5, d)),
        AggregateSpec("sf5", AggregateKind.Sum(5, DataType.Float64)),
        AggregateSpec("sx", AggregateKind.Sum(ScalarExpr.Column(1) * ScalarExpr.Column(1) + 3, DataType.Int64)),
        AggregateSpec("sd", AggregateKind.Sum(ScalarExpr.Column(5) * ScalarExpr.Column(5), d)),
        AggregateSpec("sfx", AggregateKind.Sum(ScalarExpr.Column(3) * ScalarExpr.Column(1), DataType.Float64)),
    ]


@pytest.mark.parametrize("nulls", [False, True], ids=["dense", "nulls"])
@pytest.mark.parametrize("n", [0, 1, 777, 20011])
def test_ungrouped_aggregates_match_oracle(gpu_ctx, n, nulls):
    t = mixed_table(n, seed=5 + n, nulls=nulls)
    dt = device_table(gpu_ctx, t)
    try:
        for e in (None, PREDICATES[0], PREDICATES[17], Expr.Literal(False)):
            check_agg(gpu_ctx, t, e, all_aggs(), dt=dt)
    finally:
        dt.destroy()


@pytest.mark.parametrize("nulls", [False, True], ids=["dense", "nulls"])
def test_group_by_matches_oracle(gpu_ctx, nulls):
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
    dt = device_table(gpu_ctx, t)
    try:
        for keys in [(10,), (2,), (9, 10), (8,), (2, 9, 10), (1,), (6,)]:
            for e in (None, PREDICATES[0]):
                try:
                    check_agg(gpu_ctx, t, e, specs, group_by=keys, dt=dt, group_capacity=1 << 14)
                except AssertionError as err:
                    raise AssertionError(f"keys {keys}: {err}") from err
    finally:
        dt.destroy()


def test_group_key_wider_than_64_bits(gpu_ctx):
    """11 + 9 + 59 bits: no packed key; the groups are keyed by a verified hash and the key values read back from the columns
    (tests/test_wide_keys.py has the full set)."""
    t = mixed_table(5000, seed=1)
    check_agg(gpu_ctx, t, None, [AggregateSpec("n", AggregateKind.CountStar()), AggregateSpec("s", AggregateKind.Sum(1, DataType.Int64))],
              group_by=(1, 4, 10), group_capacity=1 << 14)


def test_high_cardinality_group_by(gpu_ctx):
    t = tpch.highcard_table(300_000, 50_000, seed=4)
    got = check_agg(gpu_ctx, t, None, tpch.highcard_aggregates(), group_by=(tpch.K_FIELD,), group_capacity=1 << 17,
                    cardinality_hint=50_000)
    assert len(got) > 49_000
    # without a hint the group table has to grow on the fly
    check_agg(gpu_ctx, t, None, tpch.highcard_aggregates(), group_by=(tpch.K_FIELD,), group_capacity=1 << 17)


def test_mvcc_filter_matches_oracle(gpu_ctx):
    n = 30_000
    t, _ = tpch.int64_table(n, seed=3, with_mvcc=False)
    c, d, snap = tpch.mvcc_arrays(n, seed=3)
    t.add_mvcc(c, d)
    dt = device_table(gpu_ctx, t)
    try:
        check_filter(gpu_ctx, t, None, snap, dt=dt)
        check_filter(gpu_ctx, t, tpch.between_filter(tpch.X_FIELD, -10**8, 10**8), snap, dt=dt)
        check_agg(gpu_ctx, t, tpch.between_filter(tpch.X_FIELD, -10**8, 10**8), tpch.sum_int64(tpch.X_FIELD), snap, dt=dt)
        own = Snapshot(77, 100, (77, 101))  # the Active transaction reads its own writes
        check_filter(gpu_ctx, t, None, own, dt=dt)
    finally:
        dt.destroy()


def test_q6_and_q1_small(gpu_ctx):
    t, snap = tpch.lineitem_table(120_000, seed=6, with_q1=True, with_mvcc=True)
    dt = device_table(gpu_ctx, t)
    try:
        got = check_agg(gpu_ctx, t, tpch.q6_filter(), tpch.q6_aggregates(), dt=dt)
        assert got[0][1][0].value > 0
        check_agg(gpu_ctx, t, tpch.q6_filter(), tpch.q6_aggregates(), snap, dt=dt)
        got = check_agg(gpu_ctx, t, tpch.q1_filter(), tpch.q1_aggregates(), snap, group_by=tpch.Q1_GROUP_BY, dt=dt,
                        group_capacity=16, cardinality_hint=6)
        assert len(got) == 4
    finally:
        dt.destroy()


def test_wide_decimals_use_the_128_bit_path(gpu_ctx):
    rng = np.random.default_rng(9)
    big = [int(x) * 10**20 + int(y) for x, y in zip(rng.integers(-10**15, 10**15, 4000), rng.integers(0, 10**9, 4000))]
    t = HostTable(1).add(HostColumn(1, DataType.Decimal128(38, 4), decimal_array(big)))
    t.add(HostColumn(2, DataType.Decimal128(15, 2), decimal_from_i64(rng.integers(1, 10**12, 4000, dtype=np.int64))))
    d = DataType.Decimal128(38, 4)
    specs = [AggregateSpec("s", AggregateKind.Sum(1, d)), AggregateSpec("a", AggregateKind.Avg(1, d)),
             AggregateSpec("mn", AggregateKind.Min(1, d)), AggregateSpec("mx", AggregateKind.Max(1, d)),
             # fits i64 per value, but the products need 128 bits: the 64-bit pass must hand over
             AggregateSpec("p", AggregateKind.Sum(ScalarExpr.Column(2) * ScalarExpr.Column(2), DataType.Decimal128(31, 4)))]
    check_agg(gpu_ctx, t, None, specs)
    check_agg(gpu_ctx, t, pred(1, Operator.GreaterThan(Literal.Decimal128(10**30, 4))), specs)
    check_filter(gpu_ctx, t, Expr.Compare(ScalarExpr.Column(2) * ScalarExpr.Column(2), CompareOp.Gt,
                                          ScalarExpr.Literal(Literal.Decimal128(10**22, 4))))


def test_errors_match_oracle(gpu_ctx):
    t = HostTable(1).add(HostColumn(1, DataType.Int64, np.array([2**62, 2**62, 5], dtype=np.int64)))
    t.add(HostColumn(2, DataType.Int64, np.array([3, 0, 1], dtype=np.int64)))
    cases = [
        (None, [AggregateSpec("s", AggregateKind.Sum(1, DataType.Int64))]),                                  # integer overflow
        (None, [AggregateSpec("s", AggregateKind.Sum(ScalarExpr.Column(1) * ScalarExpr.Column(1), DataType.Int64))]),  # arrow overflow
        (Expr.Compare(ScalarExpr.Column(1) * 4, CompareOp.Gt, ScalarExpr.Literal(0)), [AggregateSpec("n", AggregateKind.CountStar())]),
        (pred(1, Operator.Equals(2.5)), [AggregateSpec("n", AggregateKind.CountStar())]),                    # literal type mismatch
        (pred(99, Operator.Equals(1)), [AggregateSpec("n", AggregateKind.CountStar())]),                     # unknown field
        (None, [AggregateSpec("s", AggregateKind.Sum(1, DataType.Decimal128(10, 2)))]),                      # Expected Decimal128 array
        (None, [AggregateSpec("s", AggregateKind.Sum(1, DataType.Date32))]),                                 # unsupported type
    ]
    dt = device_table(gpu_ctx, t)
    try:
        for e, specs in cases:
            with pytest.raises(LlkvError) as want:
                oracle.aggregate(t, e, specs)
            with pytest.raises(LlkvError) as got:
                dt.aggregate(e, specs)
            assert got.value.code == want.value.code, (got.value, want.value)
        # errors the reference raises from update() do not fire when no row reaches the accumulator
        none = Expr.Literal(False)
        spec = [AggregateSpec("s", AggregateKind.Sum(1, DataType.Decimal128(10, 2)))]
        util.assert_same_result(dt.aggregate(none, spec), oracle.aggregate(t, none, spec))
        # x / 0 is NULL, not an error (llkv-compute/src/kernels.rs:121-135)
        check_agg(gpu_ctx, t, None, [AggregateSpec("s", AggregateKind.Sum(ScalarExpr.Column(2) / ScalarExpr.Column(2), DataType.Int64))], dt=dt)
    finally:
        dt.destroy()


def test_float_min_max_nan_rules(gpu_ctx):
    nan = float("nan")
    for vals in ([nan, 1.0, 2.0], [1.0, nan, 0.5], [nan, nan], [3.0, 2.0, nan]):
        t = HostTable(1).add(HostColumn(1, DataType.Float64, np.array(vals)))
        specs = [AggregateSpec("mn", AggregateKind.Min(1, DataType.Float64)), AggregateSpec("mx", AggregateKind.Max(1, DataType.Float64))]
        check_agg(gpu_ctx, t, None, specs)
    check_filter(gpu_ctx, HostTable(1).add(HostColumn(1, DataType.Float64, np.array([nan, 1.0, -0.0, 0.0]))),
                 Expr.Or([pred(1, Operator.GreaterThanOrEquals(0.0)), Expr.Compare(ScalarExpr.Column(1), CompareOp.Gt, ScalarExpr.Literal(0.5))]))


def test_chunked_upload_paths_agree(gpu_ctx):
    """Blob ("ARR0") appends, odd chunk sizes with validity bitmaps, and re-upload after clear()."""
    t = mixed_table(10_007, seed=2, nulls=True)
    want = oracle.aggregate(t, PREDICATES[17], all_aggs())
    for kw in ({"chunk_rows": 4096}, {"chunk_rows": 1001}, {"chunk_rows": 10_007}):
        dt = device_table(gpu_ctx, t, **kw)
        try:
            util.assert_same_result(dt.aggregate(PREDICATES[17], all_aggs()), want, REL)
        finally:
            dt.destroy()
    dense = mixed_table(10_007, seed=2, nulls=False)
    dt = device_table(gpu_ctx, dense, chunk_rows=4096, as_blob=True)
    try:
        util.assert_same_result(dt.aggregate(PREDICATES[17], all_aggs()), oracle.aggregate(dense, PREDICATES[17], all_aggs()), REL)
        col = dt.columns[1]
        col.clear()
        col.append(dense.columns[1], 500)
        col.seal()
        util.assert_same_result(dt.aggregate(PREDICATES[17], all_aggs()), oracle.aggregate(dense, PREDICATES[17], all_aggs()), REL)
    finally:
        dt.destroy()


def test_accumulators_fold_across_runs(gpu_ctx):
    """update() per row range == one update over the whole table (AggregateAccumulator::update per batch)."""
    from llkv_b200 import gpu
    t = mixed_table(50_000, seed=8)
    dt = device_table(gpu_ctx, t)
    agg = gpu.Aggregation(dt, all_aggs())
    prog = gpu.Program(gpu_ctx, PREDICATES[1])
    try:
        for rb, re in [(0, 10_000), (10_000, 10_001), (10_001, 37_777), (37_777, 50_000)]:
            agg.run(prog, False, rb, re)
        util.assert_same_result(agg.finalize(), oracle.aggregate(t, PREDICATES[1], all_aggs()), REL)
        agg.reset()
        agg.run(prog, False, 0, 50_000)
        util.assert_same_result(agg.finalize(), oracle.aggregate(t, PREDICATES[1], all_aggs()), REL)
    finally:
        prog.destroy()
        agg.destroy()
        dt.destroy()


@pytest.mark.parametrize("tuning", [dict(block_threads=128, rows_per_thread=1, stages=1), dict(block_threads=256, rows_per_thread=4, stages=2),
                                    dict(block_threads=512, rows_per_thread=2, stages=4, ctas_per_sm=1), dict(force_wide=1),
                                    dict(force_wide=2), dict(force_wide=2, block_threads=128, rows_per_thread=1, stages=1),
                                    dict(block_threads=64, rows_per_thread=1, stages=2), dict(block_threads=128, rows_per_thread=8, ctas_per_sm=3)],
                         ids=["lean-128x1", "lean-256x4", "lean-512x2x4", "wide", "general64", "general64-direct", "lean-64x1", "lean-128x8"])
def test_every_kernel_geometry_agrees(gpu_ctx, tuning):
    t, snap = tpch.lineitem_table(60_000, seed=3, with_q1=True, with_mvcc=True)
    gpu_ctx.set_tuning(**tuning)
    try:
        dt = device_table(gpu_ctx, t)
        try:
            check_agg(gpu_ctx, t, tpch.q6_filter(), tpch.q6_aggregates(), snap, dt=dt)
            check_agg(gpu_ctx, t, tpch.q1_filter(), tpch.q1_aggregates(), snap, group_by=tpch.Q1_GROUP_BY, dt=dt, group_capacity=16,
                      cardinality_hint=6)
            check_filter(gpu_ctx, t, tpch.q6_filter(), snap, dt=dt)
        finally:
            dt.destroy()
    finally:
        gpu_ctx.set_tuning()


def test_lean_kernel_runs_the_benchmark_plans(gpu_ctx):
    """Q6, Q1, the filtered SUM and (since round 2) OR / NOT trees and IN lists lower to the lean kernel; suffix / substring
    patterns stay on the general interpreter."""
    from llkv_b200 import gpu
    t, snap = tpch.lineitem_table(50_000, seed=3, with_q1=True, with_mvcc=True)
    dt = device_table(gpu_ctx, t)
    try:
        dt.set_snapshot(snap)
        for expr, specs, keys, mvcc, want_fast in [
            (tpch.q6_filter(), tpch.q6_aggregates(), (), False, 1),
            (tpch.q1_filter(), tpch.q1_aggregates(), tpch.Q1_GROUP_BY, True, 1),
            (Expr.Or([tpch.q1_filter(), tpch.q6_filter()]), tpch.q6_aggregates(), (), False, 1),
            (Expr.Or([tpch.q1_filter(), pred(tpch.L_QUANTITY, Operator.In([100, 200]))]), tpch.q6_aggregates(), (), False, 1),
                (Expr.And([tpch.q1_filter(), pred(tpch.L_RETURNFLAG, Operator.EndsWith("N"))]), tpch.q6_aggregates(), (), False, 0),
        ]:
            prog = gpu.Program(gpu_ctx, expr)
            agg = gpu.Aggregation(dt, specs, keys, cardinality_hint=6 if keys else 0)
            try:
                agg.run(prog, mvcc)
                agg.finalize(16)
                assert agg.run_info().used_fast_kernel == want_fast
            finally:
                agg.destroy()
                prog.destroy()
    finally:
        dt.destroy()
    t2, snap2 = tpch.int64_table(10_000, seed=1)
    dt = device_table(gpu_ctx, t2)
    try:
        dt.set_snapshot(snap2)
        prog = gpu.Program(gpu_ctx, tpch.between_filter(tpch.X_FIELD, -5, 10**8))
        agg = gpu.Aggregation(dt, tpch.sum_int64(tpch.X_FIELD))
        agg.run(prog, True)
        got = agg.finalize(1)
        assert agg.run_info().used_fast_kernel == 1
        util.assert_same_result(got, oracle.aggregate(t2, tpch.between_filter(tpch.X_FIELD, -5, 10**8), tpch.sum_int64(tpch.X_FIELD), snap2))
        agg.destroy()
        prog.destroy()
    finally:
        dt.destroy()


def test_full_size_properties_q6(gpu_ctx):
    """At a size the oracle does not run in seconds: size-independent properties.  SUM over the whole table equals the
    sum of SUMs over disjoint row ranges, COUNT(*) of a filter equals the popcount of its bitmap, and the integer
    columns check against numpy's exact sums."""
    from llkv_b200 import gpu
    n = 6_001_215  # SF1
    a = tpch.lineitem_arrays(n, seed=6, with_q1=False)
    t, _ = tpch.lineitem_table(n, seed=6, with_q1=False)
    dt = device_table(gpu_ctx, t)
    try:
        specs = tpch.q6_aggregates() + [AggregateSpec("n", AggregateKind.CountStar()),
                                        AggregateSpec("sp", AggregateKind.Sum(tpch.L_EXTENDEDPRICE, tpch.DEC_15_2))]
        whole = dt.aggregate(tpch.q6_filter(), specs)[0][1]
        parts = [dt.aggregate(tpch.q6_filter(), specs, row_begin=lo, row_end=hi)[0][1]
                 for lo, hi in [(0, 1_000_003), (1_000_003, 4_500_000), (4_500_000, n)]]
        for i in range(3):
            assert whole[i].value == sum(p[i].value for p in parts)
        _, count = dt.filter_bitmap(tpch.q6_filter())
        assert count == whole[1].value
        m = ((a["shipdate"] >= tpch.date32(1994, 1, 1)) & (a["shipdate"] < tpch.date32(1995, 1, 1)) & (a["discount"] >= 5)
             & (a["discount"] <= 7) & (a["quantity"] < 2400))
        assert whole[1].value == int(m.sum())
        assert whole[2].value == int(a["extendedprice"][m].sum())
        prod = a["extendedprice"][m] * a["discount"][m]  # scale 4 -> scale 2, half away from zero (values are >= 0)
        assert whole[0].value == int(((prod + 50) // 100).sum())
    finally:
        dt.destroy()


def test_full_size_properties_sf10_q6_q1(gpu_ctx):
    """BASELINE.json configs 2 and 3 at their full size (SF10, 59 986 052 rows), where the oracle does not finish in
    seconds: size-independent properties.  Exact integer identities against numpy restatements of the two queries over
    the generator's arrays (per group for Q1, with the snapshot rule evaluated for the generator's transaction ids),
    additivity over disjoint row ranges, and COUNT(*) against the popcount of the selection bitmap, which a different
    kernel (the general interpreter) produces.  Runs on the specialised lean kernel from the second run on."""
    from llkv_b200 import gpu
    n = tpch.lineitem_rows(10.0)
    assert n == 59_986_052
    a = tpch.lineitem_arrays(n, seed=6, with_q1=True)
    t, _ = tpch.lineitem_table(n, seed=6, with_q1=True)
    created, deleted, snap = tpch.mvcc_arrays(n, seed=6)
    t.add_mvcc(created, deleted)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t, chunk_rows=1 << 20)
    del t
    try:
        # ---- Q6
        specs6 = tpch.q6_aggregates() + [AggregateSpec("n", AggregateKind.CountStar()),
                                         AggregateSpec("sp", AggregateKind.Sum(tpch.L_EXTENDEDPRICE, tpch.DEC_15_2))]
        cuts = [0, 7_000_001, 33_333_333, n]
        whole = dt.aggregate(tpch.q6_filter(), specs6)[0][1]
        parts = [dt.aggregate(tpch.q6_filter(), specs6, row_begin=lo, row_end=hi)[0][1] for lo, hi in zip(cuts, cuts[1:])]
        for i in range(3):
            assert whole[i].value == sum(p[i].value for p in parts)
        m = ((a["shipdate"] >= tpch.date32(1994, 1, 1)) & (a["shipdate"] < tpch.date32(1995, 1, 1)) & (a["discount"] >= 5)
             & (a["discount"] <= 7) & (a["quantity"] < 2400))
        _, count = dt.filter_bitmap(tpch.q6_filter())
        assert count == whole[1].value == int(m.sum())
        assert whole[2].value == int(a["extendedprice"][m].sum())
        prod = a["extendedprice"][m] * a["discount"][m]  # scale 4 -> scale 2, half away from zero (values are >= 0)
        assert whole[0].value == int(((prod + 50) // 100).sum())
        # ---- Q1 under the snapshot
        vis = (created == 1) & ((deleted == np.uint64(tpch.TXN_ID_NONE)) | (deleted == np.uint64(77)))
        sel = vis & (a["shipdate"] <= tpch.date32(1998, 9, 2))
        _, count1 = dt.filter_bitmap(tpch.q1_filter(), snap)
        assert count1 == int(sel.sum())
        got = dt.aggregate(tpch.q1_filter(), tpch.q1_aggregates(), snap, group_by=tpch.Q1_GROUP_BY, cardinality_hint=6, group_capacity=16)
        lo_half = dt.aggregate(tpch.q1_filter(), tpch.q1_aggregates(), snap, group_by=tpch.Q1_GROUP_BY, cardinality_hint=6, group_capacity=16,
                               row_end=29_000_003)
        hi_half = dt.aggregate(tpch.q1_filter(), tpch.q1_aggregates(), snap, group_by=tpch.Q1_GROUP_BY, cardinality_hint=6, group_capacity=16,
                               row_begin=29_000_003)
        assert len(got) == 4 and sum(v[7].value for _, v in got) == count1
        halves = {}
        for rows in (lo_half, hi_half):
            for k, v in rows:
                acc = halves.setdefault(k, [0, 0, 0, 0, 0])
                for i, j in enumerate((0, 1, 2, 3, 7)):
                    acc[i] += v[j].value
        disc_price = a["extendedprice"] * (100 - a["discount"])      # scale 4, exact
        for key, v in got:
            g = sel & (a["returnflag"] == ord(key[0])) & (a["linestatus"] == ord(key[1]))
            cnt = int(g.sum())
            assert v[7].value == cnt and cnt > 0, key
            assert v[0].value == int(a["quantity"][g].sum()), key
            assert v[1].value == int(a["extendedprice"][g].sum()), key
            assert v[2].value == int(disc_price[g].sum()), key
            charge = disc_price[g].astype(object) * (100 + a["tax"][g]).astype(object) if cnt < 200_000 else None
            if charge is not None:
                assert v[3].value == int(charge.sum()), key
            else:  # exact in int64: disc_price < 2^31, (100 + tax) < 2^7, 3e7 rows
                assert v[3].value == int((disc_price[g] * (100 + a["tax"][g])).sum()), key
            # AVG = sum / count at the input scale, rounded half away from zero (llkv-aggregate/src/lib.rs:1720-1761)
            q, r = divmod(int(a["quantity"][g].sum()), cnt)
            assert v[4].value == q + (1 if 2 * r >= cnt else 0), key
            assert [v[j].value for j in (0, 1, 2, 3, 7)] == halves[key], key
    finally:
        dt.destroy()


@pytest.mark.parametrize("as_blob", [False, True], ids=["values", "blobs"])
def test_resident_image_round_trips_chunks(gpu_ctx, as_blob):
    """Chunks appended as Arrow value buffers or as the pager's serialized blobs ("ARR0" header,
    llkv-column-map/src/serialization.rs:41-53) land as one dense resident image; reading row ranges back gives the Arrow
    values again, including Decimal128 columns the device keeps as i64 and ranges that straddle chunk boundaries."""
    t = mixed_table(10_007, seed=2)
    dt = device_table(gpu_ctx, t, chunk_rows=1000, as_blob=as_blob)
    try:
        for fid, col in t.columns.items():
            if col.dtype.type == ffi.PT_UTF8:
                continue
            dc = dt.columns[fid]
            assert dc.rows() == 10_007
            want = col.values.reshape(-1, 2) if col.dtype.type == ffi.PT_DECIMAL128 else col.values
            for lo, n in [(0, 10_007), (999, 2), (1000, 1000), (5555, 4452), (10_006, 1), (17, 0)]:
                got = dc.read(lo, n)
                assert np.array_equal(got.view(np.uint8), np.ascontiguousarray(want[lo:lo + n]).view(np.uint8)), (fid, lo, n)
        with pytest.raises(LlkvError):
            dt.columns[1].read(10_000, 8)
    finally:
        dt.destroy()


def test_column_loaded_through_its_descriptor_chain(gpu_ctx):
    """The reference's scan source end to end: ColumnDescriptor -> descriptor pages -> ChunkMetadata -> one batched get of the
    chunk blobs (llkv-column-map/src/store/scan/unsorted.rs:202-241) -> llkv_gpu_column_append_blob, then the fused scan."""
    from llkv_b200 import gpu, metadata
    from oracle import metadata as om
    rng = np.random.default_rng(21)
    n = 20_000
    x = np.sort(rng.integers(-1_000_000, 1_000_000, n, dtype=np.int64))  # clustered: chunk statistics are selective
    col = HostColumn(7, DataType.Int64, x)
    pager, metas, pk = {}, [], 100
    chunk_rows = 4096
    for lo in range(0, n, chunk_rows):
        part = HostColumn(7, DataType.Int64, x[lo:lo + chunk_rows])
        blob = part.serialize()
        st = om.chunk_stats(om.INT64, part.values)
        pager[pk] = blob
        metas.append((pk, 0, part.n_rows, len(blob), st[0], st[1], st[2], st[3]))
        pk += 1
    pages = om.descriptor_pages(  # short pages: the walk follows the chain
        metas, list(range(50, 50 + (len(metas) + 2) // 3)), per_page=3)
    pager.update(dict(pages))
    pager[49] = om.descriptor_bytes(gpu.logical_field_id(1, 7), pages[0][0], pages[-1][0], n, len(metas), data_type_code=ffi.PT_INT64)
    gets = []

    def batch_get(pks):
        gets.append(list(pks))
        return [pager[k] for k in pks]

    desc, chunks, skipped = metadata.walk_descriptor(batch_get, 49)
    assert skipped == 0 and desc.total_row_count == n and sum(c.row_count for c in chunks) == n
    blobs = batch_get([c.chunk_pk for c in chunks])  # one batched get for every chunk of the scan
    dc = gpu.DeviceColumn(gpu_ctx, int(desc.field_id), col)
    base = 0
    for c, blob in zip(chunks, blobs):
        assert c.serialized_bytes == len(blob)
        gpu._check(gpu_ctx.lib.llkv_gpu_column_append_blob(dc.handle, c.chunk_pk, blob, len(blob), None, base))
        base += c.row_count
    dt = gpu.DeviceTable(gpu_ctx, 1)
    dt.columns[7] = dc
    dt.n_rows = n
    dt.seal()
    try:
        a, b = int(x[n // 3]), int(x[n // 2])
        t = HostTable(1).add(col)
        f = tpch.between_filter(7, a, b)
        want = oracle.aggregate(t, f, tpch.sum_int64(7))
        got = dt.aggregate(f, tpch.sum_int64(7))
        util.assert_same_result(got, want, REL)
        # what a pruned fetch would have skipped (the resident column stays dense: pruning here only reports)
        _, survivors, skipped = metadata.walk_descriptor(batch_get, 49, om.INT64, (0, a), (0, b))
        assert skipped >= 2 and all(om.chunk_matches(om.INT64, (0, a), (0, b), c.min_val_u64, c.max_val_u64) for c in survivors)
        kept = np.concatenate([x[(c.chunk_pk - 100) * chunk_rows:(c.chunk_pk - 100 + 1) * chunk_rows] for c in survivors])
        assert int(kept[(kept >= a) & (kept <= b)].sum()) == got[0][1][0].value  # no matching row lives in a skipped chunk
    finally:
        dt.destroy()


def test_chunk_blobs_appended_from_registered_pager_memory(gpu_ctx):
    """The pager's mmap-backed blobs, page-locked in place (llkv_gpu_host_register): appends DMA out of the mapping and the
    resident column reads back identical."""
    import mmap
    from llkv_b200 import gpu
    rng = np.random.default_rng(17)
    x = rng.integers(-10**12, 10**12, 200_000, dtype=np.int64)
    blobs = [HostColumn(3, DataType.Int64, x[lo:lo + 50_000]).serialize() for lo in range(0, x.size, 50_000)]
    store = mmap.mmap(-1, sum(len(b) for b in blobs) + 4096)  # an anonymous mapping standing in for the pager's file mapping
    offs, o = [], 0
    for b in blobs:
        store[o:o + len(b)] = b
        offs.append(o)
        o += len(b)
    view = np.frombuffer(store, dtype=np.uint8)
    ptr = gpu.host_register(view)
    dc = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(1, 3), HostColumn(3, DataType.Int64, x[:1]))
    try:
        base = 0
        for b, off in zip(blobs, offs):
            gpu._check(gpu_ctx.lib.llkv_gpu_column_append_blob(dc.handle, 1 + base, ptr + off, len(b), None, base))
            base += 50_000
        dc.seal()
        assert np.array_equal(dc.read(), x)
    finally:
        dc.destroy()
        gpu.host_unregister(ptr)
        del view
        store.close()


@pytest.mark.parametrize("step", G["count_star_transactions"]["steps"], ids=lambda s: s["note"][:40])
def test_count_star_under_transactions(gpu_ctx, step):
    """The reference's expected COUNT(*) per connection while another one holds uncommitted deletes / appends
    (llkv-slt-tester/tests/slt/duckdb/transactions/count_star_transactions.slt), interpreted and specialised."""
    t = util.count_star_scenario_table(step)
    snap = Snapshot(step["txn_id"], step["snapshot_id"], tuple(step["noncommitted"]))
    dt = device_table(gpu_ctx, t)
    try:
        for mode in (0, 2):
            gpu_ctx.set_jit(mode)
            got = dt.aggregate(None, [AggregateSpec("n", AggregateKind.CountStar()), AggregateSpec("s", AggregateKind.Sum(1, DataType.Int64))], snap)
            assert got[0][1][0].value == step["count"], step["note"]
        words, count = dt.filter_bitmap(None, snap)
        assert count == step["count"]
    finally:
        gpu_ctx.set_jit(1)
        dt.destroy()


@pytest.mark.parametrize("case", G["nullable_aggregate_cases"], ids=lambda c: c["name"])
def test_aggregates_over_nullable_integers(gpu_ctx, case):
    """COUNT(i), SUM(i), MIN(i), MAX(i), COUNT(*) over an INTEGER column with NULLs: the reference's SLT outputs."""
    t = HostTable(1).add(util.nullable_int_column(1, case))
    names, specs = util.nullable_aggregate_specs(case)
    dt = device_table(gpu_ctx, t)
    try:
        got = dt.aggregate(None, specs)[0][1]
        assert {n: v.value for n, v in zip(names, got)} == case["expect"]
    finally:
        dt.destroy()
