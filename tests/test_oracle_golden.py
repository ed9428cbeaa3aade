"""CPU: the oracle (oracle/llkv_oracle.c) against every known answer the reference's own tests hold for this path
(tests/golden/reference_known_answers.json, each case citing reference file:line)."""
import numpy as np
import pytest

import util
from llkv_b200 import ffi
from llkv_b200.expr import AggregateKind, AggregateSpec, DataType, Expr, ScalarExpr
from llkv_b200.table import HostColumn, HostTable, Snapshot
from oracle import oracle

G = util.golden()


@pytest.mark.parametrize("case", G["filter_cases"], ids=lambda c: c["name"])
def test_filter_known_answers(case):
    t = util.table_from_json(G["table_t4"])
    words, count = oracle.filter_bitmap(t, util.expr_from_json(case["filter"]))
    pos = util.selected_positions(words, t.n_rows)
    assert count == len(pos)
    got = util.host_values(t.columns[case["select"]], pos)
    assert got == case["expect"]
    if "expect_sum" in case:
        assert sum(got) == case["expect_sum"]
    if "expect_min" in case:
        assert min(got) == case["expect_min"] and max(got) == case["expect_max"]


@pytest.mark.parametrize("case", G["computed_cases"], ids=lambda c: c["name"])
def test_computed_projection_known_answers(case):
    # the computed column is observed through MIN over single-row ranges: same evaluator, one value at a time
    t = util.table_from_json(G["table_t4"])
    e = util.sexpr_from_json(case["expr"])
    for i, want in enumerate(case["expect"]):
        rows = oracle.aggregate(t, None, [AggregateSpec("v", AggregateKind.Min(e, DataType.Float64))], row_begin=i, row_end=i + 1)
        v = rows[0][1][0]
        assert v.type == ffi.PT_FLOAT64 and v.value == want


def test_mvcc_truth_table():
    for v in G["mvcc"]["vectors"] + G["mvcc"]["rule_vectors"]:
        got = oracle.mvcc_visible(v["created_by"], v["deleted_by"], v["txn_id"], v["snapshot_id"], v["noncommitted"])
        assert got == v["visible"], v["note"]


@pytest.mark.parametrize("case", G["aggregate_cases"], ids=lambda c: c["name"])
def test_aggregate_known_answers(case):
    col = util.column_from_json(1, case["column"])
    t = HostTable(1).add(col)
    arg = util.sexpr_from_json(case["expr"]) if "expr" in case else 1
    kind = {"avg": AggregateKind.Avg, "sum": AggregateKind.Sum}[case["agg"]](arg, col.dtype)
    v = oracle.aggregate(t, None, [AggregateSpec("a", kind)])[0][1][0]
    assert v.value == case["expect"]
    if "expect_scale" in case:
        assert v.scale == case["expect_scale"]


def test_chunk_blob_header():
    # "ARR0" | layout | PrimType | precision | scale | len u64 | values bytes u32 | 0 (serialization.rs:41-53,264-307)
    col = HostColumn(1, DataType.Int64, np.arange(5, dtype=np.int64))
    blob = oracle.serialize_primitive(col)
    assert blob == col.serialize()
    assert blob[:4] == b"ARR0" and blob[4] == 0 and blob[5] == ffi.PT_INT64
    assert int.from_bytes(blob[8:16], "little") == 5 and int.from_bytes(blob[16:20], "little") == 40
    assert len(blob) == 24 + 40


def test_sum_int64_overflow_is_an_error():
    t = HostTable(1).add(HostColumn(1, DataType.Int64, np.array([2**62, 2**62], dtype=np.int64)))
    with pytest.raises(oracle.LlkvError) as e:
        oracle.aggregate(t, None, [AggregateSpec("s", AggregateKind.Sum(1, DataType.Int64))])
    assert e.value.code == ffi.ERR_INVALID_ARGUMENT and "integer overflow" in e.value.message


def test_sum_decimal_of_no_rows_is_zero_not_null():
    t = HostTable(1).add(util.column_from_json(1, {"type": "Decimal128", "precision": 10, "scale": 2, "values": []}))
    v = oracle.aggregate(t, None, [AggregateSpec("s", AggregateKind.Sum(1, DataType.Decimal128(10, 2)))])[0][1][0]
    assert v.value == 0 and v.scale == 2


@pytest.mark.parametrize("step", G["count_star_transactions"]["steps"], ids=lambda s: s["note"][:40])
def test_count_star_under_transactions(step):
    """llkv-slt-tester/tests/slt/duckdb/transactions/count_star_transactions.slt: the counts every connection sees while
    another one holds uncommitted deletes and appends."""
    t = util.count_star_scenario_table(step)
    snap = Snapshot(step["txn_id"], step["snapshot_id"], tuple(step["noncommitted"]))
    got = oracle.aggregate(t, None, [AggregateSpec("n", AggregateKind.CountStar())], snap)
    assert got[0][1][0].value == step["count"], step["note"]


@pytest.mark.parametrize("case", G["nullable_aggregate_cases"], ids=lambda c: c["name"])
def test_aggregates_over_nullable_integers(case):
    t = HostTable(1).add(util.nullable_int_column(1, case))
    names, specs = util.nullable_aggregate_specs(case)
    got = oracle.aggregate(t, None, specs)[0][1]
    assert {n: v.value for n, v in zip(names, got)} == case["expect"]


def test_count_with_a_complex_expression_filter():
    """llkv-slt-tester/tests/slt/duckdb/constraints/primarykey/test_pk_append_many_duplicates.slt:27-39: integers(i) holds
    0..99; for every val `SELECT COUNT(*) FROM integers WHERE i+i = ${val}*2` ("a complex expression to prevent index
    lookup") and `... WHERE i = ${val}` return 1.  The first is an Expr::Compare of two scalar expressions."""
    from llkv_b200.expr import CompareOp, Expr, Operator, ScalarExpr, pred
    t = HostTable(1).add(HostColumn(1, DataType.Int64, np.arange(100, dtype=np.int64)))
    count = [AggregateSpec("n", AggregateKind.CountStar())]
    i = ScalarExpr.Column(1)
    for val in range(100):
        complex_filter = Expr.Compare(i + i, CompareOp.Eq, ScalarExpr.Literal(val) * 2)
        assert oracle.aggregate(t, complex_filter, count)[0][1][0].value == 1
        assert oracle.aggregate(t, pred(1, Operator.Equals(val)), count)[0][1][0].value == 1
    # (outside the loaded range both are 0)
    assert oracle.aggregate(t, Expr.Compare(i + i, CompareOp.Eq, ScalarExpr.Literal(100) * 2), count)[0][1][0].value == 0
