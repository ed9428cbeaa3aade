"""Shared test helpers: golden-file decoding, table builders, result comparison."""
import json
import math
import os

import numpy as np

from llkv_b200 import ffi
from llkv_b200.expr import (AggregateKind, AggregateSpec, BinaryOp, Bound, CompareOp, DataType, Expr, Literal, Operator,
                            ScalarExpr, pred)
from llkv_b200.table import HostColumn, HostTable, decimal_array

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_known_answers.json")
_TYPES = {"UInt64": DataType.UInt64, "Int32": DataType.Int32, "Int64": DataType.Int64, "Float64": DataType.Float64,
          "Float32": DataType.Float32, "Date32": DataType.Date32}
_BIN = {"add": BinaryOp.Add, "sub": BinaryOp.Subtract, "mul": BinaryOp.Multiply, "div": BinaryOp.Divide, "mod": BinaryOp.Modulo}
_CMP = {"eq": CompareOp.Eq, "ne": CompareOp.NotEq, "lt": CompareOp.Lt, "le": CompareOp.LtEq, "gt": CompareOp.Gt, "ge": CompareOp.GtEq}


def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def table_from_json(t) -> HostTable:
    ht = HostTable(t.get("table_id", 1))
    for c in t["columns"]:
        ht.add(column_from_json(c["field"], c))
    return ht


def column_from_json(field, c) -> HostColumn:
    if c["type"] == "Decimal128":
        return HostColumn(field, DataType.Decimal128(c["precision"], c["scale"]), decimal_array(c["values"]))
    return HostColumn(field, _TYPES[c["type"]], np.asarray(c["values"]))


def sexpr_from_json(j) -> ScalarExpr:
    if "col" in j:
        return ScalarExpr.Column(j["col"])
    if "lit" in j:
        return ScalarExpr.Literal(j["lit"])
    if "bin" in j:
        l, op, r = j["bin"]
        return ScalarExpr.Binary(sexpr_from_json(l), _BIN[op], sexpr_from_json(r))
    raise ValueError(j)


def _bound(b):
    return Bound.Included(b[1]) if b[0] == "included" else Bound.Excluded(b[1])


def expr_from_json(j) -> Expr:
    if "pred" in j:
        p = j["pred"]
        op = p["op"]
        if op == "range":
            o = Operator.Range(_bound(p["lower"]) if "lower" in p else Bound.Unbounded,
                               _bound(p["upper"]) if "upper" in p else Bound.Unbounded)
        elif op == "in":
            o = Operator.In(p["values"])
        else:
            o = {"eq": Operator.Equals, "gt": Operator.GreaterThan, "ge": Operator.GreaterThanOrEquals,
                 "lt": Operator.LessThan, "le": Operator.LessThanOrEquals}[op](p["value"])
        return pred(p["field"], o)
    if "and" in j:
        return Expr.And([expr_from_json(c) for c in j["and"]])
    if "or" in j:
        return Expr.Or([expr_from_json(c) for c in j["or"]])
    if "not" in j:
        return Expr.Not(expr_from_json(j["not"]))
    if "compare" in j:
        c = j["compare"]
        return Expr.Compare(sexpr_from_json(c["left"]), _CMP[c["op"]], sexpr_from_json(c["right"]))
    raise ValueError(j)


def selected_positions(words: np.ndarray, n_rows: int):
    bits = np.unpackbits(words.view(np.uint8), bitorder="little")[:n_rows]
    return np.nonzero(bits)[0]


def host_values(col: HostColumn, positions):
    return [col.values[i].item() for i in positions]


def values_equal(a, b, rel=1e-12):
    """AggregateValue comparison: bit-exact for integers / decimals / NULLs, `rel` relative for f64."""
    if a.type != b.type:
        return False
    if (a.value is None) != (b.value is None):
        return False
    if a.value is None:
        return True
    if a.type == ffi.PT_FLOAT64:
        x, y = float(a.value), float(b.value)
        if math.isnan(x) or math.isnan(y):
            return math.isnan(x) and math.isnan(y)
        if x == y:
            return True
        return abs(x - y) <= rel * max(abs(x), abs(y))
    if a.type == ffi.PT_DECIMAL128:
        return a.value == b.value and a.scale == b.scale and a.precision == b.precision
    return a.value == b.value


def assert_same_result(got, want, rel=1e-12, ordered=True):
    """Both are [(key_tuple, [AggregateValue...])...]."""
    assert len(got) == len(want), f"group count {len(got)} != {len(want)}"
    if not ordered:
        got = sorted(got, key=lambda r: repr(r[0]))
        want = sorted(want, key=lambda r: repr(r[0]))
    for (gk, gv), (wk, wv) in zip(got, want):
        assert gk == wk, f"group key {gk} != {wk}"
        assert len(gv) == len(wv)
        for i, (a, b) in enumerate(zip(gv, wv)):
            assert values_equal(a, b, rel), f"group {gk} aggregate {i}: {a} != {b}"


def count_star_scenario_table(step) -> HostTable:
    """Table state of one step of the COUNT(*)-under-transactions scenario (tests/golden: count_star_transactions): an id
    column plus created_by / deleted_by built from the step's (count, created_by, deleted_by) runs; the first two runs of
    an "interleaved" step alternate row by row (DELETE ... WHERE id % 2 = 0)."""
    runs = step["rows"]
    created, deleted = [], []
    if step.get("interleaved"):
        (n0, c0, d0), (n1, c1, d1) = runs[0], runs[1]
        assert n0 == n1
        c = np.empty(2 * n0, dtype=np.uint64)
        d = np.empty(2 * n0, dtype=np.uint64)
        c[0::2], c[1::2], d[0::2], d[1::2] = c0, c1, d0, d1
        created.append(c)
        deleted.append(d)
        runs = runs[2:]
    for n, c, d in runs:
        created.append(np.full(n, c, dtype=np.uint64))
        deleted.append(np.full(n, d, dtype=np.uint64))
    created, deleted = np.concatenate(created), np.concatenate(deleted)
    t = HostTable(1).add(HostColumn(1, DataType.Int64, np.arange(created.size, dtype=np.int64)))
    t.add_mvcc(created, deleted)
    return t


def nullable_int_column(field, case) -> HostColumn:
    """Int64 column of a golden "nullable_aggregate_cases" entry: `base` (None = NULL) repeated `repeat` times."""
    from llkv_b200.table import pack_validity
    base = case["base"] * case["repeat"]
    valid = np.array([v is not None for v in base], dtype=bool)
    values = np.array([0 if v is None else v for v in base], dtype=np.int64)
    return HostColumn(field, DataType.Int64, values, pack_validity(valid))


def nullable_aggregate_specs(case):
    kinds = {"count_col": AggregateKind.Count(1), "sum": AggregateKind.Sum(1, DataType.Int64), "min": AggregateKind.Min(1, DataType.Int64),
             "max": AggregateKind.Max(1, DataType.Int64), "count_star": AggregateKind.CountStar()}
    names = list(case["expect"])
    return names, [AggregateSpec(n, kinds[n]) for n in names]
