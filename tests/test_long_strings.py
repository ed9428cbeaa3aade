"""Utf8 values longer than the 7 bytes of a packed key ("What's missing" 6; GroupKeyValue::String,
llkv-executor/src/lib.rs:99-106,9362-9456; typed predicates on String, llkv-expr/src/typed_predicate.rs:170-209).
CPU half: the oracle compares whole strings byte-wise — checked against Python's str operators.  GPU half: the column is
dictionary-coded by the host side of append, codes are ranks after seal, every leaf is an integer leaf over ranks — checked
against the oracle for predicates, GROUP BY (first-appearance order, NULL group), HAVING / ORDER BY over the keys, appends
after seal and a column that turns long in a later chunk."""
import numpy as np
import pytest

import util
from llkv_b200 import ffi
from llkv_b200.expr import AggregateKind, AggregateSpec, Bound, CompareOp, DataType, Expr, Filter, Operator, ScalarExpr, pred
from llkv_b200.table import HostColumn, HostTable, LlkvError
from oracle import oracle

MODES = ["DELIVER IN PERSON", "COLLECT COD", "NONE", "TAKE BACK RETURN", "AIR", "REG AIR", "RAIL", "", "TRUCK", "SHIP", "MAIL", "FOB",
         "take back return", "DELIVER", "DELIVER IN PERSON!", "Zürich Hauptbahnhof"]
ASCII_MODES = [m for m in MODES if all(ord(ch) < 128 for ch in m)]
LITS = ["DELIVER IN PERSON", "DELIVER", "NONE", "", "AIR", "ZZZ", "A", "R", "TAKE BACK RETURN", "take", "DELIVER IN PERSO", "0", "~"]


def table(n=20_000, seed=1, nulls=False, words=MODES, tid=61):
    rng = np.random.default_rng(seed)
    vals = [words[i] for i in rng.integers(0, len(words), n)]
    if nulls:
        keep = rng.random(n) > 0.15
        vals = [v if k else None for v, k in zip(vals, keep)]
    t = HostTable(tid).add(HostColumn.utf8(1, vals)).add(HostColumn(2, DataType.Int64, rng.integers(-100, 100, n, dtype=np.int64)))
    t.add(HostColumn(3, DataType.Int32, rng.integers(0, 5, n, dtype=np.int64).astype(np.int32)))
    return t, vals


def leaves():
    out = []
    for a in LITS:
        out += [Operator.Equals(a), Operator.GreaterThan(a), Operator.GreaterThanOrEquals(a), Operator.LessThan(a), Operator.LessThanOrEquals(a),
                Operator.StartsWith(a), Operator.EndsWith(a), Operator.Contains(a)]
        for b in ("NONE", "TAKE BACK RETURN", "", "~"):
            out.append(Operator.Range(Bound.Included(a), Bound.Excluded(b)))
            out.append(Operator.Range(Bound.Excluded(a), Bound.Included(b)))
    out.append(Operator.In(["NONE", "DELIVER IN PERSON", "nope", "TRUCK", "NONE"]))
    out.append(Operator.In([]))
    return out


def py_eval(op, v):
    if v is None:
        return False
    b = v.encode()
    lits = [bytes(l.value if isinstance(l.value, bytes) else str(l.value).encode()) for l in op.literals]
    t = op.tag
    if t == ffi.OP_EQUALS:
        return b == lits[0]
    if t == ffi.OP_GT:
        return b > lits[0]
    if t == ffi.OP_GTE:
        return b >= lits[0]
    if t == ffi.OP_LT:
        return b < lits[0]
    if t == ffi.OP_LTE:
        return b <= lits[0]
    if t == ffi.OP_IN:
        return b in lits
    if t == ffi.OP_STARTS_WITH:
        return b.startswith(lits[0])
    if t == ffi.OP_ENDS_WITH:
        return b.endswith(lits[0])
    if t == ffi.OP_CONTAINS:
        return lits[0] in b
    k = 0
    ok = True
    if op.lower.kind != ffi.BOUND_UNBOUNDED:
        ok = ok and (b >= lits[k] if op.lower.kind == ffi.BOUND_INCLUDED else b > lits[k])
        k += 1
    if op.upper.kind != ffi.BOUND_UNBOUNDED:
        ok = ok and (b <= lits[k] if op.upper.kind == ffi.BOUND_INCLUDED else b < lits[k])
    return ok


@pytest.mark.parametrize("nulls", [False, True])
def test_oracle_compares_whole_strings(nulls):
    t, vals = table(n=4000, nulls=nulls)
    for op in leaves():
        words, count = oracle.filter_bitmap(t, Expr.Pred(Filter(1, op)))
        got = set(util.selected_positions(words, t.n_rows))
        want = {i for i, v in enumerate(vals) if py_eval(op, v)}
        assert got == want and count == len(want), op


def test_oracle_groups_by_long_strings_in_first_appearance_order():
    t, vals = table(n=3000, nulls=True)
    rows = oracle.aggregate(t, None, [AggregateSpec("c", AggregateKind.CountStar())], None, (1,))
    order = []
    for v in vals:
        if v not in order:
            order.append(v)
    assert [k[0] for k, _ in rows] == order
    assert [v[0].value for _, v in rows] == [vals.count(k) for k in order]


SPECS = [AggregateSpec("c", AggregateKind.CountStar()), AggregateSpec("s", AggregateKind.Sum(2, DataType.Int64)),
         AggregateSpec("mx", AggregateKind.Max(2, DataType.Int64))]


@pytest.mark.gpu
@pytest.mark.parametrize("nulls", [False, True])
def test_gpu_long_string_leaves_match_the_oracle(gpu_ctx, nulls):
    from llkv_b200 import gpu
    t, _ = table(nulls=nulls)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t, chunk_rows=4096)
    try:
        assert dt.columns[1].dict_size() == len(set(MODES))
        validity_bytes = sum((min(4096, t.n_rows - lo) + 7) // 8 for lo in range(0, t.n_rows, 4096)) if nulls else 0
        assert dt.columns[1].h2d_bytes() == t.n_rows * 8 + validity_bytes  # codes (+ validity): no string bytes
        for op in leaves():
            flt = Expr.And([Expr.Pred(Filter(1, op)), pred(2, Operator.GreaterThan(-90))])
            prog = gpu.Program(gpu_ctx, flt)
            agg = gpu.Aggregation(dt, SPECS)
            try:
                agg.run(prog)
                got = agg.finalize(1)
                assert agg.run_info().used_fast_kernel == 1, op  # every leaf is an integer leaf over ranks
            finally:
                agg.destroy()
                prog.destroy()
            util.assert_same_result(got, oracle.aggregate(t, flt, SPECS))
        for flt in [Expr.Or([Expr.Pred(Filter(1, Operator.StartsWith("DELIVER"))), Expr.Not(Expr.Pred(Filter(1, Operator.GreaterThan("NONE"))))]),
                    Expr.Not(Expr.Pred(Filter(1, Operator.Contains("AIR"))))]:
            util.assert_same_result(dt.aggregate(flt, SPECS), oracle.aggregate(t, flt, SPECS))
            w_gpu, c_gpu = dt.filter_bitmap(flt)
            w_cpu, c_cpu = oracle.filter_bitmap(t, flt)
            assert c_gpu == c_cpu and np.array_equal(w_gpu, w_cpu)
    finally:
        dt.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("nulls", [False, True])
def test_gpu_group_by_long_strings(gpu_ctx, nulls):
    from llkv_b200 import gpu
    t, _ = table(n=60_000, seed=4, nulls=nulls)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t, chunk_rows=8192)
    try:
        for keys in [(1,), (3, 1), (1, 3)]:
            for flt in (None, pred(2, Operator.LessThan(40))):
                want = oracle.aggregate(t, flt, SPECS, None, keys)
                for mode in (0, 2):
                    gpu_ctx.set_jit(mode)
                    util.assert_same_result(dt.aggregate(flt, SPECS, None, keys, cardinality_hint=128), want)
        # HAVING / ORDER BY over the string key compare the strings, not the codes' packed form
        agg = gpu.Aggregation(dt, SPECS, (1,), cardinality_hint=64)
        agg.set_output(having=[("key", 0, CompareOp.GtEq, "NONE")], order_by=[("key", 0, True, False)])
        agg.run(None)
        rows = agg.finalize(64)
        agg.destroy()
        want = sorted([r for r in oracle.aggregate(t, None, SPECS, None, (1,)) if r[0][0] is not None and r[0][0].encode() >= b"NONE"],
                      key=lambda r: r[0][0].encode(), reverse=True)
        util.assert_same_result(rows, want)
    finally:
        gpu_ctx.set_jit(1)
        dt.destroy()


@pytest.mark.gpu
def test_gpu_column_turns_long_in_a_later_chunk_and_grows_after_seal(gpu_ctx):
    """Chunks of short strings first (packed keys), then a chunk with a long one (the column re-codes what it holds), a scan,
    then more chunks with new entries that sort before the old ones (codes are re-ranked at the next seal)."""
    from llkv_b200 import gpu
    rng = np.random.default_rng(8)
    parts = [["N", "A", "xy", ""], ["R", "N", "TAKE BACK RETURN", "xy"], ["AAAA first in order", "N", "zz", "DELIVER IN PERSON"]]
    vals = []
    dc = None
    dt = gpu.DeviceTable(gpu_ctx, 62)
    num = HostColumn(2, DataType.Int64, np.zeros(0, np.int64))
    try:
        for step, words in enumerate(parts):
            new = [words[i] for i in rng.integers(0, len(words), 5000)]
            col = HostColumn.utf8(1, new)
            if dc is None:
                dc = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(62, 1), col)
                nc = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(62, 2), num)
                dt.columns[1], dt.columns[2] = dc, nc
            dc.append(col, chunk_rows=2048)
            x = rng.integers(-50, 50, 5000, dtype=np.int64)
            nc.append(HostColumn(2, DataType.Int64, x))
            vals += new
            num = HostColumn(2, DataType.Int64, np.concatenate([num.values, x]))
            dt.n_rows = len(vals)
            t = HostTable(62).add(HostColumn.utf8(1, vals)).add(num)
            assert (dc.dict_size() > 0) == (step >= 1)
            for flt in [None, pred(1, Operator.GreaterThanOrEquals("N")), pred(1, Operator.StartsWith("A")), pred(1, Operator.Equals("xy"))]:
                util.assert_same_result(dt.aggregate(flt, SPECS, None, (1,), cardinality_hint=16), oracle.aggregate(t, flt, SPECS, None, (1,)))
    finally:
        dt.destroy()


@pytest.mark.gpu
def test_gpu_clear_drops_the_dictionary(gpu_ctx):
    """clear() + short strings: the column is a packed short-string column again (no dictionary, keys come back inline)."""
    from llkv_b200 import gpu
    long_col = HostColumn.utf8(1, ["DELIVER IN PERSON", "NONE", "DELIVER IN PERSON"])
    dc = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(65, 1), long_col)
    nc = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(65, 2), HostColumn(2, DataType.Int64, np.arange(3, dtype=np.int64)))
    dt = gpu.DeviceTable(gpu_ctx, 65)
    dt.columns[1], dt.columns[2] = dc, nc
    try:
        dc.append(long_col)
        nc.append(HostColumn(2, DataType.Int64, np.arange(3, dtype=np.int64)))
        dt.n_rows = 3
        assert dc.dict_size() == 2
        got = dt.aggregate(None, SPECS[:2], None, (1,), cardinality_hint=4)
        assert [k[0] for k, _ in got] == ["DELIVER IN PERSON", "NONE"]
        dc.clear()
        short = HostColumn.utf8(1, ["N", "xy", "N"])
        dc.append(short)
        assert dc.dict_size() == 0
        t = HostTable(65).add(short).add(HostColumn(2, DataType.Int64, np.arange(3, dtype=np.int64)))
        util.assert_same_result(dt.aggregate(None, SPECS[:2], None, (1,), cardinality_hint=4), oracle.aggregate(t, None, SPECS[:2], None, (1,)))
    finally:
        dt.destroy()


@pytest.mark.gpu
def test_gpu_long_string_limits_are_errors(gpu_ctx):
    from llkv_b200 import gpu
    t, _ = table(n=3000)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        with pytest.raises(LlkvError) as ei:  # Unicode case folding is not ASCII folding
            dt.aggregate(pred(1, Operator.Contains("air", False)), SPECS)
        assert ei.value.code == ffi.ERR_PREDICATE_BUILD
        with pytest.raises(LlkvError) as ei:  # codes of a dictionary do not compare inside scalar expressions
            dt.aggregate(Expr.Compare(ScalarExpr.Column(1), CompareOp.Eq, ScalarExpr.Literal("NONE")), SPECS)
        assert ei.value.code == ffi.ERR_INVALID_ARGUMENT
    finally:
        dt.destroy()
    t, _ = table(n=3000, words=ASCII_MODES, tid=63)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        for op in (Operator.Contains("air", False), Operator.StartsWith("take", False), Operator.EndsWith("RETURN", False)):
            util.assert_same_result(dt.aggregate(pred(1, op), SPECS), oracle.aggregate(t, pred(1, op), SPECS))
    finally:
        dt.destroy()
    # a pattern that matches more than 255 scattered entries
    words = [f"entry number {i:05d} {'x' if i % 2 else 'y'}" for i in range(1200)]
    t, _ = table(n=5000, words=words, tid=64)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        with pytest.raises(LlkvError) as ei:
            dt.aggregate(pred(1, Operator.EndsWith("x")), SPECS)
        assert ei.value.code == ffi.ERR_PREDICATE_BUILD
        flt = pred(1, Operator.StartsWith("entry number 001"))  # a prefix is one range of ranks, however many entries
        util.assert_same_result(dt.aggregate(flt, SPECS), oracle.aggregate(t, flt, SPECS))
        got = dt.aggregate(None, SPECS, None, (1,), cardinality_hint=2000)
        util.assert_same_result(got, oracle.aggregate(t, None, SPECS, None, (1,)))
    finally:
        dt.destroy()
