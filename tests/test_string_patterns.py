"""Utf8 StartsWith / EndsWith / Contains leaves (llkv-expr/src/typed_predicate.rs:187-209 — Rust's
str::starts_with / ends_with / contains, both sides through to_lowercase() when case-insensitive).
CPU half: the oracle against Python's str methods, which are the same byte-wise definitions.  GPU half: the device path
against the oracle, as conjuncts (StartsWith is a range of packed keys on the specialised kernel), inside OR / NOT trees and
over NULLs, plus the errors the path raises instead of answering wrongly."""
import numpy as np
import pytest

import util
from llkv_b200 import ffi
from llkv_b200.expr import AggregateKind, AggregateSpec, DataType, Expr, Filter, Operator
from llkv_b200.table import HostColumn, HostTable, LlkvError
from oracle import oracle

WORDS = ["", "a", "A", "ab", "Ab", "abc", "bca", "cab", "abcabca", "aaaaaaa", "aab", "baa", "xyz", "XYZ", "a\x00", "a\x00b", "\x00",
         "Abc", "zzzzzzz", "ba", "b", "c"]
PATTERNS = ["", "a", "ab", "abc", "ca", "A", "aB", "aa", "aaaaaaa", "aaaaaaaa", "abcabcab", "z", "\x00", "a\x00", "bc", "xyz", "b"]


def str_table(n=5000, seed=1, nulls=False, words=WORDS):
    rng = np.random.default_rng(seed)
    vals = [words[i] for i in rng.integers(0, len(words), n)]
    col = HostColumn.utf8(1, vals)
    if nulls:
        valid = rng.random(n) > 0.2
        col.validity = np.packbits(valid, bitorder="little")
        vals = [v if ok else None for v, ok in zip(vals, valid)]
    t = HostTable(41).add(col).add(HostColumn(2, DataType.Int64, rng.integers(-100, 100, n, dtype=np.int64)))
    return t, vals


def py_match(kind, v, pat, cs):
    if v is None:
        return False
    if not cs:
        v, pat = v.lower(), pat.lower()
    return {"starts": v.startswith, "ends": v.endswith, "contains": lambda p: p in v}[kind](pat)


def make_op(kind, pat, cs=True):
    return {"starts": Operator.StartsWith, "ends": Operator.EndsWith, "contains": Operator.Contains}[kind](pat, cs)


@pytest.mark.parametrize("kind", ["starts", "ends", "contains"])
@pytest.mark.parametrize("cs", [True, False])
@pytest.mark.parametrize("nulls", [False, True])
def test_oracle_matches_str_methods(kind, cs, nulls):
    t, vals = str_table(nulls=nulls)
    for pat in PATTERNS:
        words, count = oracle.filter_bitmap(t, Expr.Pred(Filter(1, make_op(kind, pat, cs))))
        got = set(util.selected_positions(words, t.n_rows))
        want = {i for i, v in enumerate(vals) if py_match(kind, v, pat, cs)}
        assert got == want and count == len(want), (kind, pat, cs)


def test_oracle_pattern_on_a_non_string_column_never_matches():
    t, _ = str_table()
    words, count = oracle.filter_bitmap(t, Expr.Pred(Filter(2, Operator.Contains("1"))))
    assert count == 0


def test_oracle_case_insensitive_refuses_non_ascii():
    t, _ = str_table(words=["é", "abc"])
    with pytest.raises(LlkvError) as ei:
        oracle.filter_bitmap(t, Expr.Pred(Filter(1, Operator.Contains("a", False))))
    assert ei.value.code == ffi.ERR_PREDICATE_BUILD
    with pytest.raises(LlkvError):
        oracle.filter_bitmap(str_table()[0], Expr.Pred(Filter(1, Operator.Contains("é", False))))
    # case-sensitive matching is byte-wise: non-ASCII data is fine
    _, count = oracle.filter_bitmap(t, Expr.Pred(Filter(1, Operator.Contains("é"))))
    assert count == t.n_rows - oracle.filter_bitmap(t, Expr.Pred(Filter(1, Operator.Equals("abc"))))[1]


SPECS = [AggregateSpec("c", AggregateKind.CountStar()), AggregateSpec("s", AggregateKind.Sum(2, DataType.Int64)),
         AggregateSpec("mn", AggregateKind.Min(2, DataType.Int64))]


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["starts", "ends", "contains"])
@pytest.mark.parametrize("cs", [True, False])
@pytest.mark.parametrize("nulls", [False, True])
def test_gpu_matches_the_oracle(gpu_ctx, kind, cs, nulls):
    from llkv_b200 import gpu
    t, _ = str_table(n=40_000, seed=5, nulls=nulls)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        for pat in PATTERNS:
            flt = Expr.Pred(Filter(1, make_op(kind, pat, cs)))
            util.assert_same_result(dt.aggregate(flt, SPECS), oracle.aggregate(t, flt, SPECS))
            w_gpu, c_gpu = dt.filter_bitmap(flt)
            w_cpu, c_cpu = oracle.filter_bitmap(t, flt)
            assert c_gpu == c_cpu and np.array_equal(w_gpu, w_cpu), (kind, pat, cs)
    finally:
        dt.destroy()


@pytest.mark.gpu
def test_gpu_patterns_inside_predicate_trees(gpu_ctx):
    from llkv_b200 import gpu
    t, _ = str_table(n=30_000, seed=9, nulls=True)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    P = lambda op: Expr.Pred(Filter(1, op))
    trees = [
        Expr.Or([P(Operator.StartsWith("ab")), P(Operator.EndsWith("ca"))]),
        Expr.Not(P(Operator.Contains("a"))),
        Expr.And([P(Operator.StartsWith("a")), Expr.Not(P(Operator.EndsWith("a"))), Expr.Pred(Filter(2, Operator.GreaterThan(0)))]),
        Expr.Or([Expr.Not(P(Operator.StartsWith("a", False))), P(Operator.Equals("abc"))]),
        Expr.And([P(Operator.StartsWith("a")), Expr.Pred(Filter(2, Operator.LessThan(50)))]),
    ]
    try:
        for flt in trees:
            util.assert_same_result(dt.aggregate(flt, SPECS), oracle.aggregate(t, flt, SPECS))
    finally:
        dt.destroy()


@pytest.mark.gpu
def test_gpu_starts_with_runs_on_the_specialised_kernel_and_one_byte_strings(gpu_ctx):
    """A case-sensitive prefix is a range of packed keys: the plan stays on the lean kernel; one-byte string columns
    (resident as one byte per row) answer the same way."""
    from llkv_b200 import gpu
    for words in (WORDS, ["A", "N", "R", "a"]):
        t, _ = str_table(n=50_000, seed=2, words=words)
        dt = gpu.DeviceTable.from_host(gpu_ctx, t)
        try:
            for pat in ["a", "A", "", "ab", "N"]:
                flt = Expr.And([Expr.Pred(Filter(1, Operator.StartsWith(pat))), Expr.Pred(Filter(2, Operator.LessThan(50)))])
                prog = gpu.Program(gpu_ctx, flt)
                agg = gpu.Aggregation(dt, SPECS)
                agg.run(prog)
                got = agg.finalize(1)
                assert agg.run_info().used_fast_kernel == 1
                util.assert_same_result(got, oracle.aggregate(t, flt, SPECS))
                agg.destroy()
                prog.destroy()
                for kind in ("ends", "contains"):
                    f2 = Expr.Pred(Filter(1, make_op(kind, pat, False)))
                    util.assert_same_result(dt.aggregate(f2, SPECS), oracle.aggregate(t, f2, SPECS))
        finally:
            dt.destroy()


@pytest.mark.gpu
def test_gpu_pattern_errors(gpu_ctx):
    from llkv_b200 import gpu
    t, _ = str_table(n=1000, words=["é", "abc"])
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        with pytest.raises(LlkvError) as ei:
            dt.aggregate(Expr.Pred(Filter(1, Operator.Contains("a", False))), SPECS)
        assert ei.value.code == ffi.ERR_PREDICATE_BUILD
        with pytest.raises(LlkvError) as ei:
            dt.aggregate(Expr.Pred(Filter(1, Operator.Contains("é", False))), SPECS)
        assert ei.value.code == ffi.ERR_PREDICATE_BUILD
        with pytest.raises(LlkvError) as ei:
            dt.aggregate(Expr.Pred(Filter(1, Operator.Contains(7))), SPECS)
        assert ei.value.code == ffi.ERR_PREDICATE_BUILD
        flt = Expr.Pred(Filter(1, Operator.Contains("b")))  # case-sensitive over non-ASCII data: byte-wise, fine
        util.assert_same_result(dt.aggregate(flt, SPECS), oracle.aggregate(t, flt, SPECS))
        flt = Expr.Pred(Filter(2, Operator.Contains("1")))  # not a string column: never matches
        util.assert_same_result(dt.aggregate(flt, SPECS), oracle.aggregate(t, flt, SPECS))
    finally:
        dt.destroy()
