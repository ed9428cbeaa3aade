"""GPU: randomised differential test.  Random (seeded) predicates, aggregate lists, GROUP BY keys, cardinality hints and
row ranges over a mixed-type table without NULLs, so most plans take the lean kernel; each plan runs interpreted and
specialised and both must equal the oracle (bit-exact integers / decimals / counts / min / max, 1e-12 for f64)."""
import numpy as np
import pytest

import util
from llkv_b200.expr import AggregateKind, AggregateSpec, Bound, CompareOp, DataType, Expr, Literal, Operator, ScalarExpr, pred
from llkv_b200 import tpch
from llkv_b200.table import LlkvError, Snapshot
from oracle import oracle
from test_gpu_parity import REL, device_table, mixed_table

pytestmark = pytest.mark.gpu
D = DataType.Decimal128(15, 2)


def random_filter(rng):
    leaves = [
        lambda: pred(1, Operator.Range(Bound.Included(int(rng.integers(-900, 0))), Bound.Excluded(int(rng.integers(0, 900))))),
        lambda: pred(2, Operator.LessThan(int(rng.integers(-40, 50)))),
        lambda: pred(4, Operator.GreaterThanOrEquals(int(rng.integers(0, 400)))),
        lambda: pred(5, Operator.Range(Bound.Included(Literal.Decimal128(int(rng.integers(-10**7, 0)), 2)), Bound.Included(Literal.Decimal128(int(rng.integers(0, 10**7)), 2)))),
        lambda: pred(6, Operator.LessThanOrEquals(Literal.Date32(int(rng.integers(8000, 11000))))),
        lambda: pred(2, Operator.GreaterThan(int(rng.integers(-50, 30)))),
        lambda: pred(9, Operator.Equals(bool(rng.integers(0, 2)))),
    ]
    k = int(rng.integers(0, 4))
    if k == 0:
        return None
    picks = [leaves[i]() for i in rng.choice(len(leaves), size=k, replace=False)]
    return picks[0] if k == 1 else Expr.And(picks)


def random_aggs(rng):
    c = ScalarExpr.Column
    pool = [
        AggregateSpec("n", AggregateKind.CountStar()),
        AggregateSpec("c2", AggregateKind.Count(2)),
        AggregateSpec("s1", AggregateKind.Sum(1, DataType.Int64)),
        AggregateSpec("a1", AggregateKind.Avg(1, DataType.Int64)),
        AggregateSpec("mn1", AggregateKind.Min(1, DataType.Int64)),
        AggregateSpec("mx8", AggregateKind.Max(c(8) * c(2), DataType.Int64)),
        AggregateSpec("s3", AggregateKind.Sum(3, DataType.Float64)),
        AggregateSpec("mn3", AggregateKind.Min(3, DataType.Float64)),
        AggregateSpec("mx3", AggregateKind.Max(c(3) * c(1), DataType.Float64)),
        AggregateSpec("s5", AggregateKind.Sum(5, D)),
        AggregateSpec("a5", AggregateKind.Avg(5, D)),
        AggregateSpec("mn5", AggregateKind.Min(5, D)),
        AggregateSpec("mx5", AggregateKind.Max(5, D)),
        AggregateSpec("sx", AggregateKind.Sum(c(1) * c(2) + c(8) - 7, DataType.Int64)),
        AggregateSpec("sdd", AggregateKind.Sum(c(5) * c(5), DataType.Decimal128(38, 4))),
        AggregateSpec("sdi", AggregateKind.Sum(c(5) * (1 - c(5)), DataType.Decimal128(38, 4))),
        AggregateSpec("sf", AggregateKind.Sum(c(3) * c(2), DataType.Float64)),
        AggregateSpec("t4", AggregateKind.Total(4, DataType.UInt64)),
    ]
    k = int(rng.integers(1, 7))
    return [pool[i] for i in sorted(rng.choice(len(pool), size=k, replace=False))]


@pytest.mark.parametrize("seed", range(10))
def test_random_plans_match_oracle(gpu_ctx, seed):
    from llkv_b200 import gpu
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(3000, 40000))
    t = mixed_table(n, seed=seed + 77, long_strings=False)
    c_by, d_by, snap_all = tpch.mvcc_arrays(n, seed=seed)
    t.add_mvcc(c_by, d_by)
    dt = device_table(gpu_ctx, t)
    lean_runs = 0
    part_runs = 0
    try:
        for trial in range(5):
            f = random_filter(rng)
            specs = random_aggs(rng)
            keys = [(), (), (10,), (9,), (2,), (9, 10), (8,), (6,)][int(rng.integers(0, 8))]
            hint = int(rng.choice([0, 2, 6, 100, 5000])) if keys else 0
            lo = int(rng.integers(0, n // 3)) if rng.random() < 0.5 else 0
            hi = int(rng.integers(2 * n // 3, n + 1)) if rng.random() < 0.5 else n
            snap = [None, snap_all, Snapshot(77, 100, (77, 101))][int(rng.integers(0, 3))]
            ctx_note = f"seed {seed} trial {trial} keys {keys} hint {hint} rows [{lo},{hi}) snapshot {snap} aggs {[s.alias for s in specs]}"
            try:
                want = oracle.aggregate(t, f, specs, snap, keys, row_begin=lo, row_end=hi, group_capacity=1 << 15)
            except LlkvError as e:
                want = e
            for mode, part in ((0, 1), (2, 1), (2, 2)):  # interpreted, specialised, specialised + partitioned GROUP BY
                if part == 2 and hint <= 128:
                    continue
                gpu_ctx.set_jit(mode)
                gpu_ctx.set_partitioning(part)
                prog = gpu.Program(gpu_ctx, f) if f is not None else None
                dt.set_snapshot(snap)
                agg = gpu.Aggregation(dt, specs, keys, cardinality_hint=hint)
                try:
                    if isinstance(want, LlkvError):
                        with pytest.raises(LlkvError) as got:
                            agg.run(prog, snap is not None, lo, hi)
                            agg.finalize(1 << 15)
                        assert got.value.code == want.code, ctx_note
                    else:
                        agg.run(prog, snap is not None, lo, hi)
                        got = agg.finalize(1 << 15)
                        lean_runs += agg.run_info().used_fast_kernel
                        part_runs += 1 if agg.run_info().partitions else 0
                        try:
                            util.assert_same_result(got, want, REL)
                        except AssertionError as e:
                            raise AssertionError(f"{ctx_note} jit={mode} partitioning={part}: {e}") from e
                finally:
                    agg.destroy()
                    if prog:
                        prog.destroy()
        assert lean_runs >= 2  # the point of this test is the lean kernel (Int16 arguments stay on the general interpreter)
    finally:
        gpu_ctx.set_jit(1)
        gpu_ctx.set_partitioning(1)
        dt.destroy()


# ---------------------------------------------------------------------------------- OR / NOT trees, NULLs: still the lean kernel
def random_tree(rng, depth=0, compares=False):
    """AND / OR / NOT trees over typed leaves on integer, float, decimal, date, boolean and short-string columns — ranges,
    equalities, IN lists, prefixes, IS [NOT] NULL (three-valued logic with
    domains when the columns are nullable: llkv-scan/src/predicate.rs:167-186,665-777)."""
    leaves = [
        lambda: pred(1, Operator.Range(Bound.Included(int(rng.integers(-900, 0))), Bound.Excluded(int(rng.integers(0, 900))))),
        lambda: pred(2, Operator.LessThan(int(rng.integers(-40, 50)))),
        lambda: pred(4, Operator.GreaterThanOrEquals(int(rng.integers(0, 400)))),
        lambda: pred(5, Operator.Range(Bound.Included(Literal.Decimal128(int(rng.integers(-10**7, 0)), 2)), Bound.Included(Literal.Decimal128(int(rng.integers(0, 10**7)), 2)))),
        lambda: pred(6, Operator.LessThanOrEquals(Literal.Date32(int(rng.integers(8000, 11000))))),
        lambda: pred(1, Operator.Equals(int(rng.integers(-5, 5)))),
        lambda: pred(9, Operator.Equals(bool(rng.integers(0, 2)))),
        lambda: pred(5, Operator.IsNull),
        lambda: pred(2, Operator.IsNotNull),
        lambda: pred(4, Operator.Range(Bound.Unbounded, Bound.Unbounded)),
        lambda: pred(3, Operator.Range(Bound.Excluded(float(np.round(rng.normal(-50, 40), 3))), Bound.Included(float(np.round(rng.normal(60, 40), 3))))),
        lambda: pred(3, Operator.GreaterThan(float(rng.choice([0.0, -0.0, 12.5, -75.125, float("inf"), float("-inf"), float("nan")])))),
        lambda: pred(7, Operator.LessThanOrEquals(float(rng.choice([0.0, -0.0, 1.25, -3.3, 4.75])))),
        lambda: pred(7, Operator.Equals(float(rng.choice([0.0, -0.0, 0.25, 0.3])))),
        lambda: pred(1, Operator.In([int(x) for x in rng.integers(-1000, 1000, int(rng.integers(0, 12)))])),
        lambda: pred(5, Operator.In([Literal.Decimal128(int(x), 2) for x in rng.integers(-10**7, 10**7, 3)] + [Literal.Int128(7), Literal.Decimal128(12345, 3)])),
        lambda: pred(6, Operator.In([Literal.Date32(int(x)) for x in rng.integers(8000, 11000, 5)])),
        lambda: pred(10, Operator.In(["A", "xy", "zz"])),
        lambda: pred(10, Operator.StartsWith(str(rng.choice(["", "a", "ab", "x", "N"])))),
    ]
    if compares:
        # general comparisons of two scalar expressions (compute_compare, llkv-compute/src/kernels.rs:269-297): a NULL on
        # either side is NULL; integers, decimals (rescaled to the common scale), dates, floats by total order
        c = ScalarExpr.Column
        ops = [CompareOp.Eq, CompareOp.NotEq, CompareOp.Lt, CompareOp.LtEq, CompareOp.Gt, CompareOp.GtEq]
        op = lambda: ops[int(rng.integers(0, len(ops)))]
        leaves = leaves[:6] + [
            lambda: Expr.Compare(c(1) + c(2), op(), ScalarExpr.Literal(int(rng.integers(-300, 300)))),
            lambda: Expr.Compare(c(1), op(), c(2) * int(rng.integers(-20, 20))),
            lambda: Expr.Compare(c(2) * c(2) - c(1), op(), c(1)),
            lambda: Expr.Compare(c(3), op(), c(1)),
            lambda: Expr.Compare(c(3) * 0.5, op(), c(7)),
            lambda: Expr.Compare(c(5), op(), ScalarExpr.Literal(Literal.Decimal128(int(rng.integers(-10**6, 10**6)), 2))),
            lambda: Expr.Compare(c(5) * c(5), op(), ScalarExpr.Literal(Literal.Decimal128(int(rng.integers(0, 10**12)), 4))),
            lambda: Expr.Compare(c(6), op(), ScalarExpr.Literal(Literal.Date32(int(rng.integers(8000, 11000))))),
            lambda: Expr.Compare(ScalarExpr.Literal(int(rng.integers(-50, 50))), op(), c(2)),
            # both sides computed: the left one waits in a temporary while the right one takes the accumulator
            lambda: Expr.Compare(c(1) + c(2), op(), c(2) * c(2) - int(rng.integers(0, 500))),
            lambda: Expr.Compare(c(5) * c(5), op(), c(5) * int(rng.integers(-100, 100)) * int(rng.integers(1, 1000))),
            # Int64 against UInt64 compares as Float64 (the unsigned side is proven below 2^63)
            lambda: Expr.Compare(c(1) * 2, op(), c(4)),
            # float IN lists: IEEE equality (a zero of either sign matches both, NaN matches nothing)
            lambda: pred(7, Operator.In([float(x) for x in rng.choice([0.0, -0.0, 0.25, -1.0, 3.5, 4.75], size=int(rng.integers(0, 5)))])),
            lambda: pred(3, Operator.In([float(x) for x in rng.choice([0.0, -0.0, 12.5, float("nan"), float("inf")], size=3)])),
            # Expr::IsNull over an expression: NULL where any column it reads is NULL; never NULL itself
            lambda: Expr.IsNull(c(1) + c(2), negated=bool(rng.integers(0, 2))),
            lambda: Expr.IsNull(c(5) * c(5) - c(5), negated=bool(rng.integers(0, 2))),
        ]
    r = rng.random()
    if depth >= 3 or r < 0.35:
        return leaves[int(rng.integers(0, len(leaves)))]()
    if r < 0.5:
        return Expr.Not(random_tree(rng, depth + 1, compares))
    kids = [random_tree(rng, depth + 1, compares) for _ in range(int(rng.integers(2, 4)))]
    return Expr.And(kids) if r < 0.75 else Expr.Or(kids)


@pytest.mark.parametrize("nulls", [False, True], ids=["dense", "nullable"])
@pytest.mark.parametrize("seed", range(6))
def test_random_predicate_trees_stay_on_the_lean_kernel(gpu_ctx, seed, nulls):
    _random_trees(gpu_ctx, seed, nulls, compares=False)


@pytest.mark.parametrize("nulls", [False, True], ids=["dense", "nullable"])
@pytest.mark.parametrize("seed", range(3))
def test_random_trees_with_expression_comparisons_stay_on_the_lean_kernel(gpu_ctx, seed, nulls):
    """Expr::Compare leaves (two scalar expressions) inside AND / OR / NOT trees: lowered to the lean kernel's FO_CMP, checked
    against the oracle through aggregates (both builds of the lean kernel) and through the general interpreter's bitmap."""
    _random_trees(gpu_ctx, 40 + seed, nulls, compares=True)


def _random_trees(gpu_ctx, seed, nulls, compares):
    from llkv_b200 import gpu
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(5000, 30000))
    t = mixed_table(n, seed=seed + 31, nulls=nulls, long_strings=False)
    c_by, d_by, snap_all = tpch.mvcc_arrays(n, seed=seed)
    t.add_mvcc(c_by, d_by)
    c = ScalarExpr.Column
    pool = [AggregateSpec("n", AggregateKind.CountStar()), AggregateSpec("c2", AggregateKind.Count(2)), AggregateSpec("s1", AggregateKind.Sum(1, DataType.Int64)),
            AggregateSpec("a1", AggregateKind.Avg(1, DataType.Int64)), AggregateSpec("mn1", AggregateKind.Min(1, DataType.Int64)),
            AggregateSpec("s5", AggregateKind.Sum(5, D)), AggregateSpec("a5", AggregateKind.Avg(5, D)), AggregateSpec("mx5", AggregateKind.Max(5, D)),
            AggregateSpec("sx", AggregateKind.Sum(c(1) * c(2) - 7, DataType.Int64)), AggregateSpec("nn", AggregateKind.CountNulls(5)),
            AggregateSpec("sdd", AggregateKind.Sum(c(5) * c(5), DataType.Decimal128(38, 4)))]
    dt = device_table(gpu_ctx, t)
    try:
        for trial in range(6):
            f = random_tree(rng, compares=compares)
            specs = [pool[i] for i in sorted(rng.choice(len(pool), size=int(rng.integers(1, 6)), replace=False))]
            snap = [None, snap_all][int(rng.integers(0, 2))]
            lo = int(rng.integers(0, n // 3)) if rng.random() < 0.5 else 0
            want = oracle.aggregate(t, f, specs, snap, (), row_begin=lo, row_end=n)
            w_bits, w_count = oracle.filter_bitmap(t, f, snap, lo, n)
            for mode in (0, 2):
                gpu_ctx.set_jit(mode)
                prog = gpu.Program(gpu_ctx, f)
                dt.set_snapshot(snap)
                agg = gpu.Aggregation(dt, specs)
                try:
                    agg.run(prog, snap is not None, lo, n)
                    got = agg.finalize(1)
                    info = agg.run_info()
                    util.assert_same_result(got, want, REL)
                    assert info.used_fast_kernel == 1, f"seed {seed} trial {trial}: the plan left the lean kernel: {f}"
                finally:
                    agg.destroy()
                    prog.destroy()
            g_bits, g_count = dt.filter_bitmap(f, snap, lo, n)  # (the general interpreter: the same three-valued logic)
            assert g_count == w_count and np.array_equal(g_bits, w_bits)
    finally:
        gpu_ctx.set_jit(1)
        dt.destroy()


def test_float_leaves_on_the_lean_kernel_follow_partial_cmp(gpu_ctx):
    """Float predicates compare by partial_cmp (typed_predicate.rs:75-142): NaN matches nothing — as a value or as a literal —
    and -0 == +0.  The lean kernel evaluates them as integer ranges over the order-preserving image of the bits; every
    special value on both sides, f64 and f32, against the oracle."""
    from llkv_b200 import gpu
    from llkv_b200.table import HostColumn, HostTable
    special = np.array([np.nan, -np.nan, np.inf, -np.inf, 0.0, -0.0, 5e-324, -5e-324, 1.5, -1.5, 1e300, -1e300, 3.0, 2.9999999999999996])
    rng = np.random.default_rng(4)
    n = 20_000
    v64 = np.where(rng.random(n) < 0.5, special[rng.integers(0, len(special), n)], rng.normal(0, 2, n))
    v32 = v64.astype(np.float32)
    t = HostTable(51).add(HostColumn(1, DataType.Float64, v64)).add(HostColumn(2, DataType.Float32, v32)) \
                     .add(HostColumn(3, DataType.Int64, rng.integers(-9, 9, n, dtype=np.int64)))
    dt = device_table(gpu_ctx, t)
    specs = [AggregateSpec("n", AggregateKind.CountStar()), AggregateSpec("s", AggregateKind.Sum(3, DataType.Int64))]
    lits = [float("nan"), float("inf"), float("-inf"), 0.0, -0.0, 5e-324, -5e-324, 1.5, -1.5, 3.0, 1e300, 1e-50]
    try:
        for col in (1, 2):
            ops = []
            for a in lits:
                if col == 2 and abs(a) > 3e38 and a == a and abs(a) != float("inf"):
                    continue  # (out of range for an f32 literal: PredicateBuild on both sides, covered by the parity suite)
                ops += [Operator.Equals(a), Operator.GreaterThan(a), Operator.GreaterThanOrEquals(a), Operator.LessThan(a), Operator.LessThanOrEquals(a)]
                for b in (0.0, 1.5, float("inf"), float("nan")):
                    ops.append(Operator.Range(Bound.Excluded(a), Bound.Included(b)))
                    ops.append(Operator.Range(Bound.Included(a), Bound.Excluded(b)))
            for op in ops:
                f = Expr.And([pred(col, op), pred(3, Operator.GreaterThan(-5))])
                try:
                    want = oracle.aggregate(t, f, specs)
                except LlkvError as e:
                    with pytest.raises(LlkvError) as got:
                        dt.aggregate(f, specs)
                    assert got.value.code == e.code
                    continue
                prog = gpu.Program(gpu_ctx, f)
                agg = gpu.Aggregation(dt, specs)
                try:
                    agg.run(prog)
                    got = agg.finalize(1)
                    assert agg.run_info().used_fast_kernel == 1, op
                    util.assert_same_result(got, want, REL)
                finally:
                    agg.destroy()
                    prog.destroy()
    finally:
        dt.destroy()
