"""GPU: the partitioned high-cardinality GROUP BY (LeanTile::scatter + partition_apply_kernel) against the oracle and
against the per-row global-table path: results must be identical whatever path runs."""
import numpy as np
import pytest

import util
from llkv_b200 import tpch
from llkv_b200.expr import AggregateKind, AggregateSpec, BinaryOp, DataType, ScalarExpr
from llkv_b200.table import HostColumn, HostTable, decimal_from_i64
from oracle import oracle

pytestmark = pytest.mark.gpu
REL = 1e-12


def run(gpu_ctx, dt, expr, specs, keys, hint, snap=None, lo=0, hi=None, part=2, cap=1 << 17):
    from llkv_b200 import gpu
    gpu_ctx.set_partitioning(part)
    gpu_ctx.set_jit(2)
    prog = gpu.Program(gpu_ctx, expr) if expr is not None else None
    dt.set_snapshot(snap)
    agg = gpu.Aggregation(dt, specs, keys, cardinality_hint=hint)
    try:
        agg.run(prog, snap is not None, lo, dt.n_rows if hi is None else hi)
        got = agg.finalize(cap)
        return got, agg.run_info()
    finally:
        agg.destroy()
        if prog:
            prog.destroy()
        gpu_ctx.set_partitioning(1)
        gpu_ctx.set_jit(1)


def test_partitioned_matches_oracle_and_per_row_path(gpu_ctx):
    from llkv_b200 import gpu
    t = tpch.highcard_table(300_000, 50_000, seed=4)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        want = oracle.aggregate(t, None, tpch.highcard_aggregates(), None, (tpch.K_FIELD,), group_capacity=1 << 17)
        got, info = run(gpu_ctx, dt, None, tpch.highcard_aggregates(), (tpch.K_FIELD,), 50_000)
        assert info.partitions >= 2 and info.used_jit_kernel == 1 and info.kernel_launches == 2
        util.assert_same_result(got, want, REL)
        direct, info0 = run(gpu_ctx, dt, None, tpch.highcard_aggregates(), (tpch.K_FIELD,), 50_000, part=0)
        assert info0.partitions == 0
        util.assert_same_result(direct, want, REL)
        # ragged row ranges: first / last tiles partly selected
        for lo, hi in ((1, 299_999), (12_345, 123_457), (0, 1), (77, 77)):
            want = oracle.aggregate(t, None, tpch.highcard_aggregates(), None, (tpch.K_FIELD,), row_begin=lo, row_end=hi, group_capacity=1 << 17)
            got, _ = run(gpu_ctx, dt, None, tpch.highcard_aggregates(), (tpch.K_FIELD,), 50_000, lo=lo, hi=hi)
            util.assert_same_result(got, want, REL)
    finally:
        dt.destroy()


def test_partitioned_scan_in_several_launches(gpu_ctx, monkeypatch):
    """Long scans run as a sequence of (scan, apply) launch pairs over bounded tuple buffers (2^28 rows each); here the
    bound is lowered so a small table takes five of them."""
    from llkv_b200 import gpu
    monkeypatch.setenv("LLKV_GPU_PART_BATCH_ROWS", "70000")
    t = tpch.highcard_table(300_000, 40_000, seed=9)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        want = oracle.aggregate(t, None, tpch.highcard_aggregates(), None, (tpch.K_FIELD,), group_capacity=1 << 17)
        got, info = run(gpu_ctx, dt, None, tpch.highcard_aggregates(), (tpch.K_FIELD,), 40_000)
        assert info.partitions >= 2 and info.kernel_launches >= 8
        util.assert_same_result(got, want, REL)
    finally:
        dt.destroy()


def test_partition_overflow_goes_to_the_per_row_path(gpu_ctx):
    """Half of the rows share one key: its partition fills up (capacity = uniform share + 25 %) and the surplus tuples are
    applied by the scan itself."""
    from llkv_b200 import gpu
    rng = np.random.default_rng(11)
    n = 400_000
    k = rng.integers(0, 30_000, n, dtype=np.int64)
    k[rng.random(n) < 0.5] = 4242
    v = rng.integers(-1000, 1001, n, dtype=np.int64)
    t = HostTable(1).add(HostColumn(tpch.K_FIELD, DataType.Int64, k)).add(HostColumn(tpch.V_FIELD, DataType.Int64, v))
    specs = tpch.highcard_aggregates() + [AggregateSpec("lo", AggregateKind.Min(tpch.V_FIELD, DataType.Int64)),
                                          AggregateSpec("hi", AggregateKind.Max(tpch.V_FIELD, DataType.Int64))]
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        want = oracle.aggregate(t, None, specs, None, (tpch.K_FIELD,), group_capacity=1 << 17)
        got, info = run(gpu_ctx, dt, None, specs, (tpch.K_FIELD,), 30_000)
        assert info.partitions >= 2
        util.assert_same_result(got, want, REL)
    finally:
        dt.destroy()


def test_partitioned_q1_shape_with_mvcc_and_filter(gpu_ctx):
    """Q1's aggregates (exact decimal products, AVG, COUNT) grouped by a high-cardinality key, with the Q1 filter and an
    MVCC snapshot in front: everything before GROUP runs as usual, only the accumulation is partitioned."""
    from llkv_b200 import gpu
    n = 120_000
    t, snap = tpch.lineitem_table(n, seed=3, with_mvcc=True)
    rng = np.random.default_rng(5)
    t.add(HostColumn(900, DataType.Int64, rng.integers(-5_000, 5_000, n, dtype=np.int64)))
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        want = oracle.aggregate(t, tpch.q1_filter(), tpch.q1_aggregates(), snap, (900,), group_capacity=1 << 15)
        got, info = run(gpu_ctx, dt, tpch.q1_filter(), tpch.q1_aggregates(), (900,), 10_000, snap=snap, cap=1 << 15)
        assert info.partitions >= 2
        util.assert_same_result(got, want, REL)
    finally:
        dt.destroy()


@pytest.mark.parametrize("hint", [131_072, 250_000, 1_048_576])
def test_group_table_sizes_that_end_on_an_allocation_boundary(gpu_ctx, hint):
    """2^k slots x 8 B of keys end exactly on a 2 MiB / 4 MiB / 16 MiB boundary: initialising the table must not touch the
    key array past its end (the two spare rows have accumulator words but no key slot)."""
    from llkv_b200 import gpu
    t = tpch.highcard_table(50_000, 10_000, seed=2)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        want = oracle.aggregate(t, None, tpch.highcard_aggregates(), None, (tpch.K_FIELD,), group_capacity=1 << 17)
        for part in (0, 2):
            got, _ = run(gpu_ctx, dt, None, tpch.highcard_aggregates(), (tpch.K_FIELD,), hint, part=part)
            util.assert_same_result(got, want, REL)
    finally:
        dt.destroy()


# ---------------------------------------------------------------------------------------------------- packed form
def packed_table(kind: str, n: int, seed: int):
    rng = np.random.default_rng(seed)
    if kind == "dense":       # keys fill a range: partitions are key ranges, slots indexed directly
        k = rng.integers(-3_000, 47_000, n, dtype=np.int64)
        hint = 50_000
    elif kind == "sparse":    # 30 000 distinct keys spread over 2^20: partitions are hash ranges
        pool = rng.integers(0, 1 << 20, 30_000, dtype=np.int64)
        k = pool[rng.integers(0, pool.size, n)]
        hint = 30_000
    else:                     # half of the rows on one key: its partition overflows into the per-row path
        k = rng.integers(0, 30_000, n, dtype=np.int64)
        k[rng.random(n) < 0.5] = 4242
        hint = 30_000
    v = rng.integers(0, 1001, n, dtype=np.int64)
    # (key, 21+ bits of row and both operands share 64 bits; on the skewed table the hot group's sum of w passes 2^32: the carry word)
    w = rng.integers(0, 1 << (12 if kind == "sparse" else 17), n, dtype=np.int64)
    t = HostTable(1).add(HostColumn(tpch.K_FIELD, DataType.Int64, k)).add(HostColumn(tpch.V_FIELD, DataType.Int64, v))
    t.add(HostColumn(3, DataType.Int64, w))
    return t, hint


@pytest.mark.parametrize("kind", ["dense", "sparse", "skewed"])
def test_packed_partitions_match_oracle(gpu_ctx, kind):
    """Partitioned GROUP BY with 64-bit packed tuples (LeanTile::scatter_packed + partition_fold_kernel: every partition is
    aggregated in shared memory) against the oracle, the first partitioned form and the per-row path."""
    from llkv_b200 import gpu
    t, hint = packed_table(kind, 400_000, seed=21)
    specs = tpch.highcard_aggregates() + [AggregateSpec("w", AggregateKind.Sum(3, DataType.Int64)), AggregateSpec("n", AggregateKind.Count(3))]
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        want = oracle.aggregate(t, None, specs, None, (tpch.K_FIELD,), group_capacity=1 << 17)
        got, info = run(gpu_ctx, dt, None, specs, (tpch.K_FIELD,), hint)
        assert info.packed_tuples == (2 if kind != "sparse" else 1) and info.partitions >= 2 and info.used_jit_kernel == 1, (info.packed_tuples, info.partitions)
        util.assert_same_result(got, want, REL)
        first_form, info1 = run(gpu_ctx, dt, None, specs, (tpch.K_FIELD,), hint, part=3)
        assert info1.packed_tuples == 0 and info1.partitions >= 2
        util.assert_same_result(first_form, want, REL)
        # a filter in front, ragged row ranges
        flt = tpch.between_filter(tpch.V_FIELD, 100, 800)
        for lo, hi in ((0, 400_000), (3, 399_990), (123_456, 123_999)):
            want = oracle.aggregate(t, flt, specs, None, (tpch.K_FIELD,), row_begin=lo, row_end=hi, group_capacity=1 << 17)
            got, _ = run(gpu_ctx, dt, flt, specs, (tpch.K_FIELD,), hint, lo=lo, hi=hi)
            util.assert_same_result(got, want, REL)
    finally:
        dt.destroy()


def test_packed_partitions_in_several_launches_accumulate(gpu_ctx, monkeypatch):
    """Every launch folds its partitions into the same global table: groups met again in a later launch are added to."""
    from llkv_b200 import gpu
    monkeypatch.setenv("LLKV_GPU_PART_BATCH_ROWS", "90000")
    t, hint = packed_table("dense", 400_000, seed=22)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        want = oracle.aggregate(t, None, tpch.highcard_aggregates(), None, (tpch.K_FIELD,), group_capacity=1 << 17)
        got, info = run(gpu_ctx, dt, None, tpch.highcard_aggregates(), (tpch.K_FIELD,), hint)
        assert info.packed_tuples == 2 and info.kernel_launches >= 8
        util.assert_same_result(got, want, REL)
    finally:
        dt.destroy()


def test_plans_the_packed_form_cannot_carry_keep_the_first_form(gpu_ctx):
    """Signed SUM operands and MIN / MAX have no place in a packed tuple: those plans run partitioned in the first form."""
    from llkv_b200 import gpu
    rng = np.random.default_rng(23)
    n = 200_000
    t = HostTable(1).add(HostColumn(tpch.K_FIELD, DataType.Int64, rng.integers(0, 20_000, n, dtype=np.int64)))
    t.add(HostColumn(tpch.V_FIELD, DataType.Int64, rng.integers(-1000, 1001, n, dtype=np.int64)))
    specs = tpch.highcard_aggregates() + [AggregateSpec("lo", AggregateKind.Min(tpch.V_FIELD, DataType.Int64))]
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        want = oracle.aggregate(t, None, specs, None, (tpch.K_FIELD,), group_capacity=1 << 17)
        got, info = run(gpu_ctx, dt, None, specs, (tpch.K_FIELD,), 20_000)
        assert info.packed_tuples == 0 and info.partitions >= 2
        util.assert_same_result(got, want, REL)
    finally:
        dt.destroy()
