"""GPU: row-id-sparse tables and nullable plans on the lean kernel.

Chunks arrive with their row-id shadow columns: rows appended after deletes, last-writer-wins updates, columns that skip the
rows where they are NULL.  The resident image keeps columns by position = row id - first row id with validity bitmaps for the
rows nobody holds; the reference's own fixtures (tests/golden/sparse_known_answers.json) and the oracle over the compacted
table (only the rows that exist, in row-id order) say what every query must return."""
import json
import os

import numpy as np
import pytest

import util
from llkv_b200 import ffi, tpch
from llkv_b200.expr import AggregateKind, AggregateSpec, DataType, Expr, Operator, ScalarExpr, pred
from llkv_b200.table import HostColumn, HostTable, Snapshot, decimal_from_i64, pack_validity
from oracle import oracle

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sparse_known_answers.json")))


def test_fragmented_sum_with_deletes_matches_the_reference_fixture(gpu_ctx):
    """column_sum_bench.rs: 1000 appends x 1000 rows, every 10th row deleted, 1000 more rows: the scan sums what is left."""
    from llkv_b200 import gpu
    f = G["fragmented_sum"]
    n_chunks, rows = f["chunks"], f["chunk_rows"]
    n = n_chunks * rows
    dc = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(21, 1), HostColumn(1, DataType.Int64, np.zeros(0, np.int64)))
    dt = gpu.DeviceTable(gpu_ctx, 21)
    dt.columns[1] = dc
    try:
        for i in range(n_chunks):
            ids = np.arange(i * rows, (i + 1) * rows, dtype=np.uint64)
            dc.append_rows(ids.astype(np.int64), ids)
        dc.delete_rows(np.arange(0, n, f["delete_step"], dtype=np.uint64))
        ids = np.arange(n, n + rows, dtype=np.uint64)
        dc.append_rows(ids.astype(np.int64), ids)
        dc.seal()
        dt.n_rows = dc.rows()
        assert dc.rows() == n + rows and dc.present_rows() == f["expected_rows"]
        # the bench's own visitor: PrimitiveVisitor::u64_chunk summing every chunk (column_sum_bench.rs:240-277), then the
        # with-row-ids form: row id == value in this fixture
        acc = {"sum": 0, "rows": 0, "chunks": 0}

        def on_chunk(vals, rids=None):
            acc["sum"] += int(vals.sum())
            acc["rows"] += vals.shape[0]
            acc["chunks"] += 1
            if rids is not None:
                assert np.array_equal(rids.astype(np.int64), vals)

        dc.visit(on_chunk)
        assert acc["sum"] == f["expected_final_sum"] and acc["rows"] == f["expected_rows"] and acc["chunks"] == 8  # 131 072-row chunks
        acc.update(sum=0, rows=0, chunks=0)
        dc.visit(on_chunk, chunk_rows=50_000, with_row_ids=True)
        assert acc["sum"] == f["expected_final_sum"] and acc["rows"] == f["expected_rows"] and acc["chunks"] == 21
        specs = [AggregateSpec("s", AggregateKind.Sum(1, DataType.Int64)), AggregateSpec("n", AggregateKind.CountStar()),
                 AggregateSpec("c", AggregateKind.Count(1)), AggregateSpec("mn", AggregateKind.Min(1, DataType.Int64))]
        for jit in (1, 2):  # interpreted, then specialised
            gpu_ctx.set_jit(jit)
            agg = gpu.Aggregation(dt, specs)
            agg.run(None)
            s, cnt, c, mn = agg.finalize(1)[0][1]
            info = agg.run_info()
            agg.destroy()
            assert s.value == f["expected_final_sum"] and cnt.value == f["expected_rows"] == c.value and mn.value == 1
            assert info.used_fast_kernel == 1, "a sparse column left the lean kernel"
        # a filter over the gaps: BETWEEN sees only rows that exist
        got = dt.aggregate(tpch.between_filter(1, 0, 99), specs)[0][1]
        assert got[0].value == sum(v for v in range(100) if v % 10) and got[1].value == 90
        words, count = dt.filter_bitmap(tpch.between_filter(1, 0, 99))
        assert count == 90 and not (int(words[0]) & 1) and (int(words[0]) >> 1) & 1
    finally:
        gpu_ctx.set_jit(1)
        dc.destroy()


def test_last_writer_wins_and_nulls_are_not_stored(gpu_ctx):
    """lww_tests.rs::test_lww through append_chunk(row_ids) + llkv_gpu_column_gather."""
    from llkv_b200 import gpu
    dc = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(22, 1), HostColumn(1, DataType.Int64, np.zeros(0, np.int64)))
    try:
        for step in G["lww"]["steps"]:
            ids = np.array([r for r, _ in step["append"]], dtype=np.uint64)
            vals = np.array([0 if v is None else v for _, v in step["append"]], dtype=np.int64)
            valid = pack_validity([v is not None for _, v in step["append"]])
            dc.append_rows(vals, ids, first_row_id=0, validity=valid)
            assert dc.gather(step["gather"]) == step["expect"], step
        assert dc.gather([0, 1, 2, 3, 4, 77]) == [None, 10, 20, 999, None, None]
        assert dc.present_rows() == 3
    finally:
        dc.destroy()


def sparse_lineitem(n_ids: int, seed: int):
    """A lineitem-shaped table after a life of updates and deletes: some row ids are gone, l_discount and l_tax skip rows
    (NULL by absence), some rows were updated in place.  Returns (per-column (row ids, values) chunk lists, the compacted
    HostTable of the rows that exist, MVCC arrays over existing rows)."""
    rng = np.random.default_rng(seed)
    a = tpch.lineitem_arrays(n_ids, seed, with_q1=True)
    exists = rng.random(n_ids) > 0.15
    exists[0] = False
    has_disc = exists & (rng.random(n_ids) > 0.2)
    has_tax = exists & (rng.random(n_ids) > 0.1)
    upd = exists & (rng.random(n_ids) < 0.05)  # rows whose quantity is rewritten later
    new_qty = a["quantity"].copy()
    new_qty[upd] = 100 * rng.integers(1, 51, int(upd.sum()))
    chunks = {}
    ids = np.arange(n_ids, dtype=np.uint64)

    def pieces(mask, vals, piece=7_000):
        sel = np.nonzero(mask)[0]
        return [(ids[sel[i:i + piece]], vals[sel[i:i + piece]]) for i in range(0, sel.size, piece)]

    chunks[tpch.L_QUANTITY] = pieces(exists, a["quantity"]) + pieces(upd, new_qty, 3_000)
    chunks[tpch.L_EXTENDEDPRICE] = pieces(exists, a["extendedprice"])
    chunks[tpch.L_DISCOUNT] = pieces(has_disc, a["discount"])
    chunks[tpch.L_TAX] = pieces(has_tax, a["tax"])
    chunks[tpch.L_SHIPDATE] = pieces(exists, a["shipdate"])
    created, deleted, snap = tpch.mvcc_arrays(n_ids, seed)
    chunks["created"] = pieces(exists, created)
    chunks["deleted"] = pieces(exists, deleted)
    e = np.nonzero(exists)[0]
    t = HostTable(31)
    t.add(HostColumn(tpch.L_QUANTITY, tpch.DEC_15_2, decimal_from_i64(new_qty[e])))
    t.add(HostColumn(tpch.L_EXTENDEDPRICE, tpch.DEC_15_2, decimal_from_i64(a["extendedprice"][e])))
    t.add(HostColumn(tpch.L_DISCOUNT, tpch.DEC_15_2, decimal_from_i64(a["discount"][e]), pack_validity(has_disc[e])))
    t.add(HostColumn(tpch.L_TAX, tpch.DEC_15_2, decimal_from_i64(a["tax"][e]), pack_validity(has_tax[e])))
    t.add(HostColumn(tpch.L_SHIPDATE, DataType.Date32, a["shipdate"][e]))
    t.add_mvcc(created[e], deleted[e])
    return chunks, t, snap, e


def upload_sparse(gpu_ctx, chunks, table_id=31):
    from llkv_b200 import gpu
    dt = gpu.DeviceTable(gpu_ctx, table_id)
    types = {tpch.L_QUANTITY: tpch.DEC_15_2, tpch.L_EXTENDEDPRICE: tpch.DEC_15_2, tpch.L_DISCOUNT: tpch.DEC_15_2, tpch.L_TAX: tpch.DEC_15_2,
             tpch.L_SHIPDATE: DataType.Date32}
    for fid, dtype in types.items():
        dc = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(table_id, fid), HostColumn(fid, dtype, np.zeros((0, 2), np.uint64) if dtype.type == ffi.PT_DECIMAL128 else np.zeros(0, np.int32)))
        for rid, vals in chunks[fid]:
            dc.append_rows(decimal_from_i64(vals) if dtype.type == ffi.PT_DECIMAL128 else vals.astype(np.int32), rid, first_row_id=0)
        dt.columns[fid] = dc
    u64 = HostColumn(0, DataType.UInt64, np.zeros(0, np.uint64))
    dt.created_by = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(table_id, 0xFFFFFFFF, gpu.NS_TXN_CREATED_BY), u64)
    dt.deleted_by = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(table_id, 0xFFFFFFFE, gpu.NS_TXN_DELETED_BY), u64)
    for dc, key in ((dt.created_by, "created"), (dt.deleted_by, "deleted")):
        for rid, vals in chunks[key]:
            dc.append_rows(vals, rid, first_row_id=0)
    dt.seal()
    dt.n_rows = max(c.rows() for c in dt._all())
    return dt


def test_sparse_table_queries_match_the_oracle_over_the_rows_that_exist(gpu_ctx):
    from llkv_b200 import gpu
    chunks, compact, snap, e = sparse_lineitem(120_000, seed=17)
    dt = upload_sparse(gpu_ctx, chunks)
    try:
        # rows: those of the created_by column
        specs = [AggregateSpec("n", AggregateKind.CountStar()), AggregateSpec("cd", AggregateKind.Count(tpch.L_DISCOUNT)),
                 AggregateSpec("nulls", AggregateKind.CountNulls(tpch.L_TAX)), AggregateSpec("sq", AggregateKind.Sum(tpch.L_QUANTITY, tpch.DEC_15_2)),
                 AggregateSpec("ad", AggregateKind.Avg(tpch.L_DISCOUNT, tpch.DEC_15_2)), AggregateSpec("mx", AggregateKind.Max(tpch.L_TAX, tpch.DEC_15_2))]
        for sn in (None, snap):
            want = oracle.aggregate(compact, None, specs, sn)
            util.assert_same_result(dt.aggregate(None, specs, sn), want)
            want = oracle.aggregate(compact, tpch.q6_filter(), tpch.q6_aggregates() + specs, sn)
            util.assert_same_result(dt.aggregate(tpch.q6_filter(), tpch.q6_aggregates() + specs, sn), want)
        # IS NULL / IS NOT NULL leaves
        for op in (Operator.IsNull, Operator.IsNotNull):
            flt = Expr.And([pred(tpch.L_DISCOUNT, op), tpch.q1_filter()])
            want = oracle.aggregate(compact, flt, specs, snap)
            util.assert_same_result(dt.aggregate(flt, specs, snap), want)
        # selection bitmap: positions are row ids; compare through the ids of the selected rows
        words, count = dt.filter_bitmap(tpch.q6_filter(), snap)
        w2, c2 = oracle.filter_bitmap(compact, tpch.q6_filter(), snap)
        assert count == c2
        assert np.array_equal(util.selected_positions(words, dt.n_rows), e[util.selected_positions(w2, compact.n_rows)])
        # both lean builds
        for jit in (1, 2):
            gpu_ctx.set_jit(jit)
            prog = gpu.Program(gpu_ctx, tpch.q6_filter())
            dt.set_snapshot(snap)
            agg = gpu.Aggregation(dt, tpch.q6_aggregates() + specs)
            agg.run(prog, True)
            got = agg.finalize(1)
            info = agg.run_info()
            agg.destroy()
            prog.destroy()
            util.assert_same_result(got, oracle.aggregate(compact, tpch.q6_filter(), tpch.q6_aggregates() + specs, snap))
            assert info.used_fast_kernel == 1 and (jit == 1 or info.used_jit_kernel == 1)  # (mode 1 specialises a shape from its second run on)
    finally:
        gpu_ctx.set_jit(1)
        dt.destroy()


@pytest.mark.parametrize("null_fraction", [0.2, 0.9])
def test_nullable_q6_stays_on_the_lean_kernel(gpu_ctx, null_fraction):
    """one NULL in l_discount used to send Q6 to the general interpreter (8x slower)."""
    from llkv_b200 import gpu
    n = 300_000
    t, _ = tpch.lineitem_table(n, seed=12, with_q1=False)
    rng = np.random.default_rng(3)
    t.columns[tpch.L_DISCOUNT].validity = pack_validity(rng.random(n) >= null_fraction)
    t.columns[tpch.L_EXTENDEDPRICE].validity = pack_validity(rng.random(n) >= 0.01)
    dt = gpu.DeviceTable.from_host(gpu_ctx, t)
    try:
        specs = tpch.q6_aggregates() + [AggregateSpec("n", AggregateKind.CountStar()), AggregateSpec("c", AggregateKind.Count(tpch.L_EXTENDEDPRICE)),
                                        AggregateSpec("mn", AggregateKind.Min(tpch.L_EXTENDEDPRICE, tpch.DEC_15_2))]
        want = oracle.aggregate(t, tpch.q6_filter(), specs)
        for jit in (1, 2):
            gpu_ctx.set_jit(jit)
            prog = gpu.Program(gpu_ctx, tpch.q6_filter())
            agg = gpu.Aggregation(dt, specs)
            agg.run(prog)
            got = agg.finalize(1)
            info = agg.run_info()
            agg.destroy()
            prog.destroy()
            util.assert_same_result(got, want)
            assert info.used_fast_kernel == 1
        for lo, hi in ((0, n), (1_001, 250_007)):
            util.assert_same_result(dt.aggregate(None, specs, row_begin=lo, row_end=hi), oracle.aggregate(t, None, specs, row_begin=lo, row_end=hi))
    finally:
        gpu_ctx.set_jit(1)
        dt.destroy()
