"""GPU: ColumnStore::scan with ScanOptions (llkv-column-map/src/store/scan/options.rs:13-37, scan/mod.rs:191-1080) through
llkv_gpu_column_scan: the reference's own pagination tests transcribed (llkv-column-map/tests/pagination_tests.rs:26-312,
null_pagination_tests.rs:17-129), then seeded cases against numpy — every integer width, floats in total order with NaNs
and signed zeros, value ranges, reverse scans, null runs before / after the values."""
import numpy as np
import pytest

from llkv_b200.expr import DataType
from llkv_b200.table import HostColumn, decimal_from_i64, pack_validity
from test_gpu_sort_index import CASES, total_order_key

pytestmark = pytest.mark.gpu


def column(gpu_ctx, table_id, field, dtype, values, validity=None, row_ids=None):
    from llkv_b200 import gpu
    hc = HostColumn(field, dtype, values, pack_validity(validity) if validity is not None else None)
    dc = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(table_id, field), hc)
    if row_ids is not None:
        dc.append_rows(hc.values, row_ids, 0)
    else:
        dc.append(hc)
    return dc


def collect(dc, **kw):
    vals, ids = [], []

    def on_run(v, r):
        vals.append(v)
        ids.append(r)
    dc.scan(on_run, **kw)
    return vals, ids


def flat(chunks):
    chunks = [c for c in chunks if c is not None]
    return np.concatenate(chunks) if chunks else np.zeros(0)


def test_reference_pagination_unsorted_u64(gpu_ctx):
    """pagination_tests.rs:26-157: 1000 rows of 2*i in append order; limit 25; offset 975; window (100, 37)."""
    vals = np.arange(1000, dtype=np.uint64) * 2
    dc = column(gpu_ctx, 81, 1, DataType.UInt64, vals)
    try:
        assert np.array_equal(flat(collect(dc, chunk_rows=256)[0]), vals)
        assert np.array_equal(flat(collect(dc, limit=25, chunk_rows=256)[0]), vals[:25])
        assert np.array_equal(flat(collect(dc, offset=975, chunk_rows=256)[0]), vals[975:])
        assert np.array_equal(flat(collect(dc, offset=100, limit=37, chunk_rows=16)[0]), vals[100:137])
    finally:
        dc.destroy()


def test_reference_pagination_sorted_u64(gpu_ctx):
    """pagination_tests.rs:160-312: 2048 descending values; ascending and descending scans, windows (123, 77) and limit 50."""
    vals = np.arange(2048, dtype=np.uint64)[::-1].copy()
    dc = column(gpu_ctx, 81, 2, DataType.UInt64, vals)
    try:
        asc_all = flat(collect(dc, sorted=True, chunk_rows=300)[0])
        assert len(asc_all) == 2048 and np.all(asc_all[:-1] <= asc_all[1:]) and np.array_equal(asc_all, np.arange(2048, dtype=np.uint64))
        assert np.array_equal(flat(collect(dc, sorted=True, offset=123, limit=77, chunk_rows=50)[0]), asc_all[123:200])
        desc_all = flat(collect(dc, sorted=True, reverse=True, chunk_rows=300)[0])
        assert len(desc_all) == 2048 and np.all(desc_all[:-1] >= desc_all[1:])
        assert np.array_equal(flat(collect(dc, sorted=True, reverse=True, limit=50)[0]), desc_all[:50])
    finally:
        dc.destroy()


def test_reference_sorted_with_nulls_last_pagination(gpu_ctx):
    """null_pagination_tests.rs:17-129: anchor rows 0..100, the target column holds the even ones (value = 10 * row id);
    sorted, with row ids, nulls last, offset 10, limit 15 -> row ids 20, 22, ..., 48 and 15 values."""
    from llkv_b200 import gpu
    anchor = column(gpu_ctx, 82, 1, DataType.UInt64, np.arange(100, dtype=np.uint64))
    rids = np.arange(0, 100, 2, dtype=np.uint64)
    target = column(gpu_ctx, 82, 2, DataType.UInt64, rids * 10, row_ids=rids)
    try:
        vals, ids = collect(target, sorted=True, with_row_ids=True, limit=15, offset=10, include_nulls=True, nulls_first=False, anchor=anchor)
        assert np.array_equal(flat(ids), np.arange(20, 50, 2, dtype=np.uint64))
        assert len(flat(vals)) == 15 and np.array_equal(flat(vals), np.arange(20, 50, 2, dtype=np.uint64) * 10)
        # the page that crosses from the values into the nulls: 5 values (rows 90..98), then the odd rows 1, 3, ... as null runs
        vals, ids = collect(target, sorted=True, with_row_ids=True, limit=12, offset=45, include_nulls=True, nulls_first=False, anchor=anchor)
        assert np.array_equal(flat(vals), np.arange(90, 100, 2, dtype=np.uint64) * 10)
        assert np.array_equal(flat(ids), np.concatenate([np.arange(90, 100, 2), np.arange(1, 15, 2)]).astype(np.uint64))
        assert [v is None for v in vals] == [False, True]
        # nulls first, reverse: null row ids descending, then values descending
        vals, ids = collect(target, sorted=True, reverse=True, with_row_ids=True, limit=53, offset=0, include_nulls=True, nulls_first=True, anchor=anchor)
        assert np.array_equal(flat(ids), np.concatenate([np.arange(99, 0, -2), [98, 96, 94]]).astype(np.uint64))
        assert np.array_equal(flat(vals), np.array([980, 960, 940], dtype=np.uint64))
    finally:
        anchor.destroy()
        target.destroy()


@pytest.mark.parametrize("name,dtype,make", CASES, ids=[c[0] for c in CASES])
def test_sorted_scans_match_a_stable_host_sort(gpu_ctx, name, dtype, make):
    rng = np.random.default_rng(len(name) * 7 + 1)
    n = 200_003
    v = make(rng, n)
    valid = rng.random(n) > 0.1
    dc = column(gpu_ctx, 83, 1, dtype, v, valid)
    try:
        held = np.nonzero(valid)[0]
        order = held[np.argsort(total_order_key(v[held]), kind="stable")]
        vals, ids = collect(dc, sorted=True, with_row_ids=True, chunk_rows=65_536)
        assert np.array_equal(flat(ids), order.astype(np.uint64))
        assert np.array_equal(flat(vals).view(np.uint8), v[order].view(np.uint8))
        vals, ids = collect(dc, sorted=True, reverse=True, with_row_ids=True, offset=1000, limit=5000)
        assert np.array_equal(flat(ids), order[::-1][1000:6000].astype(np.uint64))
        # a value range: bounds in the column's type
        keys = total_order_key(v[order])
        lo_v, hi_v = v[order][len(order) // 4], v[order][3 * len(order) // 4]
        vals, ids = collect(dc, sorted=True, with_row_ids=True, lower=(lo_v, True), upper=(hi_v, False))
        lo_k, hi_k = total_order_key(np.array([lo_v]))[0], total_order_key(np.array([hi_v]))[0]
        want = order[(keys >= lo_k) & (keys < hi_k)]
        assert np.array_equal(flat(ids), want.astype(np.uint64))
    finally:
        dc.destroy()


def test_sorted_scan_of_a_narrowed_decimal_column(gpu_ctx):
    rng = np.random.default_rng(5)
    v = rng.integers(-2_000_000, 2_000_000, 70_000, dtype=np.int64)
    dc = column(gpu_ctx, 84, 1, DataType.Decimal128(15, 2), decimal_from_i64(v))
    try:
        vals, ids = collect(dc, sorted=True, with_row_ids=True, limit=1000)
        order = np.argsort(v, kind="stable")[:1000]
        assert np.array_equal(flat(ids), order.astype(np.uint64))
        assert np.array_equal(np.concatenate(vals).reshape(-1, 2), decimal_from_i64(v[order]).reshape(-1, 2))
    finally:
        dc.destroy()


@pytest.mark.parametrize("dtype,np_t,values,low,high", [
    (DataType.Float64, np.float64, [12.5, -3.2, 45.6, 0.0, 18.75, 99.9, -42.1, 5.5, 18.75, 64.0, -0.01, 23.4], -10.0, 50.0),
    (DataType.Float32, np.float32, [1.25, -8.5, 12.0, 3.5, 6.75, 42.125, -16.0, 3.5, 9.0, 15.5, 27.25, -0.5], -5.0, 20.0),
], ids=["float64", "float32"])
def test_reference_float_sorted_scan_and_ranges(gpu_ctx, dtype, np_t, values, low, high):
    """float_scan_tests.rs:101-275 (`float64_sorted_scan_and_ranges`, `float32_sorted_scan_and_ranges`): the fixture values,
    an ascending scan, and the inclusive range low..=high with values and with row ids."""
    v = np.array(values, dtype=np_t)
    dc = column(gpu_ctx, 85, 1, dtype, v)
    try:
        asc = flat(collect(dc, sorted=True)[0])
        assert np.all(asc[:-1] <= asc[1:]) and len(asc) == len(v)
        pairs = sorted(((x, i) for i, x in enumerate(v)), key=lambda p: (p[0], p[1]))
        want = [(x, i) for x, i in pairs if low <= x <= high]
        vals, ids = collect(dc, sorted=True, with_row_ids=True, lower=(np_t(low), True), upper=(np_t(high), True))
        assert list(flat(vals)) == [x for x, _ in want] and list(flat(ids)) == [i for _, i in want]
    finally:
        dc.destroy()
