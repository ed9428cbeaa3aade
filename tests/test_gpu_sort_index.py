"""GPU: the per-chunk sort index (llkv-column-map/src/store/indexing/sort.rs:150-172 —
`lexsort_to_indices` over a chunk's values, serialised as a UInt32 array blob under value_order_perm_pk) built on the device.
Checked bit-for-bit against a stable host sort in the reference's order: integers by value, floats by IEEE total order
(-NaN < -inf < ... < -0 < +0 < ... < +inf < +NaN), equal values in row order."""
import struct

import numpy as np
import pytest

from llkv_b200.expr import DataType
from llkv_b200.table import HostColumn, decimal_from_i64

pytestmark = pytest.mark.gpu


def total_order_key(v: np.ndarray) -> np.ndarray:
    """u64 whose unsigned order is arrow's ascending order for the type."""
    if v.dtype.kind == "f":
        b = v.astype(np.float64).view(np.uint64)
        neg = (b >> np.uint64(63)).astype(bool)
        return np.where(neg, ~b, b | np.uint64(1 << 63))
    if v.dtype.kind == "i":
        return v.astype(np.int64).view(np.uint64) ^ np.uint64(1 << 63)
    return v.astype(np.uint64)


def parse_perm_blob(blob: bytes) -> np.ndarray:
    assert blob[:4] == b"ARR0" and blob[4] == 0 and blob[5] == 3, blob[:8]  # Primitive layout, PrimType::UInt32
    (n,) = struct.unpack_from("<Q", blob, 8)
    nbytes, zero = struct.unpack_from("<II", blob, 16)
    assert nbytes == 4 * n and zero == 0 and len(blob) == 24 + nbytes
    return np.frombuffer(blob, dtype=np.uint32, offset=24)


CASES = [
    ("i64", DataType.Int64, lambda r, n: r.integers(-(1 << 62), 1 << 62, n, dtype=np.int64)),
    ("i64_dups", DataType.Int64, lambda r, n: r.integers(-5, 5, n, dtype=np.int64)),
    ("i32", DataType.Int32, lambda r, n: r.integers(-(1 << 31), (1 << 31) - 1, n, dtype=np.int32)),
    ("i16", DataType.Int16, lambda r, n: r.integers(-(1 << 15), (1 << 15) - 1, n, dtype=np.int16)),
    ("i8", DataType.Int8, lambda r, n: r.integers(-128, 127, n, dtype=np.int8)),
    ("u64", DataType.UInt64, lambda r, n: r.integers(0, (1 << 64) - 1, n, dtype=np.uint64)),
    ("u32", DataType.UInt32, lambda r, n: r.integers(0, (1 << 32) - 1, n, dtype=np.uint32)),
    ("u16", DataType.UInt16, lambda r, n: r.integers(0, (1 << 16) - 1, n, dtype=np.uint16)),
    ("u8", DataType.UInt8, lambda r, n: r.integers(0, 255, n, dtype=np.uint8)),
    ("date32", DataType.Date32, lambda r, n: r.integers(8000, 11000, n, dtype=np.int32)),
    ("f64", DataType.Float64, lambda r, n: np.where(r.random(n) < 0.02, np.array([np.nan, -np.nan, np.inf, -np.inf, 0.0, -0.0])[r.integers(0, 6, n)],
                                                    r.standard_normal(n) * 1e6)),
    ("f32", DataType.Float32, lambda r, n: (r.standard_normal(n) * 100).astype(np.float32)),
]


@pytest.mark.parametrize("name,dtype,make", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("chunk_rows", [0, 1000])
def test_sort_index_matches_a_stable_host_sort(gpu_ctx, name, dtype, make, chunk_rows):
    from llkv_b200 import gpu
    rng = np.random.default_rng(hash(name) % 1000)
    n = 300_123 if chunk_rows == 0 else 10_500
    v = make(rng, n)
    dc = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(31, 1), HostColumn(1, dtype, v))
    try:
        dc.append_raw(v.ctypes.data, n, 0)
        dc.build_sort_index(chunk_rows)
        width = v.dtype.itemsize
        step = chunk_rows or (1 << 20) // width
        n_chunks = (n + step - 1) // step
        for ci in ([0, n_chunks - 1] if chunk_rows == 0 else range(n_chunks)):
            perm = parse_perm_blob(dc.sort_index_blob(ci))
            chunk = v[ci * step:(ci + 1) * step]
            want = np.argsort(total_order_key(chunk), kind="stable").astype(np.uint32)
            assert np.array_equal(perm, want), (name, ci)
        with pytest.raises(gpu.LlkvError):
            dc.sort_index_blob(n_chunks)
    finally:
        dc.destroy()


def test_sort_index_of_a_narrowed_decimal_column(gpu_ctx):
    """Decimal128 values resident as 4 or 8 bytes sort by their integer image."""
    from llkv_b200 import gpu
    rng = np.random.default_rng(3)
    for scale in (1, 10_000_000):
        v = rng.integers(-2_000_000, 2_000_000, 50_000, dtype=np.int64) * scale
        wide = decimal_from_i64(v)
        dc = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(32, 1), HostColumn(1, DataType.Decimal128(15, 2), wide))
        try:
            dc.append_raw(wide.ctypes.data, len(v), 0)
            dc.build_sort_index(4096)
            for ci in range(0, 13, 4):
                perm = parse_perm_blob(dc.sort_index_blob(ci))
                chunk = v[ci * 4096:(ci + 1) * 4096]
                assert np.array_equal(perm, np.argsort(chunk, kind="stable").astype(np.uint32))
        finally:
            dc.destroy()


def test_sort_index_is_dropped_by_an_append_and_refused_for_nullable_columns(gpu_ctx):
    from llkv_b200 import gpu
    v = np.arange(1000, dtype=np.int64)[::-1].copy()
    dc = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(33, 1), HostColumn(1, DataType.Int64, v))
    try:
        dc.append_raw(v.ctypes.data, 1000, 0)
        dc.build_sort_index(0)
        assert np.array_equal(parse_perm_blob(dc.sort_index_blob(0)), np.arange(1000, dtype=np.uint32)[::-1])
        dc.append_raw(v.ctypes.data, 1000, 1000)
        dc.seal()
        with pytest.raises(gpu.LlkvError):
            dc.sort_index_blob(0)  # stale: the column changed since the build
        dc.build_sort_index(0)
        assert len(parse_perm_blob(dc.sort_index_blob(0))) == 2000
    finally:
        dc.destroy()
    valid = np.ones(1000, dtype=bool)
    valid[5] = False
    hc = HostColumn(2, DataType.Int64, v, np.packbits(valid, bitorder="little"))
    dc = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(33, 2), hc)
    try:
        dc.append(hc)
        with pytest.raises(gpu.LlkvError):
            dc.build_sort_index(0)
    finally:
        dc.destroy()
