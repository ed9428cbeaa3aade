"""CPU: the column metadata readers of the C ABI (descriptor pages, sortable value images, chunk pruning) against the
oracle restatement (oracle/metadata.py) and the reference's own known answers (llkv-column-map/tests/pruning_tests.rs)."""
import struct

import numpy as np
import pytest

from llkv_b200 import ffi, metadata
from llkv_b200.table import LlkvError
from oracle import metadata as om

INC, EXC, UNB = 0, 1, 2


def f64_bits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def f32_bits(x: float) -> int:
    return struct.unpack("<I", struct.pack("<f", x))[0]


# ---- llkv-column-map/tests/pruning_tests.rs:16-64 (compute_chunk_stats) pins the oracle
def test_reference_chunk_stats_known_answers():
    st = om.chunk_stats(om.INT32, np.array([1, 5, 10, 0], dtype=np.int32), np.array([1, 1, 1, 0], dtype=bool))
    assert st == ((1 ^ 0x80000000), (10 ^ 0x80000000), 1, 3)
    st = om.chunk_stats(om.FLOAT64, np.array([1.0, -5.0, 10.5, 0.0]), np.array([1, 1, 1, 0], dtype=bool))
    assert st[0] < st[1] and st[2:] == (1, 3)
    assert st[0] == om.sortable_u64(om.FLOAT64, -5.0) and st[1] == om.sortable_u64(om.FLOAT64, 10.5)
    assert om.chunk_stats(om.INT32, np.array([0, 0], dtype=np.int32), np.array([0, 0], dtype=bool)) == (0, 0, 2, 0)
    assert om.chunk_stats(om.INT32, np.array([], dtype=np.int32)) is None


# ---- pruning_tests.rs:66-126 (IntRanges::matches) pins the oracle and the C ABI alike
def test_reference_pruning_known_answers():
    enc = lambda v: (v & 0xFFFFFFFF) ^ 0x80000000  # noqa: E731  (the test's own encoding of i32 chunk bounds)
    for impl in (lambda t, lo, hi, a, b: om.chunk_matches(t, lo, hi, a, b),
                 lambda t, lo, hi, a, b: metadata.chunk_overlaps(t, a, b, lo, hi)):
        rng = ((INC, 5), (INC, 10))
        assert not impl(om.INT32, *rng, enc(1), enc(4))
        assert impl(om.INT32, *rng, enc(1), enc(6))
        assert not impl(om.INT32, *rng, enc(11), enc(20))
    f = lambda x: om.sortable_u64(om.FLOAT64, x)  # noqa: E731  (== the test's f64_to_u64 helper)
    assert f(-2.0) == (~f64_bits(-2.0) & om.M64) and f(0.5) == (f64_bits(0.5) | om.SIGN64)
    lo, hi = (INC, f64_bits(-1.0)), (INC, f64_bits(1.0))
    assert not metadata.chunk_overlaps(om.FLOAT64, f(-2.0), f(-1.5), lo, hi)
    assert metadata.chunk_overlaps(om.FLOAT64, f(-0.5), f(0.5), lo, hi)
    assert not om.chunk_matches(om.FLOAT64, (INC, -1.0), (INC, 1.0), f(-2.0), f(-1.5))
    assert om.chunk_matches(om.FLOAT64, (INC, -1.0), (INC, 1.0), f(-0.5), f(0.5))


@pytest.mark.parametrize("prim_type,np_type,width", [(om.INT8, np.int8, 1), (om.INT16, np.int16, 2), (om.INT32, np.int32, 4), (om.INT64, np.int64, 8),
                                                     (om.DATE32, np.int32, 4), (om.DATE64, np.int64, 8), (om.UINT8, np.uint8, 1),
                                                     (om.UINT16, np.uint16, 2), (om.UINT32, np.uint32, 4), (om.UINT64, np.uint64, 8)])
def test_sortable_images_preserve_order_and_match_the_oracle(prim_type, np_type, width):
    rng = np.random.default_rng(prim_type)
    info = np.iinfo(np_type)
    vals = np.concatenate([rng.integers(info.min, info.max, 300, dtype=np_type, endpoint=True), np.array([info.min, info.max, 0, 1], dtype=np_type)])
    if info.min < 0:
        vals = np.concatenate([vals, np.array([-1], dtype=np_type)])
    vals = np.unique(vals)
    images = [metadata.sortable_u64(prim_type, int(v) & ((1 << (8 * width)) - 1)) for v in vals]
    assert images == [om.sortable_u64(prim_type, int(v)) for v in vals]
    assert images == sorted(images) and len(set(images)) == len(images)
    # sign-extended raw bits (how a register holds an i8/i16/i32) give the same image
    assert images == [metadata.sortable_u64(prim_type, int(v)) for v in vals]


def test_float_images():
    vals = [float("-inf"), -1e300, -2.5, -0.0, 0.0, 1e-300, 3.25, 1e300, float("inf")]
    img = [metadata.sortable_u64(om.FLOAT64, f64_bits(v)) for v in vals]
    assert img == [om.sortable_u64(om.FLOAT64, v) for v in vals] and img == sorted(img)
    v32 = [-3.5, -0.0, 0.0, 1.25, 7e37]
    img32 = [metadata.sortable_u64(om.FLOAT32, f32_bits(v)) for v in v32]
    assert img32 == [om.sortable_u64(om.FLOAT32, v) for v in v32] and img32 == sorted(img32)
    assert img32[3] == om.sortable_u64(om.FLOAT64, 1.25)  # Float32 goes through f64 (codecs.rs:44-46)


def test_chunk_overlaps_matches_the_oracle_on_every_bound_kind():
    rng = np.random.default_rng(7)
    for prim_type, lo_v, hi_v in ((om.INT64, -50, 50), (om.UINT32, 0, 100), (om.DATE32, -20, 20), (om.INT8, -100, 100)):
        for _ in range(400):
            a, b = sorted(int(x) for x in rng.integers(lo_v, hi_v, 2))
            cmin, cmax = om.sortable_u64(prim_type, a), om.sortable_u64(prim_type, b)
            lower = (int(rng.integers(0, 3)), int(rng.integers(lo_v, hi_v)))
            upper = (int(rng.integers(0, 3)), int(rng.integers(lo_v, hi_v)))
            want = om.chunk_matches(prim_type, lower, upper, cmin, cmax)
            got = metadata.chunk_overlaps(prim_type, cmin, cmax, lower, upper)
            assert got == want, (prim_type, a, b, lower, upper)
            # the rule never drops a chunk that holds a matching value
            inside = [v for v in range(a, b + 1)
                      if (lower[0] == UNB or (v >= lower[1] if lower[0] == INC else v > lower[1]))
                      and (upper[0] == UNB or (v <= upper[1] if upper[0] == INC else v < upper[1]))]
            if inside:
                assert got
    assert metadata.chunk_overlaps(om.INT64, 5, 9, None, None)  # no range: everything matches


def test_descriptor_and_page_chain_round_trip():
    metas = [(1000 + i, 0 if i % 3 else 7000 + i, 4096 if i < 149 else 100, 65560, om.sortable_u64(om.INT64, i * 10), om.sortable_u64(om.INT64, i * 10 + 9),
              i % 2, 10) for i in range(150)]
    pages = om.descriptor_pages(metas, [500, 501, 502])
    assert [len(b) for _, b in pages] == [16 + 63 * 64, 16 + 63 * 64, 16 + 24 * 64] and all(len(b) <= 4096 for _, b in pages)
    desc_blob = om.descriptor_bytes(field_id=(1 << 48) | 42, head=500, tail=502, rows=149 * 4096 + 100, chunks=150, data_type_code=6,
                                    index_meta=struct.pack("<IB", 1, 2))
    pager = {499: desc_blob, **dict(pages)}
    d = metadata.parse_descriptor(desc_blob)
    assert {f: int(getattr(d, f)) for f, _ in d._fields_} == om.parse_descriptor(desc_blob)
    assert d.index_meta_len == 5 and d.data_type_code == 6
    old = metadata.parse_descriptor(desc_blob[:40])  # files written before the type code existed
    assert (old.data_type_code, old.index_meta_len, old.total_chunk_count) == (0, 0, 150)
    nxt, entries = metadata.parse_descriptor_page(pages[0][1])
    assert nxt == 501 and [e.as_tuple() for e in entries] == list(om.parse_page(pages[0][1])[1]) == metas[:63]
    desc, chunks, skipped = metadata.walk_descriptor(lambda pks: [pager[k] for k in pks], 499)
    assert [c.as_tuple() for c in chunks] == metas and skipped == 0 and desc.total_row_count == 149 * 4096 + 100
    # range [205, 398]: chunks 20..39 survive (chunk i covers [10 i, 10 i + 9])
    _, chunks, skipped = metadata.walk_descriptor(lambda pks: [pager[k] for k in pks], 499, om.INT64, (INC, 205), (EXC, 399))
    assert [c.chunk_pk for c in chunks] == [1000 + i for i in range(20, 40)] and skipped == 130
    _, chunks, _ = metadata.walk_descriptor(lambda pks: [pager[k] for k in pks], 499, om.INT64, (EXC, 209), (UNB, 0))
    assert chunks[0].chunk_pk == 1021  # Excluded(209) >= chunk 20's maximum (209): skipped (pruning.rs:236-240)


def test_malformed_metadata_is_an_error_not_a_crash():
    page = om.descriptor_pages([(1, 0, 1, 1, 0, 0, 0, 0)] * 3, [9])[0][1]
    with pytest.raises(LlkvError):
        metadata.parse_descriptor_page(page[:-1])        # truncated entry
    with pytest.raises(LlkvError):
        metadata.parse_descriptor_page(page[:10])        # truncated header
    with pytest.raises(LlkvError):
        metadata.parse_descriptor_page(page, capacity=2)  # caller's buffer too small
    with pytest.raises(LlkvError):
        metadata.parse_descriptor(b"\x00" * 39)
    empty = struct.pack("<QI4x", 0, 0)
    assert metadata.parse_descriptor_page(empty) == (0, [])


# ---- compute_chunk_stats through the C ABI: the reference's known answers (pruning_tests.rs:16-64) and the oracle
def test_chunk_stats_known_answers_through_the_abi():
    from llkv_b200.table import pack_validity
    st = metadata.chunk_stats(om.INT32, np.array([1, 5, 10, 0], dtype=np.int32), pack_validity([1, 1, 1, 0]))
    assert (st.min_val_u64, st.max_val_u64, st.null_count, st.distinct_count) == ((1 ^ 0x80000000), (10 ^ 0x80000000), 1, 3)
    st = metadata.chunk_stats(om.FLOAT64, np.array([1.0, -5.0, 10.5, 0.0]), pack_validity([1, 1, 1, 0]))
    assert st.min_val_u64 < st.max_val_u64 and (st.null_count, st.distinct_count) == (1, 3)
    st = metadata.chunk_stats(om.INT32, np.array([0, 0], dtype=np.int32), pack_validity([0, 0]))
    assert (st.min_val_u64, st.max_val_u64, st.null_count, st.distinct_count) == (0, 0, 2, 0)
    with pytest.raises(LlkvError) as e:
        metadata.chunk_stats(om.INT32, np.array([], dtype=np.int32))
    assert e.value.code == 4  # NotFound: the reference returns None for an empty array


@pytest.mark.parametrize("prim_type,np_type", [(om.INT8, np.int8), (om.INT16, np.int16), (om.INT32, np.int32), (om.INT64, np.int64), (om.DATE32, np.int32),
                                               (om.UINT8, np.uint8), (om.UINT16, np.uint16), (om.UINT32, np.uint32), (om.UINT64, np.uint64),
                                               (om.FLOAT32, np.float32), (om.FLOAT64, np.float64)])
def test_chunk_stats_match_the_oracle(prim_type, np_type):
    from llkv_b200.table import pack_validity
    rng = np.random.default_rng(100 + prim_type)
    for n in (1, 7, 64, 1000):
        if np.issubdtype(np_type, np.floating):
            v = (rng.standard_normal(n) * 1e3).astype(np_type)
            v[rng.random(n) < 0.1] = np.nan
            v[rng.random(n) < 0.1] = -0.0
        else:
            info = np.iinfo(np_type)
            v = rng.integers(max(info.min, -50), min(info.max, 50), n, dtype=np_type, endpoint=True)
            if n > 2:
                v[0], v[1] = info.min, info.max
        valid = rng.random(n) < 0.8
        for vb in (None, valid):
            st = metadata.chunk_stats(prim_type, v, pack_validity(vb) if vb is not None else None)
            want = om.chunk_stats(prim_type, v, vb)
            assert (st.min_val_u64, st.max_val_u64, st.null_count, st.distinct_count) == want, (prim_type, n, vb is not None)
    # the statistics feed the pruning rule: a range that holds no value of the chunk is reported disjoint or not, never wrongly
    v = np.array([3, 9, 4], dtype=np_type)
    st = metadata.chunk_stats(prim_type, v)
    bits = lambda x: int(np.array([x], dtype=np_type).view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[np.dtype(np_type).itemsize])[0])  # noqa: E731
    assert metadata.chunk_overlaps(prim_type, st.min_val_u64, st.max_val_u64, (0, bits(4)), (0, bits(5)))
    assert not metadata.chunk_overlaps(prim_type, st.min_val_u64, st.max_val_u64, (0, bits(10)), (2, 0))
    assert not metadata.chunk_overlaps(prim_type, st.min_val_u64, st.max_val_u64, (2, 0), (1, bits(3)))


def _meta(pk, rows):
    m = metadata.ChunkMetadata()
    m.chunk_pk, m.row_count = pk, rows
    return m


def test_chunk_shards_cover_every_row_once():
    # an Int64 column (131 072 rows per chunk) and a Decimal128 column (4 096 rows per chunk) of one 1 000 000-row table
    n = 1_000_000
    wide = [_meta(10 + i, min(131072, n - i * 131072)) for i in range((n + 131071) // 131072)]
    narrow = [_meta(1000 + i, min(4096, n - i * 4096)) for i in range((n + 4095) // 4096)]
    for world in (1, 2, 3, 8):
        covered = 0
        for rank in range(world):
            mine, base = metadata.shard_chunks(wide, world, rank)
            assert base == covered
            rows = sum(int(c.row_count) for c in mine)
            # the Decimal128 column is cut at the same rows: whole chunks here, since 131 072 is a multiple of 4 096
            other, first_base, skip = metadata.rows_to_chunks(narrow, base, base + rows)
            assert first_base == base and skip == 0 and sum(int(c.row_count) for c in other) == rows
            covered += rows
        assert covered == n
    # boundaries that do not line up: the first chunk starts before the shard
    other, first_base, skip = metadata.rows_to_chunks(narrow, 5000, 9000)
    assert [c.chunk_pk for c in other] == [1001, 1002] and first_base == 4096 and skip == 904
