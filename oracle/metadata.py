"""TEST INFRASTRUCTURE ONLY (see oracle/oracle.py): CPU restatement of the column metadata the reference's scan walks
before it reads a chunk — descriptor pages, per-chunk statistics, order-preserving value images and the chunk-pruning
rule.  Pinned by the reference's own tests (llkv-column-map/tests/pruning_tests.rs:16-126, transcribed in
tests/test_metadata.py).  Nothing under rust-llkv_b200/ imports this module."""
import struct
from typing import List, Optional, Tuple

import numpy as np

SIGN64 = 1 << 63
M64 = (1 << 64) - 1
# PrimType codes (llkv-column-map/src/serialization.rs:146-166)
UINT64, INT32, UINT32, FLOAT32, INT64, INT16, INT8, UINT16, UINT8, FLOAT64, DATE32, DATE64 = 1, 2, 3, 4, 6, 7, 8, 9, 10, 11, 16, 17
INCLUDED, EXCLUDED, UNBOUNDED = 0, 1, 2  # include/llkv_gpu.h LLKV_BOUND_*


def sortable_u64(prim_type: int, value) -> int:
    """llkv-column-map/src/codecs.rs:33-65 (signed: flip the sign bit of the type's own width; floats: sign-flip trick,
    f32 through f64) and pruning.rs:123-170 (unsigned: the value itself)."""
    if prim_type == INT8:
        return (int(value) & 0xFF) ^ 0x80
    if prim_type == INT16:
        return (int(value) & 0xFFFF) ^ 0x8000
    if prim_type in (INT32, DATE32):
        return (int(value) & 0xFFFFFFFF) ^ 0x80000000
    if prim_type in (INT64, DATE64):
        return (int(value) & M64) ^ SIGN64
    if prim_type in (FLOAT32, FLOAT64):
        v = float(np.float32(value)) if prim_type == FLOAT32 else float(value)
        bits = struct.unpack("<Q", struct.pack("<d", v))[0]
        return (~bits & M64) if bits & SIGN64 else (bits | SIGN64)
    return int(value) & M64


def check_overlap(lower: Tuple[int, Optional[int]], upper: Tuple[int, Optional[int]], chunk_min: int, chunk_max: int) -> bool:
    """pruning.rs:207-247: range (lower, upper) against the inclusive chunk interval, all in the sortable u64 domain."""
    kind, u = upper
    if kind == INCLUDED and u < chunk_min:
        return False
    if kind == EXCLUDED and u <= chunk_min:
        return False
    kind, lo = lower
    if kind == INCLUDED and lo > chunk_max:
        return False
    if kind == EXCLUDED and lo >= chunk_max:
        return False
    return True


def chunk_matches(prim_type: int, lower, upper, chunk_min: int, chunk_max: int) -> bool:
    """IntRanges::matches with the range of one type set (pruning.rs:104-205): bounds are mapped with the type's codec."""
    lo = (lower[0], None if lower[0] == UNBOUNDED else sortable_u64(prim_type, lower[1]))
    hi = (upper[0], None if upper[0] == UNBOUNDED else sortable_u64(prim_type, upper[1]))
    return check_overlap(lo, hi, chunk_min, chunk_max)


def chunk_stats(prim_type: int, values: np.ndarray, valid: Optional[np.ndarray] = None):
    """compute_chunk_stats (pruning.rs:272-470) for primitive arrays: (min_u64, max_u64, null_count, distinct_count), None
    for an empty array, zeros when every value is NULL."""
    n = len(values)
    if n == 0:
        return None
    valid = np.ones(n, dtype=bool) if valid is None else np.asarray(valid, dtype=bool)
    nulls = int(n - valid.sum())
    if nulls == n:
        return (0, 0, nulls, 0)
    v = np.asarray(values)[valid]
    if prim_type in (FLOAT32, FLOAT64):
        # strict < / > from +inf / -inf: NaN never becomes a bound; distinct counts bit patterns (f32: to_bits of the f32)
        finite = v[~np.isnan(v)]
        mn = float(finite.min()) if finite.size else float("inf")
        mx = float(finite.max()) if finite.size else float("-inf")
        distinct = len(np.unique(v.view(np.uint32 if prim_type == FLOAT32 else np.uint64)))
        return (sortable_u64(prim_type, mn), sortable_u64(prim_type, mx), nulls, distinct)
    return (sortable_u64(prim_type, v.min()), sortable_u64(prim_type, v.max()), nulls, len(np.unique(v)))


# ---- descriptor blobs (llkv-column-map/src/store/descriptor.rs) -------------------------------------------------------
def chunk_metadata_bytes(m: Tuple[int, ...]) -> bytes:
    """ChunkMetadata::to_le_bytes (:37-49): chunk_pk, value_order_perm_pk, row_count, serialized_bytes, min, max, null_count,
    distinct_count."""
    return struct.pack("<8Q", *m)


def descriptor_bytes(field_id: int, head: int, tail: int, rows: int, chunks: int, data_type_code: int = 0, index_meta: bytes = b"") -> bytes:
    """ColumnDescriptor::to_le_bytes (:245-261)."""
    return struct.pack("<5Q3I", field_id, head, tail, rows, chunks, data_type_code, 0, len(index_meta)) + index_meta


def descriptor_pages(metas: List[Tuple[int, ...]], page_pks: List[int], per_page: int = 63) -> List[Tuple[int, bytes]]:
    """The page chain the append path builds: a page takes another entry while it stays within 4096 bytes and 256 entries
    (store/core.rs:2179-2180), i.e. 63 entries behind the 16-byte header."""
    pages = []
    groups = [metas[i:i + per_page] for i in range(0, len(metas), per_page)] or [[]]
    assert len(page_pks) >= len(groups)
    for i, g in enumerate(groups):
        nxt = page_pks[i + 1] if i + 1 < len(groups) else 0
        pages.append((page_pks[i], struct.pack("<QI4x", nxt, len(g)) + b"".join(chunk_metadata_bytes(m) for m in g)))
    return pages


def parse_descriptor(b: bytes):
    """ColumnDescriptor::from_le_bytes (:263-299)."""
    field_id, head, tail, rows, chunks = struct.unpack_from("<5Q", b, 0)
    code = iml = 0
    if len(b) >= 52:
        code, _pad, iml = struct.unpack_from("<3I", b, 40)
        if not (iml > 0 and len(b) >= 52 + iml):
            iml = 0
    return {"field_id": field_id, "head_page_pk": head, "tail_page_pk": tail, "total_row_count": rows, "total_chunk_count": chunks,
            "data_type_code": code, "index_meta_len": iml}


def parse_page(b: bytes):
    """DescriptorPageHeader::from_le_bytes + entries (:364-378, :419-434) -> (next_page_pk, [ChunkMetadata tuples])."""
    nxt, n = struct.unpack_from("<QI", b, 0)
    return nxt, [struct.unpack_from("<8Q", b, 16 + 64 * i) for i in range(n)]
