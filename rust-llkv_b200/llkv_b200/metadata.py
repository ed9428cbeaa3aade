"""Host mirror of the column metadata readers (include/llkv_gpu.h "column metadata in front of the scan"): what the Rust
wrapper does between the ColumnCatalog and llkv_gpu_column_append_blob — parse the ColumnDescriptor, walk its page chain,
skip chunks whose statistics cannot match, fetch the chunk blobs."""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Tuple

from .gpu import _check, load

Bound = Tuple[int, int]  # (LLKV_BOUND_INCLUDED 0 / LLKV_BOUND_EXCLUDED 1 / LLKV_BOUND_UNBOUNDED 2, raw bits)
M64 = (1 << 64) - 1


class ChunkMetadata(C.Structure):
    _fields_ = [("chunk_pk", C.c_uint64), ("value_order_perm_pk", C.c_uint64), ("row_count", C.c_uint64), ("serialized_bytes", C.c_uint64),
                ("min_val_u64", C.c_uint64), ("max_val_u64", C.c_uint64), ("null_count", C.c_uint64), ("distinct_count", C.c_uint64)]

    def as_tuple(self):
        return tuple(int(getattr(self, f)) for f, _ in self._fields_)


class ColumnDescriptor(C.Structure):
    _fields_ = [("field_id", C.c_uint64), ("head_page_pk", C.c_uint64), ("tail_page_pk", C.c_uint64), ("total_row_count", C.c_uint64),
                ("total_chunk_count", C.c_uint64), ("data_type_code", C.c_uint32), ("index_meta_len", C.c_uint32)]


class RangeBound(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("value_bits", C.c_uint64)]


def parse_descriptor(blob: bytes) -> ColumnDescriptor:
    out = ColumnDescriptor()
    _check(load().llkv_gpu_descriptor_parse(blob, len(blob), C.byref(out)))
    return out


def parse_descriptor_page(blob: bytes, capacity: int = 256) -> Tuple[int, List[ChunkMetadata]]:
    nxt, n = C.c_uint64(), C.c_uint64()
    arr = (ChunkMetadata * max(1, capacity))()
    _check(load().llkv_gpu_descriptor_page_parse(blob, len(blob), C.byref(nxt), arr, capacity, C.byref(n)))
    return int(nxt.value), [arr[i] for i in range(int(n.value))]


def sortable_u64(prim_type: int, value_bits: int) -> int:
    return int(load().llkv_gpu_sortable_u64(prim_type, value_bits & M64))


def chunk_stats(prim_type: int, values, validity=None) -> ChunkMetadata:
    """compute_chunk_stats for a numpy array (+ optional packed validity bitmap): min/max images, null and distinct counts."""
    import numpy as np
    v = np.ascontiguousarray(values)
    out = ChunkMetadata()
    vb = np.ascontiguousarray(validity, dtype=np.uint8) if validity is not None else None
    _check(load().llkv_gpu_chunk_stats(prim_type, v.ctypes.data if v.size else None, v.size, vb.ctypes.data if vb is not None else None, C.byref(out)))
    return out


def chunk_overlaps(prim_type: int, chunk_min: int, chunk_max: int, lower: Optional[Bound], upper: Optional[Bound]) -> bool:
    lo = RangeBound(lower[0], 0, lower[1] & M64) if lower else None
    hi = RangeBound(upper[0], 0, upper[1] & M64) if upper else None
    return bool(load().llkv_gpu_chunk_overlaps(prim_type, chunk_min, chunk_max, C.byref(lo) if lo else None, C.byref(hi) if hi else None))


def shard_chunks(chunks: List[ChunkMetadata], world: int, rank: int) -> Tuple[List[ChunkMetadata], int]:
    """Multi-GPU loading: rank g of G takes chunks [g C / G, (g + 1) C / G) of every column of the table —
    contiguous, so each shard is a dense row range — and the row id its first row has in the whole table (the
    `row_id_base` of its first append; llkv_gpu_agg_merge then orders groups by first appearance across shards).
    Columns of one table must be cut alike: use the chunk list of one column (rows per chunk differ between types) and
    cut the others at the same row boundaries with `rows_to_chunks`."""
    n = len(chunks)
    lo, hi = n * rank // world, n * (rank + 1) // world
    base = sum(int(c.row_count) for c in chunks[:lo])
    return chunks[lo:hi], base


def rows_to_chunks(chunks: List[ChunkMetadata], row_begin: int, row_end: int) -> Tuple[List[ChunkMetadata], int, int]:
    """The chunks of a column that hold rows [row_begin, row_end), the row id of the first of them and how many leading
    rows of it lie before row_begin (a column whose chunk boundaries differ from the one the shards were cut on)."""
    out, start, first_base, skip = [], 0, 0, 0
    for c in chunks:
        end = start + int(c.row_count)
        if end > row_begin and start < row_end:
            if not out:
                first_base, skip = start, max(0, row_begin - start)
            out.append(c)
        start = end
    return out, first_base, skip


def walk_descriptor(batch_get: Callable[[List[int]], List[bytes]], descriptor_pk: int, prim_type: int = 0,
                    lower: Optional[Bound] = None, upper: Optional[Bound] = None) -> Tuple[ColumnDescriptor, List[ChunkMetadata], int]:
    """The descriptor walk of unsorted_visit (llkv-column-map/src/store/scan/unsorted.rs:202-241): descriptor -> pages
    (one batch_get per page) -> chunk metadata, minus the chunks IntRanges::matches rules out.  Returns the descriptor,
    the surviving chunks in scan order and the number of chunks skipped."""
    desc = parse_descriptor(batch_get([descriptor_pk])[0])
    metas: List[ChunkMetadata] = []
    skipped = 0
    page_pk = desc.head_page_pk
    while page_pk:
        page_pk, entries = parse_descriptor_page(batch_get([page_pk])[0])
        for m in entries:
            if m.row_count == 0:
                continue
            if (lower or upper) and not chunk_overlaps(prim_type, m.min_val_u64, m.max_val_u64, lower, upper):
                skipped += 1
                continue
            metas.append(m)
    return desc, metas, skipped
