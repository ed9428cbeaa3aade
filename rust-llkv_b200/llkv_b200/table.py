"""Host-side column containers: Arrow-layout buffers as numpy arrays (no semantics, just bytes).

A `HostColumn` is what `Pager::batch_get` + `deserialize_array` hand to a visitor in the reference
(llkv-column-map/src/serialization.rs:438-488): a values buffer, optionally a validity bitmap, and for Utf8 the
offsets + data buffers.  Row ids are dense (0..n) as in every table the reference builds by plain appends.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import ffi
from .expr import DataType

_NP = {
    ffi.PT_UINT64: np.uint64, ffi.PT_INT32: np.int32, ffi.PT_UINT32: np.uint32, ffi.PT_FLOAT32: np.float32,
    ffi.PT_INT64: np.int64, ffi.PT_INT16: np.int16, ffi.PT_INT8: np.int8, ffi.PT_UINT16: np.uint16,
    ffi.PT_UINT8: np.uint8, ffi.PT_FLOAT64: np.float64, ffi.PT_BOOLEAN: np.uint8, ffi.PT_DATE32: np.int32,
    ffi.PT_DATE64: np.int64,
}


class LlkvError(Exception):
    """llkv_result::Error (llkv-result/src/error.rs:31-176) as surfaced through the C ABI."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{ffi.ERROR_NAMES.get(code, code)}: {message}")
        self.code = code
        self.message = message


def decimal_array(values: Iterable[int]) -> np.ndarray:
    """Python ints -> Arrow Decimal128 values buffer as an (n, 2) uint64 array of (lo, hi) words."""
    vals = list(values)
    out = np.empty((len(vals), 2), dtype=np.uint64)
    for i, v in enumerate(vals):
        lo, hi = ffi.i128_to_words(int(v))
        out[i, 0] = lo
        out[i, 1] = hi
    return out


def decimal_from_i64(values: np.ndarray) -> np.ndarray:
    """Vectorised sign extension of an int64 array into Decimal128 (lo, hi) words."""
    v = np.ascontiguousarray(values, dtype=np.int64)
    out = np.empty((v.shape[0], 2), dtype=np.uint64)
    out[:, 0] = v.view(np.uint64)
    out[:, 1] = (v >> 63).view(np.uint64)
    return out


def decimal_to_ints(buf: np.ndarray) -> List[int]:
    return [ffi.words_to_i128(int(lo), int(hi)) for lo, hi in buf]


def pack_validity(valid: Sequence[bool]) -> np.ndarray:
    """bool per row -> Arrow validity bitmap (LSB first)."""
    return np.packbits(np.asarray(valid, dtype=np.uint8), bitorder="little")


@dataclass
class HostColumn:
    field_id: int
    dtype: DataType
    values: np.ndarray               # Utf8: int32 offsets (n+1)
    validity: Optional[np.ndarray] = None  # packed bitmap, uint8
    aux: Optional[np.ndarray] = None       # Utf8: data bytes (uint8)

    def __post_init__(self):
        t = self.dtype.type
        if t == ffi.PT_DECIMAL128:
            self.values = np.ascontiguousarray(self.values, dtype=np.uint64).reshape(-1, 2)
        elif t == ffi.PT_UTF8:
            self.values = np.ascontiguousarray(self.values, dtype=np.int32)
            self.aux = np.ascontiguousarray(self.aux if self.aux is not None else np.zeros(0, np.uint8), dtype=np.uint8)
        else:
            self.values = np.ascontiguousarray(self.values, dtype=_NP[t])
        if self.validity is not None:
            self.validity = np.ascontiguousarray(self.validity, dtype=np.uint8)
            assert self.validity.shape[0] * 8 >= self.n_rows

    @property
    def n_rows(self) -> int:
        if self.dtype.type == ffi.PT_UTF8:
            return max(0, self.values.shape[0] - 1)
        return self.values.shape[0]

    @staticmethod
    def utf8(field_id: int, strings: Sequence[Optional[str]]) -> "HostColumn":
        offs = [0]
        data = bytearray()
        valid = []
        for s in strings:
            if s is not None:
                data += s.encode("utf-8")
            valid.append(s is not None)
            offs.append(len(data))
        col = HostColumn(field_id, DataType.Utf8, np.asarray(offs, np.int32), aux=np.frombuffer(bytes(data), np.uint8).copy())
        if not all(valid):
            col.validity = pack_validity(valid)
        return col

    def string_at(self, row: int) -> str:
        return bytes(self.aux[self.values[row]:self.values[row + 1]]).decode("utf-8")

    def serialize(self) -> bytes:
        """serialize_array for primitive layouts (llkv-column-map/src/serialization.rs:264-307): 24-byte header + values."""
        t = self.dtype.type
        if t == ffi.PT_UTF8:
            raise ValueError("varlen layout is not produced by this helper")
        if self.validity is not None:
            raise ValueError("nulls not supported in zero-copy format (yet)")
        body = self.values.tobytes()
        hdr = bytearray(b"ARR0")
        hdr += bytes([0, t, self.dtype.precision if t == ffi.PT_DECIMAL128 else 0,
                      (self.dtype.scale & 0xFF) if t == ffi.PT_DECIMAL128 else 0])
        hdr += int(self.n_rows).to_bytes(8, "little")
        hdr += len(body).to_bytes(4, "little") + (0).to_bytes(4, "little")
        return bytes(hdr) + body


class HostTable:
    """A set of equally long, densely row-numbered columns (one LLKV table)."""

    def __init__(self, table_id: int = 1):
        self.table_id = table_id
        self.columns: Dict[int, HostColumn] = {}
        self.created_by: Optional[HostColumn] = None
        self.deleted_by: Optional[HostColumn] = None

    def add(self, col: HostColumn) -> "HostTable":
        if self.columns:
            assert col.n_rows == self.n_rows, "columns of one table must be equally long"
        self.columns[col.field_id] = col
        return self

    def add_mvcc(self, created_by: np.ndarray, deleted_by: np.ndarray) -> "HostTable":
        """`_created_by` / `_deleted_by` UInt64 columns (llkv-transaction/src/mvcc.rs:475-481)."""
        self.created_by = HostColumn(0xFFFFFFF0, DataType.UInt64, created_by)
        self.deleted_by = HostColumn(0xFFFFFFF1, DataType.UInt64, deleted_by)
        assert self.created_by.n_rows == self.n_rows and self.deleted_by.n_rows == self.n_rows
        return self

    @property
    def n_rows(self) -> int:
        return next(iter(self.columns.values())).n_rows if self.columns else 0


@dataclass
class Snapshot:
    """TransactionSnapshot {txn_id, snapshot_id} (llkv-transaction/src/mvcc.rs:414-419) + the ids whose
    TxnIdManager::status is not Committed (mvcc.rs:157-171)."""
    txn_id: int
    snapshot_id: int
    noncommitted: Sequence[int] = ()
