"""ctypes wrapper around oracle/libllkv_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.  The product
package (rust-llkv_b200/llkv_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(os.path.dirname(_HERE), "rust-llkv_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from llkv_b200 import ffi  # noqa: E402
from llkv_b200.expr import (AggregateSpec, AggregateValue, Expr, ProgramCompiler, decode_group_key,  # noqa: E402
                            flatten_aggregates)
from llkv_b200.table import HostColumn, HostTable, LlkvError, Snapshot  # noqa: E402


class OracleColumn(C.Structure):
    _fields_ = [("field_id", C.c_uint64), ("type", C.c_int32), ("precision", C.c_uint8), ("scale", C.c_int8),
                ("_pad", C.c_uint8 * 2), ("n_rows", C.c_uint64), ("values", C.c_void_p), ("validity", C.c_void_p),
                ("aux", C.c_void_p)]


class OracleMvcc(C.Structure):
    _fields_ = [("created_by", C.POINTER(OracleColumn)), ("deleted_by", C.POINTER(OracleColumn)),
                ("txn_id", C.c_uint64), ("snapshot_id", C.c_uint64), ("noncommitted", C.POINTER(C.c_uint64)),
                ("n_noncommitted", C.c_int32)]


class OracleProgram(C.Structure):
    _fields_ = [("ops", C.POINTER(ffi.EvalOp)), ("n_ops", C.c_int32), ("literals", C.POINTER(ffi.Literal)),
                ("n_literals", C.c_int32), ("nodes", C.POINTER(ffi.ScalarNode)), ("n_nodes", C.c_int32),
                ("list_roots", C.POINTER(C.c_int32)), ("n_list_roots", C.c_int32)]


def build() -> str:
    """Compiles the C restatement (gcc) if the .so is missing or stale; returns its path."""
    so = os.path.join(_HERE, "libllkv_oracle.so")
    srcs = [os.path.join(_HERE, "llkv_oracle.c"), os.path.join(_HERE, "llkv_oracle.h"),
            os.path.join(os.path.dirname(_HERE), "include", "llkv_gpu.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.llkv_oracle_filter.restype = C.c_int32
        _lib.llkv_oracle_aggregate.restype = C.c_int32
        _lib.llkv_oracle_mvcc_visible.restype = C.c_int32
        _lib.llkv_oracle_mvcc_visible.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                                  C.POINTER(C.c_uint64), C.c_int32]
        _lib.llkv_oracle_serialize_primitive.restype = C.c_int64
        _lib.llkv_oracle_decimal_binary.restype = C.c_int32
    return _lib


def _ocol(c: HostColumn) -> OracleColumn:
    o = OracleColumn()
    o.field_id = c.field_id
    o.type = c.dtype.type
    o.precision = c.dtype.precision
    o.scale = c.dtype.scale
    o.n_rows = c.n_rows
    o.values = c.values.ctypes.data
    o.validity = c.validity.ctypes.data if c.validity is not None else None
    o.aux = c.aux.ctypes.data if c.aux is not None else None
    return o


class _Bound:
    """Keeps every ctypes buffer alive for the duration of one call."""

    def __init__(self, table: HostTable, expr: Optional[Expr], snapshot: Optional[Snapshot]):
        cols = list(table.columns.values())
        self.cols = (OracleColumn * max(1, len(cols)))(*[_ocol(c) for c in cols])
        self.n_cols = len(cols)
        self.prog = OracleProgram()
        self.keep = []
        if expr is not None:
            cp = ProgramCompiler(expr).compile()
            ops, n_ops, lits, n_lits, nodes, n_nodes, roots, n_roots = cp.c_arrays()
            self.keep += [cp, ops, lits, nodes, roots]  # (cp owns the bytes of by-reference string literals)
            self.prog.ops, self.prog.n_ops = ops, n_ops
            self.prog.literals, self.prog.n_literals = lits, n_lits
            self.prog.nodes, self.prog.n_nodes = nodes, n_nodes
            self.prog.list_roots, self.prog.n_list_roots = roots, n_roots
        self.mvcc = None
        if snapshot is not None and table.created_by is not None:
            self.cb = _ocol(table.created_by)
            self.db = _ocol(table.deleted_by)
            nc = list(snapshot.noncommitted)
            self.nc = (C.c_uint64 * max(1, len(nc)))(*nc)
            m = OracleMvcc()
            m.created_by = C.pointer(self.cb)
            m.deleted_by = C.pointer(self.db)
            m.txn_id = snapshot.txn_id
            m.snapshot_id = snapshot.snapshot_id
            m.noncommitted = self.nc
            m.n_noncommitted = len(nc)
            self.mvcc = m


def filter_bitmap(table: HostTable, expr: Optional[Expr], snapshot: Optional[Snapshot] = None,
                  row_begin: int = 0, row_end: Optional[int] = None, n_threads: int = 1) -> Tuple[np.ndarray, int]:
    row_end = table.n_rows if row_end is None else row_end
    b = _Bound(table, expr, snapshot)
    n_words = (row_end - row_begin + 63) // 64
    words = np.zeros(max(1, n_words), dtype=np.uint64)
    count = C.c_uint64(0)
    err = C.create_string_buffer(512)
    rc = lib().llkv_oracle_filter(b.cols, b.n_cols, C.byref(b.prog), C.byref(b.mvcc) if b.mvcc else None,
                                  C.c_uint64(row_begin), C.c_uint64(row_end), C.c_int32(n_threads),
                                  words.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_uint64(n_words), C.byref(count),
                                  err, C.c_size_t(512))
    if rc:
        raise LlkvError(rc, err.value.decode())
    return words[:n_words], int(count.value)


def aggregate(table: HostTable, expr: Optional[Expr], specs: Sequence[AggregateSpec],
              snapshot: Optional[Snapshot] = None, group_by: Sequence[int] = (), expr_mode: Optional[int] = None,
              row_begin: int = 0, row_end: Optional[int] = None, n_threads: int = 1, group_capacity: int = 1 << 16):
    """Returns [(key_tuple, [AggregateValue, ...]), ...] in first-appearance order of the groups."""
    row_end = table.n_rows if row_end is None else row_end
    if expr_mode is None:
        expr_mode = ffi.EXPR_EXACT if group_by else ffi.EXPR_ARROW
    b = _Bound(table, expr, snapshot)
    aggs, n_aggs, nodes, n_nodes = flatten_aggregates(specs)
    keys = (C.c_uint64 * max(1, len(group_by)))(*group_by)
    cap = group_capacity if group_by else 1
    out_vals = (ffi.AggValue * (cap * max(1, n_aggs)))()
    out_keys = (ffi.GroupKey * (cap * max(1, len(group_by))))()
    n_groups = C.c_uint64(0)
    err = C.create_string_buffer(512)
    rc = lib().llkv_oracle_aggregate(b.cols, b.n_cols, C.byref(b.prog), C.byref(b.mvcc) if b.mvcc else None, aggs,
                                     C.c_int32(n_aggs), nodes, C.c_int32(n_nodes), keys, C.c_int32(len(group_by)),
                                     C.c_int32(expr_mode), C.c_uint64(row_begin), C.c_uint64(row_end),
                                     C.c_int32(n_threads), out_vals, out_keys, C.c_uint64(cap), C.byref(n_groups), err,
                                     C.c_size_t(512))
    if rc:
        raise LlkvError(rc, err.value.decode())
    rows = []
    nk = len(group_by)
    for g in range(n_groups.value):
        key = tuple(decode_group_key(out_keys[g * nk + k], lambda kind, row, k=k: table.columns[group_by[k]].string_at(row)) for k in range(nk))
        vals = [AggregateValue.from_c(out_vals[g * n_aggs + a]) for a in range(n_aggs)]
        rows.append((key, vals))
    return rows


def mvcc_visible(created_by: int, deleted_by: int, txn_id: int, snapshot_id: int, noncommitted: Sequence[int] = ()) -> bool:
    nc = (C.c_uint64 * max(1, len(noncommitted)))(*noncommitted)
    return bool(lib().llkv_oracle_mvcc_visible(created_by, deleted_by, txn_id, snapshot_id, nc, len(noncommitted)))


def serialize_primitive(col: HostColumn) -> bytes:
    n = col.values.nbytes + 24
    out = C.create_string_buffer(n)
    rc = lib().llkv_oracle_serialize_primitive(C.c_int32(col.dtype.type), C.c_uint8(col.dtype.precision),
                                               C.c_int8(col.dtype.scale), C.c_void_p(col.values.ctypes.data),
                                               C.c_uint64(col.n_rows), out, C.c_uint64(n))
    if rc < 0:
        raise LlkvError(-rc, "serialize failed")
    return out.raw[:rc]
