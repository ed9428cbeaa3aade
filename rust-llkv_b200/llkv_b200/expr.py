"""Host-side mirror of the reference's predicate / scalar-expression / aggregate vocabulary.

Names, variants and argument meaning follow the reference so parity tests read like its own tests:
  Literal            llkv-types/src/literal.rs:26-41
  Expr / Filter / Operator / ScalarExpr / BinaryOp / CompareOp   llkv-expr/src/expr.rs:16-43,127-182,311-349,367-402
  ProgramCompiler    llkv-compute/src/program.rs:271-439 (postfix EvalOp program, FusedAnd for same-field ANDs)
  AggregateKind / AggregateSpec   llkv-aggregate/src/lib.rs:26-69

This module only builds and flattens trees into the C-ABI structs of include/llkv_gpu.h; it evaluates nothing.
"""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple, Union

from . import ffi


# --------------------------------------------------------------------------- Literal
@dataclass(frozen=True)
class Literal:
    kind: int
    value: object = None
    scale: int = 0

    @staticmethod
    def Null() -> "Literal":
        return Literal(ffi.LIT_NULL)

    @staticmethod
    def Int128(v: int) -> "Literal":
        return Literal(ffi.LIT_INT128, int(v))

    @staticmethod
    def Float64(v: float) -> "Literal":
        return Literal(ffi.LIT_FLOAT64, float(v))

    @staticmethod
    def Decimal128(raw: int, scale: int) -> "Literal":
        """DecimalValue::new(raw, scale) (llkv-types/src/decimal.rs:67-76)."""
        if len(str(abs(int(raw)))) > 38:
            raise ValueError("DecimalError::PrecisionOverflow")
        return Literal(ffi.LIT_DECIMAL128, int(raw), int(scale))

    @staticmethod
    def String(s: str) -> "Literal":
        return Literal(ffi.LIT_STRING, s)

    @staticmethod
    def Boolean(b: bool) -> "Literal":
        return Literal(ffi.LIT_BOOLEAN, bool(b))

    @staticmethod
    def Date32(days: int) -> "Literal":
        return Literal(ffi.LIT_DATE32, int(days))

    def to_c(self) -> ffi.Literal:
        out = ffi.Literal()
        out.kind = self.kind
        if self.kind in (ffi.LIT_INT128, ffi.LIT_DECIMAL128):
            out.lo, out.hi = ffi.i128_to_words(self.value)
            out.scale = self.scale
            out.precision = len(str(abs(self.value)))
        elif self.kind == ffi.LIT_FLOAT64:
            out.lo = struct.unpack("<Q", struct.pack("<d", self.value))[0]
        elif self.kind == ffi.LIT_BOOLEAN:
            out.lo = 1 if self.value else 0
        elif self.kind == ffi.LIT_DATE32:
            out.lo = self.value & 0xFFFFFFFFFFFFFFFF
        elif self.kind == ffi.LIT_STRING:
            b = self.value.encode("utf-8")
            if len(b) > 15:  # by reference (LLKV_LIT_STRING_BY_REF): the receiving call copies the bytes
                import ctypes
                out._bytes = ctypes.create_string_buffer(b, len(b))  # kept alive by this struct
                out.lo, out.hi, out.precision = ctypes.addressof(out._bytes), len(b), 255
            else:
                padded = b + b"\0" * (16 - len(b))
                out.lo, out.hi = struct.unpack("<QQ", padded)
                out.precision = len(b)
        return out


def lit(v) -> Literal:
    """`impl From<T> for Literal` (literal.rs:47-90)."""
    if isinstance(v, Literal):
        return v
    if isinstance(v, bool):
        return Literal.Boolean(v)
    if isinstance(v, int):
        return Literal.Int128(v)
    if isinstance(v, float):
        return Literal.Float64(v)
    if isinstance(v, str):
        return Literal.String(v)
    if v is None:
        return Literal.Null()
    raise TypeError(f"no Literal conversion for {type(v)}")


# --------------------------------------------------------------------------- Bound / Operator / Filter
@dataclass(frozen=True)
class Bound:
    kind: int
    value: Optional[Literal] = None

    @staticmethod
    def Included(v) -> "Bound":
        return Bound(ffi.BOUND_INCLUDED, lit(v))

    @staticmethod
    def Excluded(v) -> "Bound":
        return Bound(ffi.BOUND_EXCLUDED, lit(v))


Bound.Unbounded = Bound(ffi.BOUND_UNBOUNDED)


@dataclass(frozen=True)
class Operator:
    tag: int
    literals: Tuple[Literal, ...] = ()
    lower: Bound = Bound.Unbounded
    upper: Bound = Bound.Unbounded
    case_sensitive: bool = True  # StartsWith / EndsWith / Contains

    @staticmethod
    def Equals(v) -> "Operator":
        return Operator(ffi.OP_EQUALS, (lit(v),))

    @staticmethod
    def Range(lower: Bound, upper: Bound) -> "Operator":
        lits = tuple(b.value for b in (lower, upper) if b.kind != ffi.BOUND_UNBOUNDED)
        return Operator(ffi.OP_RANGE, lits, lower, upper)

    @staticmethod
    def GreaterThan(v) -> "Operator":
        return Operator(ffi.OP_GT, (lit(v),))

    @staticmethod
    def GreaterThanOrEquals(v) -> "Operator":
        return Operator(ffi.OP_GTE, (lit(v),))

    @staticmethod
    def LessThan(v) -> "Operator":
        return Operator(ffi.OP_LT, (lit(v),))

    @staticmethod
    def LessThanOrEquals(v) -> "Operator":
        return Operator(ffi.OP_LTE, (lit(v),))

    @staticmethod
    def In(values: Sequence) -> "Operator":
        return Operator(ffi.OP_IN, tuple(lit(v) for v in values))

    @staticmethod
    def StartsWith(pattern: str, case_sensitive: bool = True) -> "Operator":
        return Operator(ffi.OP_STARTS_WITH, (lit(pattern),), case_sensitive=case_sensitive)

    @staticmethod
    def EndsWith(pattern: str, case_sensitive: bool = True) -> "Operator":
        return Operator(ffi.OP_ENDS_WITH, (lit(pattern),), case_sensitive=case_sensitive)

    @staticmethod
    def Contains(pattern: str, case_sensitive: bool = True) -> "Operator":
        return Operator(ffi.OP_CONTAINS, (lit(pattern),), case_sensitive=case_sensitive)


Operator.IsNull = Operator(ffi.OP_IS_NULL)
Operator.IsNotNull = Operator(ffi.OP_IS_NOT_NULL)


@dataclass(frozen=True)
class Filter:
    field_id: int
    op: Operator


# --------------------------------------------------------------------------- ScalarExpr
class BinaryOp:
    Add, Subtract, Multiply, Divide, Modulo, And, Or, BitwiseShiftLeft, BitwiseShiftRight = range(9)


class CompareOp:
    Eq, NotEq, Lt, LtEq, Gt, GtEq = range(6)


@dataclass(frozen=True)
class DataType:
    """Arrow DataType subset that crosses the boundary (llkv-plan/src/translation/types.rs:22-44)."""
    type: int
    precision: int = 0
    scale: int = 0

    @staticmethod
    def Decimal128(p: int, s: int) -> "DataType":
        return DataType(ffi.PT_DECIMAL128, p, s)


DataType.Int64 = DataType(ffi.PT_INT64)
DataType.Int32 = DataType(ffi.PT_INT32)
DataType.UInt64 = DataType(ffi.PT_UINT64)
DataType.UInt32 = DataType(ffi.PT_UINT32)
DataType.Float64 = DataType(ffi.PT_FLOAT64)
DataType.Float32 = DataType(ffi.PT_FLOAT32)
DataType.Date32 = DataType(ffi.PT_DATE32)
DataType.Boolean = DataType(ffi.PT_BOOLEAN)
DataType.Utf8 = DataType(ffi.PT_UTF8)
DataType.Int16 = DataType(ffi.PT_INT16)
DataType.Int8 = DataType(ffi.PT_INT8)
DataType.UInt16 = DataType(ffi.PT_UINT16)
DataType.UInt8 = DataType(ffi.PT_UINT8)


@dataclass(frozen=True)
class ScalarExpr:
    tag: int
    field_id: int = 0
    literal: Optional[Literal] = None
    op: int = 0
    left: Optional["ScalarExpr"] = None
    right: Optional["ScalarExpr"] = None
    data_type: Optional[DataType] = None

    @staticmethod
    def Column(fid: int) -> "ScalarExpr":
        return ScalarExpr(ffi.SE_COLUMN, field_id=fid)

    column = Column

    @staticmethod
    def Literal(v) -> "ScalarExpr":
        return ScalarExpr(ffi.SE_LITERAL, literal=lit(v))

    literal_ = Literal

    @staticmethod
    def Binary(left: "ScalarExpr", op: int, right: "ScalarExpr") -> "ScalarExpr":
        return ScalarExpr(ffi.SE_BINARY, op=op, left=left, right=right)

    binary = Binary

    @staticmethod
    def Cast(expr: "ScalarExpr", data_type: DataType) -> "ScalarExpr":
        return ScalarExpr(ffi.SE_CAST, left=expr, data_type=data_type)

    @staticmethod
    def Compare(left: "ScalarExpr", op: int, right: "ScalarExpr") -> "ScalarExpr":
        return ScalarExpr(ffi.SE_COMPARE, op=op, left=left, right=right)

    @staticmethod
    def IsNull(expr: "ScalarExpr", negated: bool = False) -> "ScalarExpr":
        return ScalarExpr(ffi.SE_IS_NULL, op=int(negated), left=expr)

    # small sugar so tests stay readable
    def __add__(self, o): return ScalarExpr.Binary(self, BinaryOp.Add, _se(o))
    def __sub__(self, o): return ScalarExpr.Binary(self, BinaryOp.Subtract, _se(o))
    def __mul__(self, o): return ScalarExpr.Binary(self, BinaryOp.Multiply, _se(o))
    def __truediv__(self, o): return ScalarExpr.Binary(self, BinaryOp.Divide, _se(o))
    def __radd__(self, o): return ScalarExpr.Binary(_se(o), BinaryOp.Add, self)
    def __rsub__(self, o): return ScalarExpr.Binary(_se(o), BinaryOp.Subtract, self)
    def __rmul__(self, o): return ScalarExpr.Binary(_se(o), BinaryOp.Multiply, self)


def _se(v) -> ScalarExpr:
    return v if isinstance(v, ScalarExpr) else ScalarExpr.Literal(v)


class NodePool:
    """Flattens ScalarExpr trees into one llkv_scalar_node array (children before parents)."""

    def __init__(self):
        self.nodes: List[ffi.ScalarNode] = []

    def add(self, e: ScalarExpr) -> int:
        n = ffi.ScalarNode()
        n.tag = e.tag
        n.left = n.right = -1
        if e.tag == ffi.SE_COLUMN:
            n.field_id = e.field_id
        elif e.tag == ffi.SE_LITERAL:
            n.literal = e.literal.to_c()
        elif e.tag in (ffi.SE_BINARY, ffi.SE_COMPARE):
            n.op = e.op
            n.left = self.add(e.left)
            n.right = self.add(e.right)
        elif e.tag == ffi.SE_CAST:
            n.left = self.add(e.left)
            n.cast_type = e.data_type.type
            n.cast_precision = e.data_type.precision
            n.cast_scale = e.data_type.scale
        elif e.tag in (ffi.SE_IS_NULL, ffi.SE_NOT):
            n.op = e.op
            n.left = self.add(e.left)
        else:
            raise ValueError(f"ScalarExpr tag {e.tag} does not cross this boundary")
        self.nodes.append(n)
        return len(self.nodes) - 1

    def to_c(self):
        arr = (ffi.ScalarNode * max(1, len(self.nodes)))(*self.nodes)
        return arr, len(self.nodes)


# --------------------------------------------------------------------------- Expr (predicate tree)
@dataclass(frozen=True)
class Expr:
    tag: str
    children: Tuple["Expr", ...] = ()
    filter: Optional[Filter] = None
    left: Optional[ScalarExpr] = None
    right: Optional[ScalarExpr] = None
    op: int = 0
    list: Tuple[ScalarExpr, ...] = ()
    negated: bool = False
    value: bool = False

    @staticmethod
    def And(children: Sequence["Expr"]) -> "Expr":
        return Expr("And", tuple(children))

    @staticmethod
    def Or(children: Sequence["Expr"]) -> "Expr":
        return Expr("Or", tuple(children))

    @staticmethod
    def Not(inner: "Expr") -> "Expr":
        return Expr("Not", (inner,))

    @staticmethod
    def Pred(f: Filter) -> "Expr":
        return Expr("Pred", filter=f)

    @staticmethod
    def Compare(left: ScalarExpr, op: int, right: ScalarExpr) -> "Expr":
        return Expr("Compare", left=_se(left), op=op, right=_se(right))

    @staticmethod
    def InList(expr: ScalarExpr, items: Sequence, negated: bool = False) -> "Expr":
        return Expr("InList", left=_se(expr), list=tuple(_se(i) for i in items), negated=negated)

    @staticmethod
    def IsNull(expr: ScalarExpr, negated: bool = False) -> "Expr":
        return Expr("IsNull", left=_se(expr), negated=negated)

    @staticmethod
    def Literal(v: bool) -> "Expr":
        return Expr("Literal", value=bool(v))


def pred(field_id: int, op: Operator) -> Expr:
    return Expr.Pred(Filter(field_id, op))


class CompiledProgram:
    """The flattened EvalOp program handed to llkv_gpu_program_compile / the oracle."""

    def __init__(self):
        self.ops: List[ffi.EvalOp] = []
        self.literals: List[ffi.Literal] = []
        self.pool = NodePool()
        self.list_roots: List[int] = []

    def c_arrays(self):
        ops = (ffi.EvalOp * max(1, len(self.ops)))(*self.ops)
        lits = (ffi.Literal * max(1, len(self.literals)))(*self.literals)
        nodes, n_nodes = self.pool.to_c()
        roots = (C.c_int32 * max(1, len(self.list_roots)))(*self.list_roots)
        return ops, len(self.ops), lits, len(self.literals), nodes, n_nodes, roots, len(self.list_roots)


class ProgramCompiler:
    """compile_eval (llkv-compute/src/program.rs:313-413) + gather_fused (:415-439)."""

    def __init__(self, root: Expr):
        self.root = root

    def compile(self) -> CompiledProgram:
        prog = CompiledProgram()
        self._emit(self.root, prog)
        return prog

    @staticmethod
    def _gather_fused(children: Sequence[Expr]):
        if not children:
            return None
        fid = None
        for c in children:
            if c.tag != "Pred":
                return None
            if fid is None:
                fid = c.filter.field_id
            elif fid != c.filter.field_id:
                return None
        return fid

    def _filter_op(self, tag: int, f: Filter, prog: CompiledProgram) -> ffi.EvalOp:
        op = ffi.EvalOp()
        op.tag = tag
        op.operator_tag = f.op.tag
        op.field_id = f.field_id
        op.lower_kind = f.op.lower.kind
        op.upper_kind = f.op.upper.kind
        op.lit_begin = len(prog.literals)
        op.lit_count = len(f.op.literals)
        op.literal_bool = 0 if f.op.case_sensitive else 1
        prog.literals.extend(l.to_c() for l in f.op.literals)
        return op

    def _emit(self, node: Expr, prog: CompiledProgram):
        if node.tag == "And":
            if not node.children:
                raise ValueError("AND expression requires at least one predicate")
            fid = self._gather_fused(node.children)
            if fid is not None:
                op = ffi.EvalOp()
                op.tag = ffi.EV_FUSED_AND
                op.field_id = fid
                op.child_count = len(node.children)
                prog.ops.append(op)
                for c in node.children:
                    prog.ops.append(self._filter_op(ffi.EV_FILTER_ITEM, c.filter, prog))
                return
            for c in node.children:
                self._emit(c, prog)
            op = ffi.EvalOp()
            op.tag = ffi.EV_AND
            op.child_count = len(node.children)
            prog.ops.append(op)
        elif node.tag == "Or":
            if not node.children:
                raise ValueError("OR expression requires at least one predicate")
            for c in node.children:
                self._emit(c, prog)
            op = ffi.EvalOp()
            op.tag = ffi.EV_OR
            op.child_count = len(node.children)
            prog.ops.append(op)
        elif node.tag == "Not":
            self._emit(node.children[0], prog)
            op = ffi.EvalOp()
            op.tag = ffi.EV_NOT
            prog.ops.append(op)
        elif node.tag == "Pred":
            prog.ops.append(self._filter_op(ffi.EV_PUSH_PREDICATE, node.filter, prog))
        elif node.tag == "Compare":
            op = ffi.EvalOp()
            op.tag = ffi.EV_PUSH_COMPARE
            op.expr_left = prog.pool.add(node.left)
            op.expr_right = prog.pool.add(node.right)
            op.cmp_op = node.op
            prog.ops.append(op)
        elif node.tag == "InList":
            op = ffi.EvalOp()
            op.tag = ffi.EV_PUSH_IN_LIST
            op.expr_left = prog.pool.add(node.left)
            op.expr_right = len(prog.list_roots)
            op.child_count = len(node.list)
            op.negated = int(node.negated)
            for item in node.list:
                prog.list_roots.append(prog.pool.add(item))
            prog.ops.append(op)
        elif node.tag == "IsNull":
            op = ffi.EvalOp()
            op.tag = ffi.EV_PUSH_IS_NULL
            op.expr_left = prog.pool.add(node.left)
            op.negated = int(node.negated)
            prog.ops.append(op)
        elif node.tag == "Literal":
            op = ffi.EvalOp()
            op.tag = ffi.EV_PUSH_LITERAL
            op.literal_bool = int(node.value)
            prog.ops.append(op)
        else:
            raise ValueError(f"Expr::{node.tag} is not supported in storage evaluation")


# --------------------------------------------------------------------------- aggregates
@dataclass(frozen=True)
class AggregateKind:
    kind: int
    expr: Optional[ScalarExpr]
    data_type: DataType = DataType.Int64
    distinct: bool = False

    @staticmethod
    def CountStar() -> "AggregateKind":
        return AggregateKind(ffi.AGG_COUNT, None)

    @staticmethod
    def Count(expr, distinct: bool = False) -> "AggregateKind":
        return AggregateKind(ffi.AGG_COUNT, _col(expr), distinct=distinct)

    @staticmethod
    def Sum(expr, data_type: DataType, distinct: bool = False) -> "AggregateKind":
        return AggregateKind(ffi.AGG_SUM, _col(expr), data_type, distinct)

    @staticmethod
    def Total(expr, data_type: DataType, distinct: bool = False) -> "AggregateKind":
        return AggregateKind(ffi.AGG_TOTAL, _col(expr), data_type, distinct)

    @staticmethod
    def Avg(expr, data_type: DataType, distinct: bool = False) -> "AggregateKind":
        return AggregateKind(ffi.AGG_AVG, _col(expr), data_type, distinct)

    @staticmethod
    def Min(expr, data_type: DataType) -> "AggregateKind":
        return AggregateKind(ffi.AGG_MIN, _col(expr), data_type)

    @staticmethod
    def Max(expr, data_type: DataType) -> "AggregateKind":
        return AggregateKind(ffi.AGG_MAX, _col(expr), data_type)

    @staticmethod
    def CountNulls(expr) -> "AggregateKind":
        return AggregateKind(ffi.AGG_COUNT_NULLS, _col(expr))


def _col(e) -> ScalarExpr:
    return ScalarExpr.Column(e) if isinstance(e, int) else e


@dataclass(frozen=True)
class AggregateSpec:
    alias: str
    kind: AggregateKind


def flatten_aggregates(specs: Sequence[AggregateSpec]):
    pool = NodePool()
    out = []
    for s in specs:
        a = ffi.AggSpec()
        a.kind = s.kind.kind
        a.expr_root = -1 if s.kind.expr is None else pool.add(s.kind.expr)
        a.data_type = s.kind.data_type.type
        a.precision = s.kind.data_type.precision
        a.scale = s.kind.data_type.scale
        a.distinct = int(s.kind.distinct)
        out.append(a)
    arr = (ffi.AggSpec * max(1, len(out)))(*out)
    nodes, n_nodes = pool.to_c()
    return arr, len(out), nodes, n_nodes


# --------------------------------------------------------------------------- results
@dataclass(frozen=True)
class AggregateValue:
    """AggregateValue (llkv-executor/src/lib.rs:6590-6662): Null | Int64 | Float64 | Decimal128{value,scale}."""
    type: int
    value: object  # None for NULL
    precision: int = 0
    scale: int = 0

    @staticmethod
    def from_c(v: ffi.AggValue) -> "AggregateValue":
        if not v.valid:
            return AggregateValue(v.type, None, v.precision, v.scale)
        if v.type == ffi.PT_FLOAT64:
            return AggregateValue(v.type, struct.unpack("<d", struct.pack("<Q", v.lo))[0])
        if v.type == ffi.PT_DECIMAL128:
            return AggregateValue(v.type, ffi.words_to_i128(v.lo, v.hi), v.precision, v.scale)
        if v.type == ffi.PT_UINT64:
            return AggregateValue(v.type, int(v.lo))
        val = v.lo - (1 << 64) if v.lo >> 63 else v.lo
        return AggregateValue(v.type, int(val))


def decode_group_key(k: ffi.GroupKey, resolve=None):
    """GroupKeyValue (llkv-executor/src/lib.rs:99-106).  `resolve(dict_kind, bits)` turns a dictionary-coded string key
    (llkv_group_key.dict != 0) into its string."""
    if not k.valid:
        return None
    if k.type == ffi.PT_UTF8 and k.dict:
        return resolve(int(k.dict), int(k.bits))
    if k.type == ffi.PT_UTF8:
        n = k.bits & 0xFF
        return bytes((k.bits >> (56 - 8 * i)) & 0xFF for i in range(n)).decode("utf-8")
    if k.type == ffi.PT_BOOLEAN:
        return bool(k.bits)
    if k.type in (ffi.PT_UINT64, ffi.PT_UINT32, ffi.PT_UINT16, ffi.PT_UINT8):
        return int(k.bits)
    return int(k.bits - (1 << 64) if k.bits >> 63 else k.bits)
