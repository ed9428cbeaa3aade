//! `Expr<FieldId>` / `ScalarExpr<FieldId>` / `AggregateCall<FieldId>` -> the flat arrays of `include/llkv_gpu.h`.
//!
//! The predicate side mirrors `ProgramCompiler::compile` (`llkv-compute/src/program.rs:271-439`): a postfix walk that
//! emits one `llkv_eval_op` per `EvalOp`, with the `gather_fused` rule (an AND whose children are all `Pred`s on one field
//! becomes `FusedAnd` + its filter items) and `Not` carrying no domain program of its own — the device tracks the domain
//! of every operand (three-valued logic, `llkv-scan/src/predicate.rs:167-186,665-777`).  The C++ mirror of this file is
//! `rust-llkv_b200/host/llkv_gpu.hpp: ProgramCompiler`; `tests/test_cpp_host.py` checks that mirror byte for byte against
//! the Python one, and `tests/test_rust_ffi.py` checks the tag tables and struct layouts below against the header.
//! Authored against jzombie/rust-llkv v0.8.5-alpha; not compiled here (no Rust toolchain in the image).
use std::ops::Bound;

use arrow::datatypes::DataType;
use llkv_expr::expr::{AggregateCall, BinaryOp, CompareOp, Expr, Filter, Operator, ScalarExpr};
use llkv_gpu_sys as sys;
use llkv_result::{Error, Result};
use llkv_types::{FieldId, Literal};

// ---- tag tables: the header's enums, in the order of the reference's variants -------------------------------------
pub const EV_PUSH_PREDICATE: i32 = 0;
pub const EV_PUSH_COMPARE: i32 = 1;
pub const EV_PUSH_IN_LIST: i32 = 2;
pub const EV_PUSH_IS_NULL: i32 = 3;
pub const EV_PUSH_LITERAL: i32 = 4;
pub const EV_FUSED_AND: i32 = 5;
pub const EV_AND: i32 = 6;
pub const EV_OR: i32 = 7;
pub const EV_NOT: i32 = 8;
pub const EV_FILTER_ITEM: i32 = 100;

pub const OP_EQUALS: i32 = 0;
pub const OP_RANGE: i32 = 1;
pub const OP_GT: i32 = 2;
pub const OP_GTE: i32 = 3;
pub const OP_LT: i32 = 4;
pub const OP_LTE: i32 = 5;
pub const OP_IN: i32 = 6;
pub const OP_STARTS_WITH: i32 = 7;
pub const OP_ENDS_WITH: i32 = 8;
pub const OP_CONTAINS: i32 = 9;
pub const OP_IS_NULL: i32 = 10;
pub const OP_IS_NOT_NULL: i32 = 11;

pub const BOUND_INCLUDED: i32 = 0;
pub const BOUND_EXCLUDED: i32 = 1;
pub const BOUND_UNBOUNDED: i32 = 2;

pub const SE_COLUMN: i32 = 0;
pub const SE_LITERAL: i32 = 1;
pub const SE_BINARY: i32 = 2;
pub const SE_NOT: i32 = 3;
pub const SE_IS_NULL: i32 = 4;
pub const SE_CAST: i32 = 7;
pub const SE_COMPARE: i32 = 8;
pub const SE_COALESCE: i32 = 9;

pub const LIT_NULL: i32 = 0;
pub const LIT_INT128: i32 = 1;
pub const LIT_FLOAT64: i32 = 2;
pub const LIT_DECIMAL128: i32 = 3;
pub const LIT_STRING: i32 = 4;
pub const LIT_BOOLEAN: i32 = 5;
pub const LIT_DATE32: i32 = 6;

pub const AGG_COUNT: i32 = 0;
pub const AGG_SUM: i32 = 1;
pub const AGG_TOTAL: i32 = 2;
pub const AGG_AVG: i32 = 3;
pub const AGG_MIN: i32 = 4;
pub const AGG_MAX: i32 = 5;
pub const AGG_COUNT_NULLS: i32 = 6;

// on-disk PrimType codes (llkv-column-map/src/serialization.rs:146-166)
/// Bytes per value of the Arrow layout a `PrimType` crosses the boundary in (Boolean: one byte; Utf8: its int32 offsets).
pub fn prim_type_width(prim_type: i32) -> usize {
    match prim_type {
        PT_INT8 | PT_UINT8 | PT_BOOLEAN => 1,
        PT_INT16 | PT_UINT16 => 2,
        PT_INT32 | PT_UINT32 | PT_FLOAT32 | PT_DATE32 | PT_UTF8 => 4,
        PT_DECIMAL128 => 16,
        _ => 8,
    }
}

pub const PT_NULL: i32 = 0;
pub const PT_UINT64: i32 = 1;
pub const PT_INT32: i32 = 2;
pub const PT_UINT32: i32 = 3;
pub const PT_FLOAT32: i32 = 4;
pub const PT_INT64: i32 = 6;
pub const PT_INT16: i32 = 7;
pub const PT_INT8: i32 = 8;
pub const PT_UINT16: i32 = 9;
pub const PT_UINT8: i32 = 10;
pub const PT_FLOAT64: i32 = 11;
pub const PT_UTF8: i32 = 12;
pub const PT_BOOLEAN: i32 = 15;
pub const PT_DATE32: i32 = 16;
pub const PT_DATE64: i32 = 17;
pub const PT_DECIMAL128: i32 = 18;

fn binary_op_code(op: BinaryOp) -> i32 {
    match op {
        BinaryOp::Add => 0,
        BinaryOp::Subtract => 1,
        BinaryOp::Multiply => 2,
        BinaryOp::Divide => 3,
        BinaryOp::Modulo => 4,
        BinaryOp::And => 5,
        BinaryOp::Or => 6,
        BinaryOp::BitwiseShiftLeft => 7,
        BinaryOp::BitwiseShiftRight => 8,
    }
}

fn compare_op_code(op: CompareOp) -> i32 {
    match op {
        CompareOp::Eq => 0,
        CompareOp::NotEq => 1,
        CompareOp::Lt => 2,
        CompareOp::LtEq => 3,
        CompareOp::Gt => 4,
        CompareOp::GtEq => 5,
    }
}

/// Arrow `DataType` -> (LLKV_PT_*, precision, scale) for the types that cross the boundary.
pub fn prim_type_of(dt: &DataType) -> Result<(i32, u8, i8)> {
    Ok(match dt {
        DataType::Null => (PT_NULL, 0, 0),
        DataType::UInt64 => (PT_UINT64, 0, 0),
        DataType::Int32 => (PT_INT32, 0, 0),
        DataType::UInt32 => (PT_UINT32, 0, 0),
        DataType::Float32 => (PT_FLOAT32, 0, 0),
        DataType::Int64 => (PT_INT64, 0, 0),
        DataType::Int16 => (PT_INT16, 0, 0),
        DataType::Int8 => (PT_INT8, 0, 0),
        DataType::UInt16 => (PT_UINT16, 0, 0),
        DataType::UInt8 => (PT_UINT8, 0, 0),
        DataType::Float64 => (PT_FLOAT64, 0, 0),
        DataType::Utf8 => (PT_UTF8, 0, 0),
        DataType::Boolean => (PT_BOOLEAN, 0, 0),
        DataType::Date32 => (PT_DATE32, 0, 0),
        DataType::Date64 => (PT_DATE64, 0, 0),
        DataType::Decimal128(p, s) => (PT_DECIMAL128, *p, *s),
        other => return Err(Error::InvalidArgumentError(format!("type {other:?} does not cross the GPU boundary"))),
    })
}

fn empty_literal() -> sys::llkv_literal {
    sys::llkv_literal { kind: LIT_NULL, precision: 0, scale: 0, _pad: [0; 2], lo: 0, hi: 0 }
}

/// `llkv_types::Literal` (`llkv-types/src/literal.rs:26-41`) -> `llkv_literal`.  Strings of up to 15 bytes travel inline
/// (longer ones by reference: `FlatProgram::literal`); struct and interval literals have no place on this path.
pub fn literal_to_c(lit: &Literal) -> Result<sys::llkv_literal> {
    let mut out = empty_literal();
    match lit {
        Literal::Null => {}
        Literal::Int128(v) => {
            out.kind = LIT_INT128;
            out.lo = *v as u64;
            out.hi = ((*v as u128) >> 64) as u64;
        }
        Literal::Float64(v) => {
            out.kind = LIT_FLOAT64;
            out.lo = v.to_bits();
        }
        Literal::Decimal128(d) => {
            out.kind = LIT_DECIMAL128;
            let raw = d.raw_value();
            out.lo = raw as u64;
            out.hi = ((raw as u128) >> 64) as u64;
            out.scale = d.scale();
            out.precision = d.precision();
        }
        Literal::String(s) => {
            let b = s.as_bytes();
            if b.len() > 15 {
                return Err(Error::PredicateBuild(format!("string literal of {} bytes is too long for the GPU path", b.len())));
            }
            out.kind = LIT_STRING;
            out.precision = b.len() as u8;
            let mut bytes = [0u8; 16];
            bytes[..b.len()].copy_from_slice(b);
            out.lo = u64::from_le_bytes(bytes[..8].try_into().unwrap());
            out.hi = u64::from_le_bytes(bytes[8..].try_into().unwrap());
        }
        Literal::Boolean(v) => {
            out.kind = LIT_BOOLEAN;
            out.lo = *v as u64;
        }
        Literal::Date32(d) => {
            out.kind = LIT_DATE32;
            out.lo = *d as i64 as u64;
        }
        other => return Err(Error::PredicateBuild(format!("literal {other:?} is not supported on the GPU path"))),
    }
    Ok(out)
}

/// Everything `llkv_gpu_program_compile` takes.
#[derive(Default)]
pub struct FlatProgram {
    pub ops: Vec<sys::llkv_eval_op>,
    pub literals: Vec<sys::llkv_literal>,
    pub nodes: Vec<sys::llkv_scalar_node>,
    pub list_roots: Vec<i32>,
    /// Bytes behind string literals passed by reference (`LLKV_LIT_STRING_BY_REF`, strings over 15 bytes).  Boxed, so the
    /// addresses stored in `literals` survive this vector growing; `llkv_gpu_program_compile` copies them.
    pub strings: Vec<Box<[u8]>>,
}

pub const LIT_STRING_BY_REF: u8 = 255;

impl FlatProgram {
    /// The C form of `lit`: long strings are owned by the program and referenced, everything else travels inline.
    fn literal(&mut self, lit: &Literal) -> Result<sys::llkv_literal> {
        match lit {
            Literal::String(s) if s.len() > 15 => Ok(self.string_by_ref(s.as_bytes())),
            other => literal_to_c(other),
        }
    }
    fn string_by_ref(&mut self, bytes: &[u8]) -> sys::llkv_literal {
        self.strings.push(bytes.to_vec().into_boxed_slice());
        let stored = self.strings.last().unwrap();
        let mut out = empty_literal();
        out.kind = LIT_STRING;
        out.precision = LIT_STRING_BY_REF;
        out.lo = stored.as_ptr() as u64;
        out.hi = stored.len() as u64;
        out
    }
}

fn empty_op(tag: i32) -> sys::llkv_eval_op {
    sys::llkv_eval_op {
        tag, operator_tag: 0, field_id: 0, lower_kind: BOUND_UNBOUNDED, upper_kind: BOUND_UNBOUNDED, lit_begin: 0, lit_count: 0,
        expr_left: -1, expr_right: -1, cmp_op: 0, negated: 0, child_count: 0, literal_bool: 0,
    }
}

fn empty_node(tag: i32) -> sys::llkv_scalar_node {
    sys::llkv_scalar_node {
        tag, op: 0, left: -1, right: -1, field_id: 0, literal: empty_literal(), cast_type: 0, cast_precision: 0, cast_scale: 0, _pad: [0; 2],
    }
}

/// `ScalarExpr` -> node pool, children before parents; returns the root's index.
pub fn flatten_scalar(expr: &ScalarExpr<FieldId>, nodes: &mut Vec<sys::llkv_scalar_node>) -> Result<i32> {
    let node = match expr {
        ScalarExpr::Column(fid) => {
            let mut n = empty_node(SE_COLUMN);
            n.field_id = *fid as u64;
            n
        }
        ScalarExpr::Literal(lit) => {
            let mut n = empty_node(SE_LITERAL);
            n.literal = literal_to_c(lit)?;
            n
        }
        ScalarExpr::Binary { left, op, right } => {
            let l = flatten_scalar(left, nodes)?;
            let r = flatten_scalar(right, nodes)?;
            let mut n = empty_node(SE_BINARY);
            n.op = binary_op_code(*op);
            n.left = l;
            n.right = r;
            n
        }
        ScalarExpr::Not(inner) => {
            let l = flatten_scalar(inner, nodes)?;
            let mut n = empty_node(SE_NOT);
            n.left = l;
            n
        }
        ScalarExpr::IsNull { expr, negated } => {
            let l = flatten_scalar(expr, nodes)?;
            let mut n = empty_node(SE_IS_NULL);
            n.left = l;
            n.op = *negated as i32;
            n
        }
        ScalarExpr::Cast { expr, data_type } => {
            let l = flatten_scalar(expr, nodes)?;
            let (t, p, s) = prim_type_of(data_type)?;
            let mut n = empty_node(SE_CAST);
            n.left = l;
            n.cast_type = t;
            n.cast_precision = p;
            n.cast_scale = s;
            n
        }
        ScalarExpr::Compare { left, op, right } => {
            let l = flatten_scalar(left, nodes)?;
            let r = flatten_scalar(right, nodes)?;
            let mut n = empty_node(SE_COMPARE);
            n.op = compare_op_code(*op);
            n.left = l;
            n.right = r;
            n
        }
        ScalarExpr::Coalesce(items) if items.len() == 2 => {
            let l = flatten_scalar(&items[0], nodes)?;
            let r = flatten_scalar(&items[1], nodes)?;
            let mut n = empty_node(SE_COALESCE);
            n.left = l;
            n.right = r;
            n
        }
        // Aggregate calls inside expressions, struct field access, CASE, subqueries, RANDOM: the recogniser keeps such plans
        // on the reference's own path (`GpuPath::recognise`)
        other => return Err(Error::InvalidArgumentError(format!("scalar expression {other:?} is not supported on the GPU path"))),
    };
    nodes.push(node);
    Ok(nodes.len() as i32 - 1)
}

fn push_operator(out: &mut FlatProgram, op: &Operator<'_>, item: &mut sys::llkv_eval_op) -> Result<()> {
    item.lit_begin = out.literals.len() as i32;
    let mut one = |out: &mut FlatProgram, tag: i32, lit: &Literal| -> Result<()> {
        item.operator_tag = tag;
        let c = out.literal(lit)?;
        out.literals.push(c);
        Ok(())
    };
    match op {
        Operator::Equals(l) => one(out, OP_EQUALS, l)?,
        Operator::GreaterThan(l) => one(out, OP_GT, l)?,
        Operator::GreaterThanOrEquals(l) => one(out, OP_GTE, l)?,
        Operator::LessThan(l) => one(out, OP_LT, l)?,
        Operator::LessThanOrEquals(l) => one(out, OP_LTE, l)?,
        Operator::Range { lower, upper } => {
            item.operator_tag = OP_RANGE;
            // lower (if bounded) then upper (if bounded): the order the header documents for lit_begin
            for (bound, kind) in [(lower, &mut item.lower_kind), (upper, &mut item.upper_kind)] {
                match bound {
                    Bound::Included(l) => {
                        *kind = BOUND_INCLUDED;
                        let c = out.literal(l)?;
                        out.literals.push(c);
                    }
                    Bound::Excluded(l) => {
                        *kind = BOUND_EXCLUDED;
                        let c = out.literal(l)?;
                        out.literals.push(c);
                    }
                    Bound::Unbounded => *kind = BOUND_UNBOUNDED,
                }
            }
        }
        Operator::In(list) => {
            item.operator_tag = OP_IN;
            for l in list.iter() {
                let c = out.literal(l)?;
                        out.literals.push(c);
            }
        }
        Operator::IsNull => item.operator_tag = OP_IS_NULL,
        Operator::IsNotNull => item.operator_tag = OP_IS_NOT_NULL,
        // the pattern travels as a string literal; llkv_eval_op.literal_bool = 1 asks for the case-insensitive form
        Operator::StartsWith { pattern, case_sensitive }
        | Operator::EndsWith { pattern, case_sensitive }
        | Operator::Contains { pattern, case_sensitive } => {
            item.operator_tag = match op {
                Operator::StartsWith { .. } => OP_STARTS_WITH,
                Operator::EndsWith { .. } => OP_ENDS_WITH,
                _ => OP_CONTAINS,
            };
            item.literal_bool = i32::from(!*case_sensitive);
            let c = out.literal(&Literal::String(pattern.clone()))?; // (long patterns are copied into out.strings)
            out.literals.push(c);
        }
    }
    item.lit_count = out.literals.len() as i32 - item.lit_begin;
    Ok(())
}

fn push_filter(out: &mut FlatProgram, tag: i32, f: &Filter<'_, FieldId>) -> Result<()> {
    let mut op = empty_op(tag);
    op.field_id = f.field_id as u64;
    push_operator(out, &f.op, &mut op)?;
    out.ops.push(op);
    Ok(())
}

/// `gather_fused` (`llkv-compute/src/program.rs:415-439`): the children of an AND when every one is a `Pred` on the same field.
fn fused_field<'a, 'e>(children: &'a [Expr<'e, FieldId>]) -> Option<FieldId> {
    let mut field = None;
    for c in children {
        match c {
            Expr::Pred(f) => match field {
                None => field = Some(f.field_id),
                Some(x) if x == f.field_id => {}
                _ => return None,
            },
            _ => return None,
        }
    }
    if children.len() >= 2 { field } else { None }
}

fn flatten_into(expr: &Expr<'_, FieldId>, out: &mut FlatProgram) -> Result<()> {
    match expr {
        Expr::Pred(f) => push_filter(out, EV_PUSH_PREDICATE, f)?,
        Expr::And(children) | Expr::Or(children) if children.is_empty() => {
            return Err(Error::InvalidArgumentError("AND / OR without operands".into()));
        }
        Expr::And(children) => {
            if let Some(field) = fused_field(children) {
                let mut op = empty_op(EV_FUSED_AND);
                op.field_id = field as u64;
                op.child_count = children.len() as i32;
                out.ops.push(op);
                for c in children {
                    if let Expr::Pred(f) = c {
                        push_filter(out, EV_FILTER_ITEM, f)?;
                    }
                }
            } else {
                for c in children {
                    flatten_into(c, out)?;
                }
                let mut op = empty_op(EV_AND);
                op.child_count = children.len() as i32;
                out.ops.push(op);
            }
        }
        Expr::Or(children) => {
            for c in children {
                flatten_into(c, out)?;
            }
            let mut op = empty_op(EV_OR);
            op.child_count = children.len() as i32;
            out.ops.push(op);
        }
        Expr::Not(inner) => {
            flatten_into(inner, out)?;
            out.ops.push(empty_op(EV_NOT));
        }
        Expr::Compare { left, op, right } => {
            let l = flatten_scalar(left, &mut out.nodes)?;
            let r = flatten_scalar(right, &mut out.nodes)?;
            let mut e = empty_op(EV_PUSH_COMPARE);
            e.expr_left = l;
            e.expr_right = r;
            e.cmp_op = compare_op_code(*op);
            out.ops.push(e);
        }
        Expr::InList { expr, list, negated } => {
            let target = flatten_scalar(expr, &mut out.nodes)?;
            let first = out.list_roots.len() as i32;
            for item in list {
                let root = flatten_scalar(item, &mut out.nodes)?;
                out.list_roots.push(root);
            }
            let mut e = empty_op(EV_PUSH_IN_LIST);
            e.expr_left = target;
            e.expr_right = first;
            e.child_count = list.len() as i32;
            e.negated = *negated as i32;
            out.ops.push(e);
        }
        Expr::IsNull { expr, negated } => {
            let root = flatten_scalar(expr, &mut out.nodes)?;
            let mut e = empty_op(EV_PUSH_IS_NULL);
            e.expr_left = root;
            e.negated = *negated as i32;
            out.ops.push(e);
        }
        Expr::Literal(v) => {
            let mut e = empty_op(EV_PUSH_LITERAL);
            e.literal_bool = *v as i32;
            out.ops.push(e);
        }
        Expr::Exists(_) => return Err(Error::InvalidArgumentError("correlated subqueries stay on the reference's path".into())),
    }
    Ok(())
}

/// `ProgramCompiler::new(Arc::new(expr)).compile()` for the device.
pub fn flatten_expr(expr: &Expr<'_, FieldId>) -> Result<FlatProgram> {
    let mut out = FlatProgram::default();
    flatten_into(expr, &mut out)?;
    Ok(out)
}

/// `validate_aggregate_type` (`llkv-executor/src/lib.rs:5946-5988`): SUM / AVG / TOTAL / MIN / MAX take Int64, Float64 and
/// Decimal128 as they are and see Utf8, Boolean, Date32 and Null inputs as Float64; anything else is an error.
pub fn normalise_aggregate_type(dt: &DataType, func_name: &str) -> Result<DataType> {
    match dt {
        DataType::Int64 | DataType::Float64 | DataType::Decimal128(_, _) => Ok(dt.clone()),
        DataType::Utf8 | DataType::Boolean | DataType::Date32 | DataType::Null => Ok(DataType::Float64),
        other => Err(Error::InvalidArgumentError(format!("{func_name} aggregate not supported for column type {other:?}"))),
    }
}

/// The aggregate list of `compute_aggregate_values` (`llkv-executor/src/lib.rs:6087-6665`) -> `llkv_agg_spec[]` + the node
/// pool of their argument expressions.  `type_of` is the planner's type inference for an argument
/// (`llkv-plan/src/translation/schema.rs:80-88`; for a bare column, the column's type).  DISTINCT aggregates and
/// GROUP_CONCAT are not on this path: the recogniser has already turned such plans away.
pub fn flatten_aggregates(calls: &[(String, AggregateCall<FieldId>)], type_of: &dyn Fn(&ScalarExpr<FieldId>) -> Option<DataType>)
                          -> Result<(Vec<sys::llkv_agg_spec>, Vec<sys::llkv_scalar_node>)> {
    let mut specs = Vec::with_capacity(calls.len());
    let mut nodes = Vec::new();
    let mut spec = |kind: i32, root: i32, dt: Option<&DataType>, distinct: bool| -> Result<sys::llkv_agg_spec> {
        if distinct {
            return Err(Error::InvalidArgumentError("DISTINCT aggregates stay on the reference's path".into()));
        }
        let (t, p, s) = match dt {
            Some(d) => prim_type_of(d)?,
            None => (PT_INT64, 0, 0),
        };
        Ok(sys::llkv_agg_spec { kind, expr_root: root, data_type: t, precision: p, scale: s, distinct: 0, _pad: 0 })
    };
    for (_key, call) in calls {
        let typed = |name: &str, e: &ScalarExpr<FieldId>| -> Result<DataType> {
            let dt = type_of(e).ok_or_else(|| Error::Internal(format!("missing input type metadata for {name} aggregate")))?;
            normalise_aggregate_type(&dt, name)
        };
        let s = match call {
            AggregateCall::CountStar => spec(AGG_COUNT, -1, None, false)?,
            AggregateCall::Count { expr, distinct } => {
                let r = flatten_scalar(expr, &mut nodes)?;
                spec(AGG_COUNT, r, None, *distinct)?
            }
            AggregateCall::Sum { expr, distinct } => {
                let dt = typed("SUM", expr)?;
                let r = flatten_scalar(expr, &mut nodes)?;
                spec(AGG_SUM, r, Some(&dt), *distinct)?
            }
            AggregateCall::Total { expr, distinct } => {
                let dt = typed("TOTAL", expr)?;
                let r = flatten_scalar(expr, &mut nodes)?;
                spec(AGG_TOTAL, r, Some(&dt), *distinct)?
            }
            AggregateCall::Avg { expr, distinct } => {
                let dt = typed("AVG", expr)?;
                let r = flatten_scalar(expr, &mut nodes)?;
                spec(AGG_AVG, r, Some(&dt), *distinct)?
            }
            AggregateCall::Min(expr) => {
                let dt = typed("MIN", expr)?;
                let r = flatten_scalar(expr, &mut nodes)?;
                spec(AGG_MIN, r, Some(&dt), false)?
            }
            AggregateCall::Max(expr) => {
                let dt = typed("MAX", expr)?;
                let r = flatten_scalar(expr, &mut nodes)?;
                spec(AGG_MAX, r, Some(&dt), false)?
            }
            AggregateCall::CountNulls(expr) => {
                let r = flatten_scalar(expr, &mut nodes)?;
                spec(AGG_COUNT_NULLS, r, None, false)?
            }
            AggregateCall::GroupConcat { .. } => return Err(Error::InvalidArgumentError("GROUP_CONCAT stays on the reference's path".into())),
        };
        specs.push(s);
    }
    Ok((specs, nodes))
}
