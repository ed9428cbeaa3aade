//! The executor branch: which `SelectPlan`s the GPU path takes, and how one is run.
//!
//! `QueryExecutor::execute_select_with_filter` (`llkv-executor/src/lib.rs:523-567`) asks `GpuPath::recognise` right after
//! it has ruled out compound selects and FROM-less selects; a plan that is turned away goes down the reference's own
//! branches untouched (there is no CPU fallback inside the GPU path: a plan it accepts either runs on the device or fails
//! with the reference's error).  Authored against jzombie/rust-llkv v0.8.5-alpha; not compiled here.
use std::sync::Arc;

use arrow::datatypes::DataType;
use llkv_executor::types::ExecutorTable;
use llkv_expr::expr::{AggregateCall, Expr, ScalarExpr};
use llkv_gpu_sys as sys;
use llkv_plan::plans::{AggregateExpr, AggregateFunction, SelectPlan};
use llkv_result::{Error, Result};
use llkv_storage::pager::Pager;
use llkv_transaction::mvcc::TransactionSnapshot;
use llkv_types::FieldId;
use simd_r_drive_entry_handle::EntryHandle;

use crate::flatten::{self, FlatProgram};
use crate::{Aggregation, Context, GroupKey, Program, ResidentTable};

/// `LLKV_EXPR_ARROW` / `LLKV_EXPR_EXACT`: ungrouped aggregates type their arguments with the arrow kernels
/// (`llkv-compute/src/eval.rs:565-614`), GROUP BY evaluates them per row with exact decimal arithmetic
/// (`llkv-executor/src/lib.rs:7229-7332`).  SURVEY.md section 8a notes D1 / D2.
pub const EXPR_ARROW: i32 = 0;
pub const EXPR_EXACT: i32 = 1;

/// A plan the device can run as one fused scan: filter program, aggregate list, GROUP BY keys.
pub struct GpuQuery {
    pub filter: Option<FlatProgram>,
    pub aggregates: Vec<(String, AggregateCall<FieldId>)>,
    pub specs: Vec<sys::llkv_agg_spec>,
    pub nodes: Vec<sys::llkv_scalar_node>,
    pub group_by: Vec<FieldId>,
    pub expr_mode: i32,
    pub cardinality_hint: u64,
}

pub struct GpuPath;

impl GpuPath {
    /// `None` = not ours (the reference's branches run it); `Some(Err)` = ours, and it fails the way the reference would
    /// (unknown column, aggregate over an unsupported type).
    pub fn recognise<P>(plan: &SelectPlan, table: &ExecutorTable<P>) -> Option<Result<GpuQuery>>
    where
        P: Pager<Blob = EntryHandle> + Send + Sync,
    {
        // one table, no joins, no set operations, no subqueries, nothing to order or page that the device does not do
        if plan.compound.is_some() || plan.tables.len() != 1 || !plan.joins.is_empty() || !plan.scalar_subqueries.is_empty() || plan.distinct {
            return None;
        }
        if plan.having.is_some() || plan.value_table_mode.is_some() {
            return None;
        }
        if plan.filter.as_ref().is_some_and(|f| !f.subqueries.is_empty()) {
            return None;
        }
        // aggregate-only outputs: `plan.aggregates` (execute_aggregates, lib.rs:5357-5682) — computed aggregates inside
        // projections (execute_computed_aggregates, :5686-5855) reach us as AggregateCalls through `from_calls`
        if plan.aggregates.is_empty() {
            return None;
        }
        let schema = table.schema.as_ref();
        let mut calls: Vec<(String, AggregateCall<FieldId>)> = Vec::with_capacity(plan.aggregates.len());
        for agg in &plan.aggregates {
            match agg {
                AggregateExpr::CountStar { alias, distinct } => {
                    if *distinct {
                        return None;
                    }
                    calls.push((alias.clone(), AggregateCall::CountStar));
                }
                AggregateExpr::Column { column, alias, function, distinct } => {
                    if *distinct {
                        return None;
                    }
                    let Some(col) = schema.column_by_name(column) else {
                        return Some(Err(Error::InvalidArgumentError(format!("unknown column '{column}' in aggregate"))));
                    };
                    let arg = Box::new(ScalarExpr::Column(col.field_id));
                    let call = match function {
                        AggregateFunction::Count => AggregateCall::Count { expr: arg, distinct: false },
                        AggregateFunction::SumInt64 => AggregateCall::Sum { expr: arg, distinct: false },
                        AggregateFunction::TotalInt64 => AggregateCall::Total { expr: arg, distinct: false },
                        AggregateFunction::MinInt64 => AggregateCall::Min(arg),
                        AggregateFunction::MaxInt64 => AggregateCall::Max(arg),
                        AggregateFunction::CountNulls => AggregateCall::CountNulls(arg),
                        AggregateFunction::GroupConcat => return None,
                    };
                    calls.push((alias.clone(), call));
                }
            }
        }
        // GROUP BY keys: integers, dates, booleans and short strings (`group_key_value`, lib.rs:9362-9456; anything else is
        // the reference's "GROUP BY does not support column type" error, which the device raises too)
        let mut group_by = Vec::with_capacity(plan.group_by.len());
        for name in &plan.group_by {
            let Some(col) = schema.column_by_name(name) else {
                return Some(Err(Error::InvalidArgumentError(format!("unknown column '{name}' in GROUP BY"))));
            };
            group_by.push(col.field_id);
        }
        // WHERE: names -> field ids exactly as the reference's branches do it (expression::translate_predicate, lib.rs:5477)
        let filter = match &plan.filter {
            None => None,
            Some(f) => {
                let translated: Expr<'static, FieldId> = match llkv_executor::translation::expression::translate_predicate(
                    f.predicate.clone(), schema, |name| Error::InvalidArgumentError(format!("unknown column '{name}' in filter"))) {
                    Ok(e) => e,
                    Err(e) => return Some(Err(e)),
                };
                if translated.is_trivially_true() {
                    None
                } else {
                    match flatten::flatten_expr(&translated) {
                        Ok(p) => Some(p),
                        // string patterns, struct literals, ...: not an error of the query, just not ours
                        Err(Error::PredicateBuild(_)) => return None,
                        Err(e) => return Some(Err(e)),
                    }
                }
            }
        };
        Some(Self::from_calls(calls, filter, group_by, &|e| match e {
            ScalarExpr::Column(fid) => schema.column_by_field_id(*fid).map(|c| c.data_type.clone()),
            other => llkv_plan::translation::schema::infer_computed_data_type(schema, other).ok(),
        }))
    }

    /// The shape `compute_aggregate_values` works from (`llkv-executor/src/lib.rs:6087-6103`): (result key, aggregate call).
    pub fn from_calls(calls: Vec<(String, AggregateCall<FieldId>)>, filter: Option<FlatProgram>, group_by: Vec<FieldId>,
                      type_of: &dyn Fn(&ScalarExpr<FieldId>) -> Option<DataType>) -> Result<GpuQuery> {
        let (specs, nodes) = flatten::flatten_aggregates(&calls, type_of)?;
        let expr_mode = if group_by.is_empty() { EXPR_ARROW } else { EXPR_EXACT };
        Ok(GpuQuery { filter, aggregates: calls, specs, nodes, group_by, expr_mode, cardinality_hint: 0 })
    }

    /// One query: new aggregate states, one fused scan of the resident table under the snapshot, finalize.  The result is
    /// what `AggregateAccumulator::finalize` would have produced per aggregate (and per group, in first-appearance order).
    pub fn execute(ctx: &Arc<Context>, table: &ResidentTable, query: &GpuQuery, snapshot: Option<(&TransactionSnapshot, &[u64])>)
                   -> Result<(Vec<sys::llkv_agg_value>, Vec<sys::llkv_group_key>)> {
        let program = match &query.filter {
            Some(f) => Some(Program::from_flat(ctx, &f.ops, &f.literals, &f.nodes, &f.list_roots)?),
            None => None,
        };
        if let Some((snap, noncommitted)) = snapshot {
            table.set_snapshot(snap.txn_id, snap.snapshot_id, noncommitted)?;
        }
        let keys: Vec<u64> = query.group_by.iter().map(|f| *f as u64).collect();
        let mut agg = Aggregation::new(ctx, table.table_id(), &query.specs, &query.nodes, &keys, query.expr_mode, query.cardinality_hint)?;
        agg.execute(program.as_ref(), snapshot.is_some(), 0, table.rows()?, true)?;
        agg.finalize()
    }

    /// The key tuple of every group as `GroupKeyValue`s (`llkv-executor/src/lib.rs:99-106`): dictionary-coded string keys are
    /// resolved through their column's dictionary.
    pub fn group_keys(table: &ResidentTable, query: &GpuQuery, keys: &[sys::llkv_group_key]) -> Result<Vec<Vec<GroupKey>>> {
        let nk = query.group_by.len();
        if nk == 0 {
            return Ok(Vec::new());
        }
        keys.chunks(nk)
            .map(|row| {
                row.iter()
                    .zip(&query.group_by)
                    .map(|(k, field)| table.column(*field as u32).ok_or(Error::NotFound)?.group_key_value(k))
                    .collect()
            })
            .collect()
    }
}
