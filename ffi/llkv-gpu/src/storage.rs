//! `GpuStorageTable`: the `StorageTable<P>` swap point of the executor (`llkv-executor/src/types/storage.rs:20-50`,
//! `ExecutorTable.storage`, `types/executor_types.rs:26-56`).
//!
//! It wraps the table's ordinary adapter and a `ResidentTable` (its columns in HBM).  Row-id predicates
//! (`filter_row_ids`) run on the device; `scan_stream` hands batches to a caller closure, so nothing can be fused behind it
//! and it stays with the wrapped adapter — the fused aggregate path enters through `GpuPath` instead (path.rs), which finds
//! the resident image by downcasting `ExecutorTable.storage` with `as_any`.  Authored, not compiled here.
use std::any::Any;
use std::sync::Arc;

use arrow::array::RecordBatch;
use croaring::Treemap;
use llkv_executor::types::StorageTable;
use llkv_expr::Expr;
use llkv_join::{JoinKey, JoinOptions};
use llkv_result::Result as LlkvResult;
use llkv_storage::pager::Pager;
use llkv_table::table::{ScanProjection, ScanStreamOptions};
use llkv_types::{FieldId, TableId};
use simd_r_drive_entry_handle::EntryHandle;

use crate::flatten;
use crate::{Context, Program, ResidentTable};

pub struct GpuStorageTable<P>
where
    P: Pager<Blob = EntryHandle> + Send + Sync,
{
    inner: Arc<dyn StorageTable<P>>,
    ctx: Arc<Context>,
    resident: Arc<ResidentTable>,
}

impl<P> GpuStorageTable<P>
where
    P: Pager<Blob = EntryHandle> + Send + Sync,
{
    /// `inner` is the adapter the executor built (`TableStorageAdapter`); `resident` was loaded from the same table's
    /// descriptors through the same pager (`ResidentTable::load`).
    pub fn new(inner: Arc<dyn StorageTable<P>>, ctx: Arc<Context>, resident: Arc<ResidentTable>) -> Self {
        Self { inner, ctx, resident }
    }

    pub fn resident(&self) -> &Arc<ResidentTable> {
        &self.resident
    }

    pub fn context(&self) -> &Arc<Context> {
        &self.ctx
    }
}

impl<P> StorageTable<P> for GpuStorageTable<P>
where
    P: Pager<Blob = EntryHandle> + Send + Sync + 'static,
{
    fn table_id(&self) -> TableId {
        self.inner.table_id()
    }

    fn scan_stream<'expr>(&self, projections: &[ScanProjection], filter_expr: &Expr<'expr, FieldId>, options: ScanStreamOptions<P>,
                          on_batch: &mut dyn FnMut(RecordBatch)) -> LlkvResult<()> {
        // RecordBatch streams are the reference's business: the device path produces aggregate states and selection
        // bitmaps, never row batches
        self.inner.scan_stream(projections, filter_expr, options, on_batch)
    }

    /// `Table::filter_row_ids`: the predicate program runs on the device over the resident columns; the selection bitmap
    /// over positions becomes a `Treemap` of row ids (position + the table's first row id).
    fn filter_row_ids<'expr>(&self, filter_expr: &Expr<'expr, FieldId>) -> LlkvResult<Treemap> {
        let flat = match flatten::flatten_expr(filter_expr) {
            Ok(f) => f,
            Err(_) => return self.inner.filter_row_ids(filter_expr), // an operator this path does not carry (string patterns)
        };
        let program = Program::from_flat(&self.ctx, &flat.ops, &flat.literals, &flat.nodes, &flat.list_roots)?;
        let rows = self.resident.rows()?;
        let words = self.resident.filter_bitmap(Some(&program), false, 0, rows)?;
        let base = self.resident.first_row_id();
        let mut out = Treemap::new();
        for (w, word) in words.iter().enumerate() {
            let mut bits = *word;
            while bits != 0 {
                let b = bits.trailing_zeros() as u64;
                out.add(base + (w as u64) * 64 + b);
                bits &= bits - 1;
            }
        }
        Ok(out)
    }

    fn join_stream(&self, right: &dyn StorageTable<P>, keys: &[JoinKey], options: &JoinOptions, on_batch: &mut dyn FnMut(RecordBatch))
                   -> LlkvResult<()> {
        self.inner.join_stream(right, keys, options, on_batch)
    }

    fn as_any(&self) -> &dyn Any {
        self
    }
}
