//! Safe handles over `llkv-gpu-sys` in the vocabulary of the reference: what `INTEGRATION.md` section 3 calls `GpuPath`
//! is built from these.  Authored against jzombie/rust-llkv v0.8.5-alpha, not compiled here (no Rust toolchain in the
//! image); every call below is exercised through the same C ABI by the Python / C++ host mirrors in the test-suite.
//!
//! Ownership and threading follow `include/llkv_gpu.h`: handles are `Send`, calls on one handle are serialised by
//! `&mut self`, calls that share a context are serialised by the library, the library never calls back into Rust, and
//! page-locked host buffers handed to `append_*` stay alive until `seal` / `flush` returns (pageable ones only for the call).
//! Every child handle keeps its `Arc<Context>`: a context is destroyed after its last column, program and aggregate.
pub mod flatten;
pub mod path;
pub mod storage;

use std::ffi::c_void;
use std::ptr;
use std::sync::{Arc, Mutex};

use llkv_gpu_sys as sys;
use llkv_result::{Error, Result};
use llkv_storage::pager::{BatchGet, GetResult, Pager};
use llkv_storage::types::PhysicalKey;

/// Status codes are `llkv_result::Error` in variant order (`llkv-result/src/error.rs:31-176`).
fn check(rc: i32) -> Result<()> {
    if rc == 0 {
        return Ok(());
    }
    let mut buf = vec![0u8; 1024];
    let n = unsafe { sys::llkv_gpu_last_error(buf.as_mut_ptr().cast(), buf.len()) }.min(buf.len() - 1);
    let msg = String::from_utf8_lossy(&buf[..n]).into_owned();
    // LLKV_ERR_* (include/llkv_gpu.h) = the variant order of llkv_result::Error
    Err(match rc {
        1 => Error::Io(std::io::Error::other(msg)),
        2 => Error::Arrow(arrow::error::ArrowError::ComputeError(msg)),
        3 => Error::InvalidArgumentError(msg),
        4 => Error::NotFound,
        5 => Error::CatalogError(msg),
        6 => Error::ConstraintError(msg),
        7 => Error::TransactionContextError(msg),
        9 => Error::ExprCast(msg),
        10 => Error::PredicateBuild(msg),
        _ => Error::Internal(msg), // 8 and anything unknown
    })
}

/// One GPU: streams, pinned staging ring, the registry of resident columns.  There is no CPU fallback: creating a
/// context on a machine without a CUDA device is an error.
pub struct Context {
    raw: *mut sys::llkv_gpu_ctx,
}
// Every entry point of the library takes the context's own lock (include/llkv_gpu.h, "Conventions"): columns, programs and
// aggregates of one context may live on different threads.  Child handles hold an `Arc<Context>`, so the context is
// destroyed after the last of them.
unsafe impl Send for Context {}
unsafe impl Sync for Context {}

impl Context {
    pub fn new(device_ordinal: i32) -> Result<Arc<Self>> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::llkv_gpu_ctx_create(device_ordinal, 4, 64 << 20, &mut raw) })?;
        Ok(Arc::new(Self { raw }))
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { sys::llkv_gpu_ctx_destroy(self.raw) }
    }
}

/// A column resident in HBM (`ColumnStore` scan source, `llkv-column-map/src/store/scan/mod.rs:191`).
pub struct ResidentColumn {
    ctx: Arc<Context>,
    raw: *mut sys::llkv_gpu_column,
    logical_field_id: u64,
}
unsafe impl Send for ResidentColumn {}

impl ResidentColumn {
    /// The descriptor walk of `unsorted_visit` (`llkv-column-map/src/store/scan/unsorted.rs:202-241`): descriptor ->
    /// pages (one get per page) -> one batched get of every chunk -> `llkv_gpu_column_append_blob` per chunk -> seal.
    /// Tables are dense (row ids `0..n`, `dense_row_runs`, `scan/filter.rs:1510-1582`), so each chunk continues the last.
    pub fn load<P: Pager>(ctx: &Arc<Context>, pager: &P, descriptor_pk: PhysicalKey, logical_field_id: u64, prim_type: i32,
                          precision: u8, scale: i8) -> Result<Self> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::llkv_gpu_column_register(ctx.raw, logical_field_id, prim_type, precision, scale, &mut raw) })?;
        let col = Self { ctx: ctx.clone(), raw, logical_field_id };
        let get_one = |key: PhysicalKey| -> Result<P::Blob> {
            match pager.batch_get(&[BatchGet::Raw { key }])?.pop() {
                Some(GetResult::Raw { bytes, .. }) => Ok(bytes),
                _ => Err(Error::NotFound),
            }
        };
        let desc_blob = get_one(descriptor_pk)?;
        let mut desc = sys::llkv_column_descriptor::default();
        let b: &[u8] = desc_blob.as_ref();
        check(unsafe { sys::llkv_gpu_descriptor_parse(b.as_ptr().cast(), b.len() as u64, &mut desc) })?;
        check(unsafe { sys::llkv_gpu_column_reserve(col.raw, desc.total_row_count) })?;

        let mut metas: Vec<sys::llkv_chunk_metadata> = Vec::with_capacity(desc.total_chunk_count as usize);
        let mut page_pk = desc.head_page_pk;
        let mut page = vec![sys::llkv_chunk_metadata::default(); 256]; // DESCRIPTOR_ENTRIES_PER_PAGE
        while page_pk != 0 {
            let blob = get_one(page_pk)?;
            let b: &[u8] = blob.as_ref();
            let mut n = 0u64;
            check(unsafe { sys::llkv_gpu_descriptor_page_parse(b.as_ptr().cast(), b.len() as u64, &mut page_pk, page.as_mut_ptr(), 256, &mut n) })?;
            metas.extend(page[..n as usize].iter().filter(|m| m.row_count > 0));
        }
        let gets: Vec<BatchGet> = metas.iter().map(|m| BatchGet::Raw { key: m.chunk_pk }).collect();
        let mut row_id_base = 0u64;
        // the blobs (mmap-backed EntryHandles) stay alive until seal() returns: append_blob may still be reading them
        let blobs = pager.batch_get(&gets)?;
        for (meta, got) in metas.iter().zip(blobs.iter()) {
            let GetResult::Raw { bytes, .. } = got else { return Err(Error::NotFound) };
            let b: &[u8] = bytes.as_ref();
            check(unsafe { sys::llkv_gpu_column_append_blob(col.raw, meta.chunk_pk, b.as_ptr().cast(), b.len() as u64, ptr::null(), row_id_base) })?;
            row_id_base += meta.row_count;
        }
        check(unsafe { sys::llkv_gpu_column_seal(col.raw) })?;
        drop(blobs);
        Ok(col)
    }

    pub fn rows(&self) -> Result<u64> {
        let mut n = 0u64;
        check(unsafe { sys::llkv_gpu_column_rows(self.raw, &mut n) })?;
        Ok(n)
    }

    /// The field id part of the column's `LogicalFieldId` (`llkv-types/src/ids.rs:133-152`: the low 32 bits).
    pub fn field_id(&self) -> u32 {
        self.logical_field_id as u32
    }

    /// Entries of the column's dictionary (a Utf8 column that holds strings longer than 7 bytes), 0 otherwise.
    pub fn dict_size(&self) -> Result<u64> {
        let mut n = 0u64;
        check(unsafe { sys::llkv_gpu_column_dict_size(self.raw, &mut n) })?;
        Ok(n)
    }

    /// The string behind a dictionary code (`llkv_group_key.dict == 1`).
    pub fn dict_entry(&self, code: u64) -> Result<String> {
        let (mut p, mut n) = (ptr::null(), 0u64);
        check(unsafe { sys::llkv_gpu_column_dict_entry(self.raw, code, &mut p, &mut n) })?;
        let bytes = if n == 0 { &[][..] } else { unsafe { std::slice::from_raw_parts(p, n as usize) } };
        String::from_utf8(bytes.to_vec()).map_err(|e| Error::Internal(e.to_string()))
    }

    /// `GroupKeyValue` (`llkv-executor/src/lib.rs:99-106`) of one finalized key cell of this column.
    pub fn group_key_value(&self, k: &sys::llkv_group_key) -> Result<GroupKey> {
        if k.valid == 0 {
            return Ok(GroupKey::Null);
        }
        Ok(match k.type_ {
            flatten::PT_UTF8 if k.dict == 1 => GroupKey::String(self.dict_entry(k.bits)?),
            flatten::PT_UTF8 => {
                // packed short string: bytes big-endian from the top byte, length in the low byte
                let n = (k.bits & 0xff) as usize;
                let bytes: Vec<u8> = (0..n).map(|i| (k.bits >> (56 - 8 * i)) as u8).collect();
                GroupKey::String(String::from_utf8(bytes).map_err(|e| Error::Internal(e.to_string()))?)
            }
            flatten::PT_BOOLEAN => GroupKey::Bool(k.bits != 0),
            _ => GroupKey::Int(k.bits as i64),
        })
    }

    /// `SortIndexOps::stage_build_for_chunk` for every chunk, on the device; the blobs go to the pager under
    /// `ChunkMetadata.value_order_perm_pk` (`llkv-column-map/src/store/indexing/sort.rs:126-172`).
    pub fn sort_index_blobs(&self, chunk_rows: u64) -> Result<Vec<Vec<u8>>> {
        check(unsafe { sys::llkv_gpu_column_build_sort_index(self.raw, chunk_rows) })?;
        let mut out = Vec::new();
        for chunk in 0u64.. {
            let mut len = 0u64;
            match unsafe { sys::llkv_gpu_column_sort_index_blob(self.raw, chunk, ptr::null_mut(), 0, &mut len) } {
                0 => {}
                4 => break, // NotFound: past the last chunk
                rc => check(rc)?,
            }
            let mut blob = vec![0u8; len as usize];
            check(unsafe { sys::llkv_gpu_column_sort_index_blob(self.raw, chunk, blob.as_mut_ptr().cast(), len, &mut len) })?;
            out.push(blob);
        }
        Ok(out)
    }

    /// `ColumnStore::scan(field, ScanOptions, visitor)` (`llkv-column-map/src/store/scan/mod.rs:191-1080`): `on_run` gets
    /// each chunk's values (`None` = a null run) and row ids; sorted scans are sorted on the device.
    pub fn scan<F>(&self, anchor: Option<&ResidentColumn>, options: &sys::llkv_scan_options, chunk_rows: u64, mut on_run: F) -> Result<()>
    where
        F: FnMut(i32, Option<&[u8]>, Option<&[u64]>, u64),
    {
        unsafe extern "C" fn trampoline<F: FnMut(i32, Option<&[u8]>, Option<&[u64]>, u64)>(
            user: *mut c_void, prim_type: i32, values: *const c_void, row_ids: *const u64, n_rows: u64,
        ) -> i32 {
            let f = &mut *(user as *mut F);
            let width = flatten::prim_type_width(prim_type);
            let vals = if values.is_null() { None } else { Some(std::slice::from_raw_parts(values as *const u8, n_rows as usize * width)) };
            let ids = if row_ids.is_null() { None } else { Some(std::slice::from_raw_parts(row_ids, n_rows as usize)) };
            f(prim_type, vals, ids, n_rows);
            0
        }
        check(unsafe {
            sys::llkv_gpu_column_scan(self.raw, anchor.map_or(ptr::null_mut(), |a| a.raw), options, chunk_rows, Some(trampoline::<F>),
                                      (&mut on_run as *mut F).cast())
        })
    }
}

/// `GroupKeyValue` (`llkv-executor/src/lib.rs:99-106`).
#[derive(Clone, Debug, PartialEq, Eq, Hash)]
pub enum GroupKey {
    Null,
    Int(i64),
    Bool(bool),
    String(String),
}

impl Drop for ResidentColumn {
    fn drop(&mut self) {
        unsafe { sys::llkv_gpu_column_destroy(self.raw) };
        let _ = &self.ctx; // the context outlives its columns
    }
}

/// `ProgramCompiler::compile` output (`llkv-compute/src/program.rs:271-298`) on the device side of the boundary.  The
/// flattening of `EvalOp` / `OwnedFilter` / `ScalarExpr` into the `llkv_*` arrays is mechanical (children before parents;
/// `rust-llkv_b200/host/llkv_gpu.hpp: ProgramCompiler` is the same code in C++).
pub struct Program {
    raw: *mut sys::llkv_gpu_program,
    _ctx: Arc<Context>,
}
unsafe impl Send for Program {}

impl Program {
    pub fn from_flat(ctx: &Arc<Context>, ops: &[sys::llkv_eval_op], literals: &[sys::llkv_literal], nodes: &[sys::llkv_scalar_node],
                     list_roots: &[i32]) -> Result<Self> {
        let mut raw = ptr::null_mut();
        check(unsafe {
            sys::llkv_gpu_program_compile(ctx.raw, ops.as_ptr(), ops.len() as i32, literals.as_ptr(), literals.len() as i32, nodes.as_ptr(),
                                          nodes.len() as i32, list_roots.as_ptr(), list_roots.len() as i32, &mut raw)
        })?;
        Ok(Self { raw, _ctx: ctx.clone() })
    }
}
impl Drop for Program {
    fn drop(&mut self) {
        unsafe { sys::llkv_gpu_program_destroy(self.raw) }
    }
}

/// The resident image of one table: its user columns and, when the table carries them, the MVCC columns.  The snapshot
/// registered with the context refers to these columns, so it lives and dies with the table (it is cleared on drop).
pub struct ResidentTable {
    ctx: Arc<Context>,
    table_id: u64,
    columns: Vec<ResidentColumn>,
    mvcc: Option<(ResidentColumn, ResidentColumn)>,
    first_row_id: u64,
    snapshot_set: Mutex<bool>,
}

impl ResidentTable {
    pub fn new(ctx: &Arc<Context>, table_id: u64, columns: Vec<ResidentColumn>, mvcc: Option<(ResidentColumn, ResidentColumn)>, first_row_id: u64) -> Self {
        Self { ctx: ctx.clone(), table_id, columns, mvcc, first_row_id, snapshot_set: Mutex::new(false) }
    }
    pub fn table_id(&self) -> u64 {
        self.table_id
    }
    pub fn first_row_id(&self) -> u64 {
        self.first_row_id
    }
    pub fn rows(&self) -> Result<u64> {
        self.columns.first().map_or(Ok(0), |c| c.rows())
    }
    /// The resident column of a field (group keys of dictionary-coded columns are resolved through it).
    pub fn column(&self, field_id: u32) -> Option<&ResidentColumn> {
        self.columns.iter().find(|c| c.field_id() == field_id)
    }
    /// `MvccRowIdFilter::new(txn_manager, snapshot)` (`llkv-transaction/src/helpers.rs:259-312`): the snapshot plus every
    /// transaction id whose `TxnIdManager::status` is Active or Aborted.  Without MVCC columns every row is visible
    /// (`helpers.rs:141-152`).
    pub fn set_snapshot(&self, txn_id: u64, snapshot_id: u64, noncommitted: &[u64]) -> Result<()> {
        let Some((created, deleted)) = &self.mvcc else { return Ok(()) };
        let mut set = self.snapshot_set.lock().unwrap();
        check(unsafe {
            sys::llkv_gpu_mvcc_set(self.ctx.raw, self.table_id, created.raw, deleted.raw, txn_id, snapshot_id, noncommitted.as_ptr(), noncommitted.len() as i32)
        })?;
        *set = true;
        Ok(())
    }
    /// Selection bitmap over positions (bit i = row `first_row_id + row_begin + i`): `ScanStorage::filter_leaf` /
    /// `RowIdFilter::filter` on the device.
    pub fn filter_bitmap(&self, program: Option<&Program>, apply_mvcc: bool, row_begin: u64, row_end: u64) -> Result<Vec<u64>> {
        let n_words = ((row_end - row_begin + 63) / 64) as usize;
        let mut words = vec![0u64; n_words.max(1)];
        let mut count = 0u64;
        check(unsafe {
            sys::llkv_gpu_filter_bitmap(self.ctx.raw, self.table_id, program.map_or(ptr::null(), |p| p.raw.cast_const()), apply_mvcc as i32,
                                        row_begin, row_end, words.as_mut_ptr(), n_words as u64, &mut count)
        })?;
        words.truncate(n_words);
        Ok(words)
    }
}
impl Drop for ResidentTable {
    fn drop(&mut self) {
        if *self.snapshot_set.lock().unwrap() {
            unsafe { sys::llkv_gpu_mvcc_clear(self.ctx.raw, self.table_id) };
        }
    }
}

/// A set of `AggregateState`s fused with the scan that feeds them: one call per query instead of one
/// `AggregateAccumulator::update` per 65 536-row batch (`llkv-aggregate/src/lib.rs:463,759,1488`;
/// `llkv-scan/src/execute.rs:47-295`).
pub struct Aggregation {
    raw: *mut sys::llkv_gpu_agg,
    n_aggs: usize,
    n_keys: usize,
    _ctx: Arc<Context>,
}
unsafe impl Send for Aggregation {}

impl Aggregation {
    pub fn new(ctx: &Arc<Context>, table_id: u64, specs: &[sys::llkv_agg_spec], nodes: &[sys::llkv_scalar_node], group_key_fields: &[u64],
               expr_mode: i32, cardinality_hint: u64) -> Result<Self> {
        let mut raw = ptr::null_mut();
        check(unsafe {
            sys::llkv_gpu_agg_create(ctx.raw, table_id, specs.as_ptr(), specs.len() as i32, nodes.as_ptr(), nodes.len() as i32,
                                     group_key_fields.as_ptr(), group_key_fields.len() as i32, expr_mode, cardinality_hint, &mut raw)
        })?;
        Ok(Self { raw, n_aggs: specs.len(), n_keys: group_key_fields.len(), _ctx: ctx.clone() })
    }

    /// reset + run + (merge across the context's peers) in one call; replayed as one CUDA graph once the step repeats.
    pub fn execute(&mut self, filter: Option<&Program>, apply_mvcc: bool, row_begin: u64, row_end: u64, merge: bool) -> Result<()> {
        check(unsafe {
            sys::llkv_gpu_agg_execute(self.raw, filter.map_or(ptr::null(), |p| p.raw.cast_const()), apply_mvcc as i32, row_begin, row_end, merge as i32)
        })
    }

    /// Scans rows `[row_begin, row_end)`; asynchronous — `finalize` settles the run (and reruns wider / with a larger group
    /// table if the device asked for it).
    pub fn run(&mut self, filter: Option<&Program>, apply_mvcc: bool, row_begin: u64, row_end: u64) -> Result<()> {
        check(unsafe { sys::llkv_gpu_agg_run(self.raw, filter.map_or(ptr::null(), |p| p.raw.cast_const()), apply_mvcc as i32, row_begin, row_end) })
    }

    /// `AggregateAccumulator::finalize` for every aggregate of every group, groups in first-appearance order
    /// (`llkv-executor/src/lib.rs:5064-5089`): (values, keys), `n_aggs` / `n_keys` entries per group.
    pub fn finalize(&mut self) -> Result<(Vec<sys::llkv_agg_value>, Vec<sys::llkv_group_key>)> {
        let mut groups = 0u64;
        check(unsafe { sys::llkv_gpu_agg_group_count(self.raw, &mut groups) })?;
        let g = groups as usize;
        let mut values: Vec<sys::llkv_agg_value> = Vec::with_capacity(g * self.n_aggs);
        let mut keys: Vec<sys::llkv_group_key> = Vec::with_capacity(g * self.n_keys);
        check(unsafe { sys::llkv_gpu_agg_finalize(self.raw, values.as_mut_ptr(), keys.as_mut_ptr(), groups, &mut groups) })?;
        unsafe {
            values.set_len(groups as usize * self.n_aggs);
            keys.set_len(groups as usize * self.n_keys);
        }
        Ok((values, keys))
    }
}
impl Drop for Aggregation {
    fn drop(&mut self) {
        unsafe { sys::llkv_gpu_agg_destroy(self.raw) }
    }
}

/// Page-locks the pager's mapping for the lifetime of the guard so chunk appends DMA straight out of it.
pub struct RegisteredMapping(*const c_void);
impl RegisteredMapping {
    /// # Safety
    /// `bytes` must stay mapped until the guard is dropped.
    pub unsafe fn new(bytes: &[u8]) -> Result<Self> {
        check(sys::llkv_gpu_host_register(bytes.as_ptr().cast(), bytes.len() as u64))?;
        Ok(Self(bytes.as_ptr().cast()))
    }
}
impl Drop for RegisteredMapping {
    fn drop(&mut self) {
        unsafe { sys::llkv_gpu_host_unregister(self.0) };
    }
}
