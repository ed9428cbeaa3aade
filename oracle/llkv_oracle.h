/*
 * llkv_oracle.h — CPU restatement of LLKV's scan -> filter -> MVCC -> aggregate path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and only as
 * the checker / the timed CPU baseline.  The product path (rust-llkv_b200/csrc) never links or calls it.
 *
 * Parity pin: the reference (jzombie/rust-llkv v0.8.5-alpha, pure Rust) cannot be compiled in this image
 * (no cargo/rustc), so this is a restatement, pinned against the known answers the reference's own tests
 * hold for this path (tests/golden/ JSON files, each entry citing file:line).  Arithmetic that lives in the
 * un-vendored `arrow-arith`/`arrow-cast`/`arrow-ord` 57.1.0 crates (checked integer ops, Decimal128
 * mul/cast rounding, total-order float compare) is restated from their published semantics and has no
 * local golden vector: parity for those ops is UNPINNED.
 *
 * The oracle consumes the same flattened trees as the C ABI (include/llkv_gpu.h) and runs them the way
 * the reference does: one column scan per predicate leaf -> row-id bitmaps -> bitmap algebra -> per-row
 * MVCC rule -> gather in 65 536-row windows -> per-node temporaries -> scalar accumulator loops.
 */
#ifndef LLKV_ORACLE_H
#define LLKV_ORACLE_H

#include "../include/llkv_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_column {
  uint64_t field_id; /* FieldId within the table */
  int32_t type;      /* LLKV_PT_* */
  uint8_t precision;
  int8_t scale;
  uint8_t _pad[2];
  uint64_t n_rows;
  const void* values;      /* Arrow values buffer; UTF8: int32 offsets (n_rows+1) */
  const uint8_t* validity; /* optional Arrow validity bitmap, LSB first; NULL = all valid */
  const void* aux;         /* UTF8: data bytes */
} oracle_column;

typedef struct oracle_mvcc {
  const oracle_column* created_by; /* UInt64 */
  const oracle_column* deleted_by; /* UInt64 */
  uint64_t txn_id;
  uint64_t snapshot_id;
  const uint64_t* noncommitted; /* txn ids whose status is Active/Aborted */
  int32_t n_noncommitted;
} oracle_mvcc;

typedef struct oracle_program {
  const llkv_eval_op* ops;
  int32_t n_ops; /* 0 => trivially true filter */
  const llkv_literal* literals;
  int32_t n_literals;
  const llkv_scalar_node* nodes;
  int32_t n_nodes;
  const int32_t* list_roots;
  int32_t n_list_roots;
} oracle_program;

/* Selection bitmap (bit i = row row_begin+i) of rows passing program AND (optionally) MVCC. */
int32_t llkv_oracle_filter(const oracle_column* cols, int32_t n_cols, const oracle_program* prog,
                           const oracle_mvcc* mvcc, uint64_t row_begin, uint64_t row_end, int32_t n_threads,
                           uint64_t* out_words, uint64_t n_words, uint64_t* out_count, char* err, size_t errcap);

/* Filter -> MVCC -> gather -> evaluate -> accumulate -> finalize, ungrouped (n_keys==0) or GROUP BY. */
int32_t llkv_oracle_aggregate(const oracle_column* cols, int32_t n_cols, const oracle_program* prog,
                              const oracle_mvcc* mvcc, const llkv_agg_spec* specs, int32_t n_aggs,
                              const llkv_scalar_node* nodes, int32_t n_nodes, const uint64_t* key_fields,
                              int32_t n_keys, int32_t expr_mode, uint64_t row_begin, uint64_t row_end,
                              int32_t n_threads, llkv_agg_value* out_values, llkv_group_key* out_keys,
                              uint64_t group_capacity, uint64_t* out_groups, char* err, size_t errcap);

/* RowVersion::is_visible_for for one row (llkv-transaction/src/mvcc.rs:282-334). */
int32_t llkv_oracle_mvcc_visible(uint64_t created_by, uint64_t deleted_by, uint64_t txn_id, uint64_t snapshot_id,
                                 const uint64_t* noncommitted, int32_t n_noncommitted);

/* Chunk blob codec (llkv-column-map/src/serialization.rs:41-53,264-307). Returns bytes written / status. */
int64_t llkv_oracle_serialize_primitive(int32_t prim_type, uint8_t precision, int8_t scale, const void* values,
                                        uint64_t n_rows, uint8_t* out, uint64_t out_cap);

/* Exact decimal scalar ops (llkv-compute/src/scalar/decimal.rs:128-242). status 0 ok, else DecimalError. */
int32_t llkv_oracle_decimal_binary(int32_t op, const llkv_literal* a, const llkv_literal* b, llkv_literal* out);

#ifdef __cplusplus
}
#endif
#endif
