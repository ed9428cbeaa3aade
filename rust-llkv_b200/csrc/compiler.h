// compiler.h — host-side plan compiler: llkv-expr predicate programs, scalar expressions and aggregate specs
// (as flattened by include/llkv_gpu.h) -> one Plan for the fused scan kernel (plan.h).
//
// Typing follows the reference, not SQL intuition:
//   * predicate leaves:   llkv-expr/src/typed_predicate.rs:75-167,252-315 + llkv-types/src/literal.rs:368-519
//   * program structure:  llkv-compute/src/program.rs:313-520, llkv-scan/src/predicate.rs:32-193 (rows/domain pairs)
//   * arrow-mode scalars: llkv-compute/src/eval.rs:30-38,71-148,565-750, kernels.rs:38-45,99-297, fast_numeric.rs:19-356
//   * exact-mode scalars: llkv-executor/src/lib.rs:7008-7440, llkv-compute/src/scalar/decimal.rs:30-242
//   * accumulators:       llkv-aggregate/src/lib.rs:400-449,463-748,759-1477
// Pure C++ (no CUDA): everything here runs at plan time on the host.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/llkv_gpu.h"
#include "plan.h"

namespace llkv {

// One registered column of the scanned table, as the compiler needs to see it.
struct ColumnMeta {
  uint64_t field_id = 0;  // FieldId within the table
  int32_t type = 0;       // LLKV_PT_*
  int precision = 0, scale = 0;
  bool nullable = false;  // a validity bitmap exists
  uint8_t load_kind = 0;  // LoadKind of the device representation
  uint32_t elem_bytes = 0;   // device bytes per row
  uint32_t arrow_bytes = 0;  // Arrow-layout value bytes per row (algorithmic traffic)
  const void* dev_values = nullptr;
  const unsigned char* dev_validity = nullptr;
  uint64_t n_rows = 0;
  // statistics gathered on the device while chunks are appended (the reference keeps min/max per chunk,
  // llkv-column-map/src/store/descriptor.rs:23-32)
  bool dec_fits_i64 = true;  // Decimal128: every value is a sign-extended i64
  uint64_t min_bits = 0, max_bits = 0;  // integers/dates/bool: order by the column's signedness
  bool has_minmax = false;
  uint8_t max_strlen = 0;  // Utf8
  bool str_non_ascii = false;  // Utf8: some string holds a byte >= 0x80
  // Utf8 column in dictionary form (strings longer than 7 bytes): the resident values are ranks in this byte-ordered list
  const std::vector<std::string>* dict_sorted = nullptr;
  uint64_t dict_epoch = 0;
};

// bytes of a string literal: inline (precision <= 15; lo and hi are adjacent, little endian) or by reference
inline void literal_bytes(const llkv_literal& l, const char** p, size_t* n) {
  if (l.precision == LLKV_LIT_STRING_BY_REF) {
    *p = reinterpret_cast<const char*>(static_cast<uintptr_t>(l.lo));
    *n = static_cast<size_t>(l.hi);
  } else {
    *p = reinterpret_cast<const char*>(&l.lo);
    *n = l.precision > 15 ? 15 : l.precision;
  }
}

struct ProgramView {
  const llkv_eval_op* ops = nullptr;
  int32_t n_ops = 0;
  const llkv_literal* literals = nullptr;
  int32_t n_literals = 0;
  const llkv_scalar_node* nodes = nullptr;
  int32_t n_nodes = 0;
  const int32_t* list_roots = nullptr;
  int32_t n_list_roots = 0;
};

struct MvccView {
  bool enabled = false;
  const ColumnMeta* created_by = nullptr;
  const ColumnMeta* deleted_by = nullptr;
  uint64_t txn_id = 0, snapshot_id = 0;
  std::vector<uint64_t> noncommitted;
};

// accumulator variants (llkv-aggregate/src/lib.rs:95-249), same order as the oracle uses
enum AccKind {
  ACC_COUNT_STAR, ACC_COUNT_COL, ACC_SUM_I64, ACC_SUM_F64, ACC_SUM_DEC, ACC_TOTAL_I64, ACC_TOTAL_F64, ACC_TOTAL_DEC,
  ACC_AVG_I64, ACC_AVG_F64, ACC_AVG_DEC, ACC_MIN_I64, ACC_MIN_F64, ACC_MIN_DEC, ACC_MAX_I64, ACC_MAX_F64, ACC_MAX_DEC,
  ACC_COUNT_NULLS
};

// Where one aggregate's state lives inside a group row (indices into gwords), for finalize.
struct AggLayout {
  int acc = 0;  // AccKind
  int precision = 0, scale = 0;
  int w_count = -1;       // non-NULL argument count (may alias word 0 when the argument cannot be NULL)
  int w_val = -1;         // first value word: sum limbs (2 or 4), f64 sum, or min/max (1 or 2)
  int n_limbs = 0;        // integer sums: 2 (i64 inputs) or 4 (i128 inputs)
  int w_first_valid = -1, w_first_nan = -1;  // MinFloat64/MaxFloat64 leading-NaN rule
  bool dead = false;      // argument is statically NULL-typed: the accumulator never sees a value
  // exact-mode computed argument of a Decimal128 accumulator: a group whose every value is NULL hands the accumulator
  // an all-NULL Int64 array (plan_values_to_arrow_array, llkv-executor/src/lib.rs:298-405) and the update fails
  bool all_null_group_is_error = false;
  // an error the reference raises from update() (type mismatch...), reported only if a row reaches the accumulator
  int raise_code = 0;
  std::string raise_message;
};

struct KeyLayout {
  uint64_t field_id = 0;
  int32_t type = 0;
  uint8_t kind = 0, bits = 0, strlen = 0, nullable = 0;
  uint64_t min = 0;
  bool is_signed = false;
  bool dict = false;  // Utf8 key whose field is a dictionary code
};

struct CompileRequest {
  std::vector<ColumnMeta> cols;  // every column of the table
  const ProgramView* prog = nullptr;  // nullptr / n_ops == 0: trivially true
  MvccView mvcc;
  const llkv_agg_spec* specs = nullptr;
  int32_t n_aggs = 0;
  const llkv_scalar_node* agg_nodes = nullptr;
  int32_t n_agg_nodes = 0;
  std::vector<uint64_t> key_fields;
  int32_t expr_mode = LLKV_EXPR_ARROW;
  bool bitmap_mode = false;
  bool force_wide = false;
  bool no_fast = false;  // keep only the general interpreter's program
};

struct CompileResult {
  Plan plan;  // program, literals, columns, MVCC, keys, accumulator layout.  Geometry / table pointers are the caller's.
  std::vector<AggLayout> aggs;
  std::vector<KeyLayout> keys;
  bool fast = false;             // plan.fcode holds a lean-kernel program (fast_kernel.cu)
  bool wide = false;             // compiled for the 128-bit interpreter
  bool can_narrow_fail = false;  // 64-bit interpreter may raise FLAG_NARROW_FAIL
  uint32_t algorithmic_bytes_per_row = 0, physical_bytes_per_row = 0;
  int32_t status = 0;
  std::string error;
};

// Returns LLKV_OK or an llkv_result::Error code with `out.error` set.
int32_t compile_plan(const CompileRequest& req, CompileResult& out);

// helpers shared with the ABI layer
int prim_type_width(int32_t type);  // Arrow value bytes, 0 = not a fixed-width type at this boundary
double powi_f64(double a, int b);   // Rust f64::powi lowering (compiler-rt __powidf2)

}  // namespace llkv
