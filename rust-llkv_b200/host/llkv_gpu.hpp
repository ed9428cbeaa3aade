// llkv_gpu.hpp — C++17 host side above the C ABI (include/llkv_gpu.h), header only.
//
// The reference's host language (Rust) has no toolchain in this image, so this is the compiled-language mirror of the
// reference's own vocabulary for this path, same names and argument meaning, flattening into the C ABI structs:
//   Literal                       llkv-types/src/literal.rs:26-41 (+ `impl From<T> for Literal`, :47-90)
//   Bound / Operator / Filter     llkv-expr/src/expr.rs:367-402, std::ops::Bound
//   ScalarExpr / BinaryOp / CompareOp   llkv-expr/src/expr.rs:127-182,311-349
//   Expr                          llkv-expr/src/expr.rs:16-43
//   ProgramCompiler               llkv-compute/src/program.rs:271-439 (postfix EvalOp program; gather_fused :415-439)
//   AggregateKind / AggregateSpec llkv-aggregate/src/lib.rs:26-69
// and RAII owners of the ABI's opaque handles (Context, Column, Program, Aggregation).  Errors are llkv::Error with the
// llkv_result::Error code (llkv-result/src/error.rs:31-176) and the library's message.  Nothing here evaluates anything:
// every row is processed by the CUDA kernels behind the ABI.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/llkv_gpu.h"

namespace llkv {

struct Error : std::runtime_error {
  int32_t code;
  Error(int32_t c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int32_t rc) {
  if (rc == LLKV_OK) return;
  char buf[1024];
  llkv_gpu_last_error(buf, sizeof(buf));
  throw Error(rc, buf);
}

// ------------------------------------------------------------------------------------------------ Literal
struct Literal {
  llkv_literal c{};
  std::shared_ptr<std::string> long_string;  // bytes behind a by-reference string literal (> 15 bytes)
  static Literal Null() { return Literal(); }
  static Literal Int128(__int128 v) {
    Literal l;
    l.c.kind = LLKV_LIT_INT128;
    l.c.lo = (uint64_t)(unsigned __int128)v;
    l.c.hi = (uint64_t)((unsigned __int128)v >> 64);
    l.c.precision = digits(v);
    return l;
  }
  static Literal Float64(double v) {
    Literal l;
    l.c.kind = LLKV_LIT_FLOAT64;
    std::memcpy(&l.c.lo, &v, 8);
    return l;
  }
  // DecimalValue::new(raw, scale) (llkv-types/src/decimal.rs:67-76)
  static Literal Decimal128(__int128 raw, int scale) {
    if (digits(raw) > 38) throw std::invalid_argument("DecimalError::PrecisionOverflow");
    Literal l = Int128(raw);
    l.c.kind = LLKV_LIT_DECIMAL128;
    l.c.scale = (int8_t)scale;
    return l;
  }
  static Literal String(const std::string& s) {
    Literal l;
    l.c.kind = LLKV_LIT_STRING;
    if (s.size() > 15) {  // by reference: the library copies the bytes in the call that receives the literal
      l.long_string = std::make_shared<std::string>(s);
      l.c.lo = (uint64_t)(uintptr_t)l.long_string->data();
      l.c.hi = l.long_string->size();
      l.c.precision = LLKV_LIT_STRING_BY_REF;
      return l;
    }
    unsigned char b[16] = {0};
    std::memcpy(b, s.data(), s.size());
    std::memcpy(&l.c.lo, b, 8);
    std::memcpy(&l.c.hi, b + 8, 8);
    l.c.precision = (uint8_t)s.size();
    return l;
  }
  static Literal Boolean(bool v) {
    Literal l;
    l.c.kind = LLKV_LIT_BOOLEAN;
    l.c.lo = v ? 1 : 0;
    return l;
  }
  static Literal Date32(int32_t days) {
    Literal l;
    l.c.kind = LLKV_LIT_DATE32;
    l.c.lo = (uint64_t)(int64_t)days;
    return l;
  }
  // impl From<T> for Literal
  Literal() { c.kind = LLKV_LIT_NULL; }
  Literal(int v) { *this = Int128(v); }
  Literal(long v) { *this = Int128(v); }
  Literal(long long v) { *this = Int128(v); }
  Literal(double v) { *this = Float64(v); }
  Literal(bool v) { *this = Boolean(v); }
  Literal(const char* s) { *this = String(s); }

 private:
  static uint8_t digits(__int128 v) {
    unsigned __int128 a = v < 0 ? (unsigned __int128)0 - (unsigned __int128)v : (unsigned __int128)v;
    uint8_t n = 1;
    while (a >= 10) {
      a /= 10;
      ++n;
    }
    return n;
  }
};

// ------------------------------------------------------------------------------------------------ Bound / Operator / Filter
struct Bound {
  int32_t kind = LLKV_BOUND_UNBOUNDED;
  Literal value;
  static Bound Included(Literal v) { return Bound{LLKV_BOUND_INCLUDED, v}; }
  static Bound Excluded(Literal v) { return Bound{LLKV_BOUND_EXCLUDED, v}; }
  static Bound Unbounded() { return Bound{}; }
};

struct Operator {
  int32_t tag = LLKV_OP_EQUALS;
  std::vector<Literal> literals;
  Bound lower, upper;
  bool case_sensitive = true;  // StartsWith / EndsWith / Contains
  static Operator Equals(Literal v) { return Operator{LLKV_OP_EQUALS, {v}, {}, {}}; }
  static Operator Range(Bound lower, Bound upper) {
    Operator o{LLKV_OP_RANGE, {}, lower, upper};
    if (lower.kind != LLKV_BOUND_UNBOUNDED) o.literals.push_back(lower.value);
    if (upper.kind != LLKV_BOUND_UNBOUNDED) o.literals.push_back(upper.value);
    return o;
  }
  static Operator GreaterThan(Literal v) { return Operator{LLKV_OP_GT, {v}, {}, {}}; }
  static Operator GreaterThanOrEquals(Literal v) { return Operator{LLKV_OP_GTE, {v}, {}, {}}; }
  static Operator LessThan(Literal v) { return Operator{LLKV_OP_LT, {v}, {}, {}}; }
  static Operator LessThanOrEquals(Literal v) { return Operator{LLKV_OP_LTE, {v}, {}, {}}; }
  static Operator In(std::vector<Literal> values) { return Operator{LLKV_OP_IN, std::move(values), {}, {}}; }
  // Operator::{StartsWith, EndsWith, Contains} { pattern, case_sensitive } (llkv-expr/src/expr.rs:388-399)
  static Operator StartsWith(const std::string& pattern, bool case_sensitive = true) { return pattern_op(LLKV_OP_STARTS_WITH, pattern, case_sensitive); }
  static Operator EndsWith(const std::string& pattern, bool case_sensitive = true) { return pattern_op(LLKV_OP_ENDS_WITH, pattern, case_sensitive); }
  static Operator Contains(const std::string& pattern, bool case_sensitive = true) { return pattern_op(LLKV_OP_CONTAINS, pattern, case_sensitive); }
  static Operator pattern_op(int32_t tag, const std::string& pattern, bool case_sensitive) {
    Operator o{tag, {Literal::String(pattern)}, {}, {}};
    o.case_sensitive = case_sensitive;
    return o;
  }
  static Operator IsNull() { return Operator{LLKV_OP_IS_NULL, {}, {}, {}}; }
  static Operator IsNotNull() { return Operator{LLKV_OP_IS_NOT_NULL, {}, {}, {}}; }
};

struct Filter {
  uint64_t field_id = 0;
  Operator op;
};

// ------------------------------------------------------------------------------------------------ ScalarExpr
enum class BinaryOp : int32_t { Add = LLKV_BIN_ADD, Subtract = LLKV_BIN_SUB, Multiply = LLKV_BIN_MUL, Divide = LLKV_BIN_DIV, Modulo = LLKV_BIN_MOD };
enum class CompareOp : int32_t { Eq = LLKV_CMP_EQ, NotEq = LLKV_CMP_NE, Lt = LLKV_CMP_LT, LtEq = LLKV_CMP_LE, Gt = LLKV_CMP_GT, GtEq = LLKV_CMP_GE };

struct DataType {
  int32_t type = LLKV_PT_INT64;
  uint8_t precision = 0;
  int8_t scale = 0;
  static DataType Int64() { return {LLKV_PT_INT64, 0, 0}; }
  static DataType Float64() { return {LLKV_PT_FLOAT64, 0, 0}; }
  static DataType Date32() { return {LLKV_PT_DATE32, 0, 0}; }
  static DataType UInt64() { return {LLKV_PT_UINT64, 0, 0}; }
  static DataType Utf8() { return {LLKV_PT_UTF8, 0, 0}; }
  static DataType Decimal128(int p, int s) { return {LLKV_PT_DECIMAL128, (uint8_t)p, (int8_t)s}; }
};

struct ScalarExpr {
  struct Node {
    int32_t tag = LLKV_SE_COLUMN;
    uint64_t field_id = 0;
    Literal literal;
    int32_t op = 0;
    std::shared_ptr<const Node> left, right;
    DataType data_type;
  };
  std::shared_ptr<const Node> n;

  static ScalarExpr Column(uint64_t fid) {
    Node x;
    x.tag = LLKV_SE_COLUMN;
    x.field_id = fid;
    return wrap(x);
  }
  static ScalarExpr Lit(Literal v) {
    Node x;
    x.tag = LLKV_SE_LITERAL;
    x.literal = v;
    return wrap(x);
  }
  static ScalarExpr Binary(const ScalarExpr& l, BinaryOp op, const ScalarExpr& r) {
    Node x;
    x.tag = LLKV_SE_BINARY;
    x.op = (int32_t)op;
    x.left = l.n;
    x.right = r.n;
    return wrap(x);
  }
  static ScalarExpr Compare(const ScalarExpr& l, CompareOp op, const ScalarExpr& r) {
    Node x;
    x.tag = LLKV_SE_COMPARE;
    x.op = (int32_t)op;
    x.left = l.n;
    x.right = r.n;
    return wrap(x);
  }
  static ScalarExpr Cast(const ScalarExpr& e, DataType t) {
    Node x;
    x.tag = LLKV_SE_CAST;
    x.left = e.n;
    x.data_type = t;
    return wrap(x);
  }
  static ScalarExpr IsNull(const ScalarExpr& e, bool negated = false) {
    Node x;
    x.tag = LLKV_SE_IS_NULL;
    x.op = negated ? 1 : 0;
    x.left = e.n;
    return wrap(x);
  }
  ScalarExpr operator+(const ScalarExpr& o) const { return Binary(*this, BinaryOp::Add, o); }
  ScalarExpr operator-(const ScalarExpr& o) const { return Binary(*this, BinaryOp::Subtract, o); }
  ScalarExpr operator*(const ScalarExpr& o) const { return Binary(*this, BinaryOp::Multiply, o); }
  ScalarExpr operator/(const ScalarExpr& o) const { return Binary(*this, BinaryOp::Divide, o); }

 private:
  static ScalarExpr wrap(const Node& x) {
    ScalarExpr e;
    e.n = std::make_shared<const Node>(x);
    return e;
  }
};

// Flattens ScalarExpr trees into one llkv_scalar_node array, children before parents.
struct NodePool {
  std::vector<llkv_scalar_node> nodes;
  int32_t add(const ScalarExpr& e) { return add(*e.n); }

 private:
  int32_t add(const ScalarExpr::Node& e) {
    llkv_scalar_node n;
    std::memset(&n, 0, sizeof(n));
    n.tag = e.tag;
    n.left = n.right = -1;
    n.literal.kind = LLKV_LIT_NULL;
    switch (e.tag) {
      case LLKV_SE_COLUMN: n.field_id = e.field_id; break;
      case LLKV_SE_LITERAL: n.literal = e.literal.c; break;
      case LLKV_SE_BINARY: case LLKV_SE_COMPARE:
        n.op = e.op;
        n.left = add(*e.left);
        n.right = add(*e.right);
        break;
      case LLKV_SE_CAST:
        n.left = add(*e.left);
        n.cast_type = e.data_type.type;
        n.cast_precision = e.data_type.precision;
        n.cast_scale = e.data_type.scale;
        break;
      case LLKV_SE_IS_NULL: case LLKV_SE_NOT:
        n.op = e.op;
        n.left = add(*e.left);
        break;
      default: throw std::invalid_argument("ScalarExpr variant does not cross this boundary");
    }
    nodes.push_back(n);
    return (int32_t)nodes.size() - 1;
  }
};

// ------------------------------------------------------------------------------------------------ Expr (predicate tree)
struct Expr {
  enum Tag { kAnd, kOr, kNot, kPred, kCompare, kInList, kIsNull, kLiteral } tag = kLiteral;
  std::vector<Expr> children;
  Filter filter;
  ScalarExpr left, right;
  CompareOp op = CompareOp::Eq;
  std::vector<ScalarExpr> list;
  bool negated = false;
  bool value = false;

  static Expr And(std::vector<Expr> c) { Expr e; e.tag = kAnd; e.children = std::move(c); return e; }
  static Expr Or(std::vector<Expr> c) { Expr e; e.tag = kOr; e.children = std::move(c); return e; }
  static Expr Not(Expr inner) { Expr e; e.tag = kNot; e.children.push_back(std::move(inner)); return e; }
  static Expr Pred(Filter f) { Expr e; e.tag = kPred; e.filter = std::move(f); return e; }
  static Expr Compare(ScalarExpr l, CompareOp op, ScalarExpr r) { Expr e; e.tag = kCompare; e.left = std::move(l); e.op = op; e.right = std::move(r); return e; }
  static Expr InList(ScalarExpr x, std::vector<ScalarExpr> items, bool negated = false) {
    Expr e; e.tag = kInList; e.left = std::move(x); e.list = std::move(items); e.negated = negated; return e;
  }
  static Expr IsNull(ScalarExpr x, bool negated = false) { Expr e; e.tag = kIsNull; e.left = std::move(x); e.negated = negated; return e; }
  static Expr Literal(bool v) { Expr e; e.tag = kLiteral; e.value = v; return e; }
};
inline Expr pred(uint64_t field_id, Operator op) { return Expr::Pred(Filter{field_id, std::move(op)}); }

struct CompiledProgram {
  std::vector<llkv_eval_op> ops;
  std::vector<llkv_literal> literals;
  NodePool pool;
  std::vector<int32_t> list_roots;
};

// compile_eval (llkv-compute/src/program.rs:313-413) + gather_fused (:415-439)
class ProgramCompiler {
 public:
  explicit ProgramCompiler(Expr root) : root_(std::move(root)) {}
  CompiledProgram compile() const {
    CompiledProgram p;
    emit(root_, p);
    return p;
  }

 private:
  Expr root_;
  static llkv_eval_op blank(int32_t tag) {
    llkv_eval_op op;
    std::memset(&op, 0, sizeof(op));
    op.tag = tag;
    return op;
  }
  static bool gather_fused(const std::vector<Expr>& children, uint64_t* fid) {
    if (children.empty()) return false;
    bool have = false;
    for (const Expr& c : children) {
      if (c.tag != Expr::kPred) return false;
      if (!have) { *fid = c.filter.field_id; have = true; }
      else if (*fid != c.filter.field_id) return false;
    }
    return true;
  }
  static llkv_eval_op filter_op(int32_t tag, const Filter& f, CompiledProgram& p) {
    llkv_eval_op op = blank(tag);
    op.operator_tag = f.op.tag;
    op.field_id = f.field_id;
    op.lower_kind = f.op.lower.kind;
    op.upper_kind = f.op.upper.kind;
    op.lit_begin = (int32_t)p.literals.size();
    op.lit_count = (int32_t)f.op.literals.size();
    op.literal_bool = f.op.case_sensitive ? 0 : 1;
    for (const Literal& l : f.op.literals) p.literals.push_back(l.c);  // (by-reference strings stay owned by the Expr tree)
    return op;
  }
  static void emit(const Expr& node, CompiledProgram& p) {
    switch (node.tag) {
      case Expr::kAnd: {
        if (node.children.empty()) throw std::invalid_argument("AND expression requires at least one predicate");
        uint64_t fid = 0;
        if (gather_fused(node.children, &fid)) {
          llkv_eval_op op = blank(LLKV_EV_FUSED_AND);
          op.field_id = fid;
          op.child_count = (int32_t)node.children.size();
          p.ops.push_back(op);
          for (const Expr& c : node.children) p.ops.push_back(filter_op(LLKV_EV_FILTER_ITEM, c.filter, p));
          return;
        }
        for (const Expr& c : node.children) emit(c, p);
        llkv_eval_op op = blank(LLKV_EV_AND);
        op.child_count = (int32_t)node.children.size();
        p.ops.push_back(op);
        return;
      }
      case Expr::kOr: {
        if (node.children.empty()) throw std::invalid_argument("OR expression requires at least one predicate");
        for (const Expr& c : node.children) emit(c, p);
        llkv_eval_op op = blank(LLKV_EV_OR);
        op.child_count = (int32_t)node.children.size();
        p.ops.push_back(op);
        return;
      }
      case Expr::kNot:
        emit(node.children.at(0), p);
        p.ops.push_back(blank(LLKV_EV_NOT));
        return;
      case Expr::kPred: p.ops.push_back(filter_op(LLKV_EV_PUSH_PREDICATE, node.filter, p)); return;
      case Expr::kCompare: {
        llkv_eval_op op = blank(LLKV_EV_PUSH_COMPARE);
        op.expr_left = p.pool.add(node.left);
        op.expr_right = p.pool.add(node.right);
        op.cmp_op = (int32_t)node.op;
        p.ops.push_back(op);
        return;
      }
      case Expr::kInList: {
        llkv_eval_op op = blank(LLKV_EV_PUSH_IN_LIST);
        op.expr_left = p.pool.add(node.left);
        op.expr_right = (int32_t)p.list_roots.size();
        op.child_count = (int32_t)node.list.size();
        op.negated = node.negated ? 1 : 0;
        for (const ScalarExpr& item : node.list) p.list_roots.push_back(p.pool.add(item));
        p.ops.push_back(op);
        return;
      }
      case Expr::kIsNull: {
        llkv_eval_op op = blank(LLKV_EV_PUSH_IS_NULL);
        op.expr_left = p.pool.add(node.left);
        op.negated = node.negated ? 1 : 0;
        p.ops.push_back(op);
        return;
      }
      case Expr::kLiteral: {
        llkv_eval_op op = blank(LLKV_EV_PUSH_LITERAL);
        op.literal_bool = node.value ? 1 : 0;
        p.ops.push_back(op);
        return;
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------ aggregates
struct AggregateKind {
  int32_t kind = LLKV_AGG_COUNT;
  bool has_expr = false;
  ScalarExpr expr;
  DataType data_type;
  bool distinct = false;
  static AggregateKind CountStar() { return AggregateKind{}; }
  static AggregateKind Count(ScalarExpr e) { return make(LLKV_AGG_COUNT, std::move(e), DataType::Int64()); }
  static AggregateKind Sum(ScalarExpr e, DataType t) { return make(LLKV_AGG_SUM, std::move(e), t); }
  static AggregateKind Total(ScalarExpr e, DataType t) { return make(LLKV_AGG_TOTAL, std::move(e), t); }
  static AggregateKind Avg(ScalarExpr e, DataType t) { return make(LLKV_AGG_AVG, std::move(e), t); }
  static AggregateKind Min(ScalarExpr e, DataType t) { return make(LLKV_AGG_MIN, std::move(e), t); }
  static AggregateKind Max(ScalarExpr e, DataType t) { return make(LLKV_AGG_MAX, std::move(e), t); }
  static AggregateKind CountNulls(ScalarExpr e) { return make(LLKV_AGG_COUNT_NULLS, std::move(e), DataType::Int64()); }

 private:
  static AggregateKind make(int32_t k, ScalarExpr e, DataType t) {
    AggregateKind a;
    a.kind = k;
    a.has_expr = true;
    a.expr = std::move(e);
    a.data_type = t;
    return a;
  }
};
struct AggregateSpec {
  std::string alias;
  AggregateKind kind;
};
struct FlatAggregates {
  std::vector<llkv_agg_spec> specs;
  NodePool pool;
};
inline FlatAggregates flatten_aggregates(const std::vector<AggregateSpec>& specs) {
  FlatAggregates out;
  for (const AggregateSpec& s : specs) {
    llkv_agg_spec a;
    std::memset(&a, 0, sizeof(a));
    a.kind = s.kind.kind;
    a.expr_root = s.kind.has_expr ? out.pool.add(s.kind.expr) : -1;
    a.data_type = s.kind.data_type.type;
    a.precision = s.kind.data_type.precision;
    a.scale = s.kind.data_type.scale;
    a.distinct = s.kind.distinct ? 1 : 0;
    out.specs.push_back(a);
  }
  return out;
}

// LogicalFieldId packing (llkv-types/src/ids.rs:133-152): namespace << 48 | table << 32 | field
inline uint64_t logical_field_id(uint64_t table_id, uint64_t field_id, uint64_t ns = 0) { return (ns << 48) | ((table_id & 0xffff) << 32) | (field_id & 0xffffffffull); }

// ------------------------------------------------------------------------------------------------ column metadata
// The descriptor walk of unsorted_visit (llkv-column-map/src/store/scan/unsorted.rs:202-241): `batch_get` is the pager
// (key -> blob); returns the ChunkMetadata of every chunk the scan has to read, in scan order.  `lower` / `upper` prune
// with IntRanges::matches (store/pruning.rs:104-258).
template <class BatchGet>  // std::string-like blob = batch_get(uint64_t physical_key)
inline std::vector<llkv_chunk_metadata> walk_descriptor(BatchGet&& batch_get, uint64_t descriptor_pk, llkv_column_descriptor* descriptor_out = nullptr,
                                                        int32_t prim_type = 0, const llkv_range_bound* lower = nullptr,
                                                        const llkv_range_bound* upper = nullptr) {
  const auto desc_blob = batch_get(descriptor_pk);
  llkv_column_descriptor desc;
  check(llkv_gpu_descriptor_parse(desc_blob.data(), desc_blob.size(), &desc));
  if (descriptor_out) *descriptor_out = desc;
  std::vector<llkv_chunk_metadata> metas;
  llkv_chunk_metadata page[256];  // DESCRIPTOR_ENTRIES_PER_PAGE (store/constants.rs:14)
  for (uint64_t pk = desc.head_page_pk; pk;) {
    const auto blob = batch_get(pk);
    uint64_t n = 0;
    check(llkv_gpu_descriptor_page_parse(blob.data(), blob.size(), &pk, page, 256, &n));
    for (uint64_t i = 0; i < n; ++i)
      if (page[i].row_count > 0 && llkv_gpu_chunk_overlaps(prim_type, page[i].min_val_u64, page[i].max_val_u64, lower, upper)) metas.push_back(page[i]);
  }
  return metas;
}

// ------------------------------------------------------------------------------------------------ RAII owners of the ABI handles
class Context {
 public:
  explicit Context(int device = 0, int n_streams = 4, uint64_t pinned_bytes = 64ull << 20) { check(llkv_gpu_ctx_create(device, n_streams, pinned_bytes, &h_)); }
  ~Context() { if (h_) llkv_gpu_ctx_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  llkv_gpu_ctx* get() const { return h_; }
  void synchronize() { check(llkv_gpu_ctx_synchronize(h_)); }
  void set_timing(bool on) { check(llkv_gpu_ctx_set_timing(h_, on ? 1 : 0)); }
  void set_jit(int mode) { check(llkv_gpu_ctx_set_jit(h_, mode)); }
  void set_partitioning(int mode) { check(llkv_gpu_ctx_set_partitioning(h_, mode)); }
  void set_pruning(int mode) { check(llkv_gpu_ctx_set_pruning(h_, mode)); }
  // MvccRowIdFilter::new(txn_manager, snapshot) (llkv-transaction/src/helpers.rs:259-312)
  void set_snapshot(uint64_t table_id, llkv_gpu_column* created_by, llkv_gpu_column* deleted_by, uint64_t txn_id, uint64_t snapshot_id,
                    const std::vector<uint64_t>& noncommitted = {}) {
    check(llkv_gpu_mvcc_set(h_, table_id, created_by, deleted_by, txn_id, snapshot_id, noncommitted.data(), (int32_t)noncommitted.size()));
  }

 private:
  llkv_gpu_ctx* h_ = nullptr;
};

class Column {
 public:
  Column(Context& ctx, uint64_t lfid, DataType t) { check(llkv_gpu_column_register(ctx.get(), lfid, t.type, t.precision, t.scale, &h_)); }
  ~Column() { if (h_) llkv_gpu_column_destroy(h_); }
  Column(const Column&) = delete;
  Column& operator=(const Column&) = delete;
  llkv_gpu_column* get() const { return h_; }
  // one chunk of Arrow values (dense row ids continuing the column)
  void append(const void* values, uint64_t n_rows, uint64_t row_id_base, const uint8_t* validity = nullptr) {
    check(llkv_gpu_column_append_chunk(h_, next_pk_++, values, n_rows, validity, nullptr, row_id_base, nullptr));
  }
  // a serialized chunk blob exactly as the pager hands it out (serialization.rs:41-53)
  void append_blob(const void* blob, uint64_t len, uint64_t row_id_base) { check(llkv_gpu_column_append_blob(h_, next_pk_++, blob, len, nullptr, row_id_base)); }
  void seal() { check(llkv_gpu_column_seal(h_)); }
  uint64_t rows() const {
    uint64_t n = 0;
    check(llkv_gpu_column_rows(h_, &n));
    return n;
  }
  // ColumnStore::append with the row-id shadow column: rows land at (row id - first_row_id), an existing row id is overwritten
  void append_rows(const void* values, const uint64_t* row_ids, uint64_t n_rows, uint64_t first_row_id = 0, const uint8_t* validity = nullptr) {
    check(llkv_gpu_column_append_chunk(h_, next_pk_++, values, n_rows, validity, row_ids, first_row_id, nullptr));
  }
  void delete_rows(const std::vector<uint64_t>& row_ids) { check(llkv_gpu_column_delete_rows(h_, row_ids.data(), row_ids.size())); }
  // page-locked sources may be reused after this returns
  void flush() { check(llkv_gpu_column_flush(h_)); }
  // the string behind a group key whose `dict` flag is set (columns holding strings longer than 7 bytes)
  std::string dict_entry(uint64_t code) const {
    const uint8_t* p = nullptr;
    uint64_t n = 0;
    check(llkv_gpu_column_dict_entry(h_, code, &p, &n));
    return std::string(reinterpret_cast<const char*>(p), (size_t)n);
  }
  // ColumnStore::scan(field, ScanOptions, visitor): `visit(prim_type, values or nullptr (a null run), row ids or nullptr, rows)`
  template <typename F>
  void scan(const llkv_scan_options& options, F&& visit, const Column* anchor = nullptr, uint64_t chunk_rows = 0) {
    auto tramp = [](void* user, int32_t prim_type, const void* values, const uint64_t* row_ids, uint64_t n) -> int32_t {
      (*static_cast<typename std::remove_reference<F>::type*>(user))(prim_type, values, row_ids, n);
      return 0;
    };
    check(llkv_gpu_column_scan(h_, anchor ? anchor->get() : nullptr, &options, chunk_rows, tramp, &visit));
  }
  // per-chunk value_order_perm blobs (SortIndexOps::stage_build_for_chunk), built on the device
  std::vector<std::vector<uint8_t>> sort_index_blobs(uint64_t chunk_rows = 0) {
    check(llkv_gpu_column_build_sort_index(h_, chunk_rows));
    std::vector<std::vector<uint8_t>> out;
    for (uint64_t chunk = 0;; ++chunk) {
      uint64_t len = 0;
      const int32_t rc = llkv_gpu_column_sort_index_blob(h_, chunk, nullptr, 0, &len);
      if (rc == LLKV_ERR_NOT_FOUND) break;
      check(rc);
      out.emplace_back((size_t)len);
      check(llkv_gpu_column_sort_index_blob(h_, chunk, out.back().data(), len, &len));
    }
    return out;
  }

 private:
  llkv_gpu_column* h_ = nullptr;
  uint64_t next_pk_ = 1;
};

class Program {
 public:
  Program(Context& ctx, const Expr& filter) {
    CompiledProgram p = ProgramCompiler(filter).compile();
    check(llkv_gpu_program_compile(ctx.get(), p.ops.data(), (int32_t)p.ops.size(), p.literals.data(), (int32_t)p.literals.size(), p.pool.nodes.data(),
                                   (int32_t)p.pool.nodes.size(), p.list_roots.data(), (int32_t)p.list_roots.size(), &h_));
  }
  ~Program() { if (h_) llkv_gpu_program_destroy(h_); }
  Program(const Program&) = delete;
  Program& operator=(const Program&) = delete;
  const llkv_gpu_program* get() const { return h_; }

 private:
  llkv_gpu_program* h_ = nullptr;
};

struct GroupRow {
  std::vector<llkv_group_key> keys;
  std::vector<llkv_agg_value> values;
};

// A set of AggregateStates fused with the scan that feeds them.
class Aggregation {
 public:
  Aggregation(Context& ctx, uint64_t table_id, const std::vector<AggregateSpec>& specs, const std::vector<uint64_t>& group_by = {},
              uint64_t cardinality_hint = 0)
      : n_aggs_(specs.size()), n_keys_(group_by.size()) {
    FlatAggregates f = flatten_aggregates(specs);
    // GROUP BY expressions run in exact decimal mode, ungrouped ones through the arrow kernels
    const int32_t mode = group_by.empty() ? LLKV_EXPR_ARROW : LLKV_EXPR_EXACT;
    check(llkv_gpu_agg_create(ctx.get(), table_id, f.specs.data(), (int32_t)f.specs.size(), f.pool.nodes.data(), (int32_t)f.pool.nodes.size(),
                              group_by.data(), (int32_t)group_by.size(), mode, cardinality_hint, &h_));
  }
  ~Aggregation() { if (h_) llkv_gpu_agg_destroy(h_); }
  Aggregation(const Aggregation&) = delete;
  Aggregation& operator=(const Aggregation&) = delete;
  void reset() { check(llkv_gpu_agg_reset(h_)); }
  void run(const Program* prog, bool apply_mvcc, uint64_t row_begin, uint64_t row_end) {
    check(llkv_gpu_agg_run(h_, prog ? prog->get() : nullptr, apply_mvcc ? 1 : 0, row_begin, row_end));
  }
  void merge() { check(llkv_gpu_agg_merge(h_)); }
  // one step: new states + scan + (multi-GPU) merge; replayed as a CUDA graph once it repeats unchanged
  void execute(const Program* prog, bool apply_mvcc, uint64_t row_begin, uint64_t row_end, bool merge = true) {
    check(llkv_gpu_agg_execute(h_, prog ? prog->get() : nullptr, apply_mvcc ? 1 : 0, row_begin, row_end, merge ? 1 : 0));
  }
  // HAVING / ORDER BY / OFFSET / LIMIT over the group rows (llkv-executor/src/lib.rs:5256-5355), applied at finalize
  void set_output(const std::vector<llkv_having_term>& having, const std::vector<llkv_order_key>& order_by, uint64_t offset = 0, uint64_t limit = 0) {
    check(llkv_gpu_agg_set_output(h_, having.data(), (int32_t)having.size(), order_by.data(), (int32_t)order_by.size(), offset, limit));
  }
  std::vector<GroupRow> finalize(uint64_t group_capacity = 1) {
    std::vector<llkv_agg_value> vals(group_capacity * (n_aggs_ ? n_aggs_ : 1));
    std::vector<llkv_group_key> keys(group_capacity * (n_keys_ ? n_keys_ : 1));
    uint64_t n = 0;
    check(llkv_gpu_agg_finalize(h_, vals.data(), keys.data(), group_capacity, &n));
    std::vector<GroupRow> out(n);
    for (uint64_t g = 0; g < n; ++g) {
      out[g].keys.assign(keys.begin() + g * n_keys_, keys.begin() + (g + 1) * n_keys_);
      out[g].values.assign(vals.begin() + g * n_aggs_, vals.begin() + (g + 1) * n_aggs_);
    }
    return out;
  }
  llkv_run_info run_info() const {
    llkv_run_info info;
    check(llkv_gpu_agg_run_info(h_, &info));
    return info;
  }

 private:
  llkv_gpu_agg* h_ = nullptr;
  size_t n_aggs_, n_keys_;
};

}  // namespace llkv
