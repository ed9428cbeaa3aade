/*
 * llkv_oracle.c — CPU restatement of LLKV's scan -> filter -> MVCC -> aggregate path (see llkv_oracle.h).
 * TEST INFRASTRUCTURE ONLY; never linked into the product.  Pinned by tests/golden/: the reference's own known answers
 * (reference_known_answers.json) and, for the arithmetic that lives in the arrow crates (not vendored in the reference tree,
 * no network here), known answers derived from arrow-rs' published rules and cross-checked with Python decimal and pyarrow
 * (arrow_known_answers.json + make_arrow_golden.py).  Still unpinned: the float total order of arrow-ord and the
 * integer/float casts of arrow-cast (no vector anywhere).
 *
 * Shape follows the reference, not the GPU design:
 *   leaf scan per predicate  -> Vec<u64> row ids -> bitmap           llkv-column-map/src/store/scan/filter.rs:931-958
 *   bitmap AND/OR/NOT(domain)                                       llkv-scan/src/predicate.rs:32-193
 *   per-row MVCC rule                                               llkv-transaction/src/helpers.rs:178-245
 *   gather in 65 536-row windows                                    llkv-scan/src/execute.rs:31,268-292
 *   per-node temporaries, arrow-arith semantics                     llkv-compute/src/eval.rs:565-750, kernels.rs:99-297
 *   scalar accumulator loops, error at first prefix overflow        llkv-aggregate/src/lib.rs:759-1477
 *   GROUP BY: key map in first-appearance order, exact decimal ops   llkv-executor/src/lib.rs:5028-5355,7229-7332
 */
#define _GNU_SOURCE
#include "llkv_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef __int128 i128;
typedef unsigned __int128 u128;

#define ROW_STREAM_CHUNK_SIZE 65536 /* llkv-scan/src/execute.rs:31 */
#define TXN_ID_NONE UINT64_MAX     /* llkv-transaction/src/mvcc.rs:25-31 */
#define TXN_ID_AUTO_COMMIT 1ull

/* ------------------------------------------------------------------ errors */
typedef struct {
  char* buf;
  size_t cap;
} Err;

static int32_t fail(Err* e, int32_t code, const char* fmt, ...) {
  if (e && e->buf && e->cap) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(e->buf, e->cap, fmt, ap);
    va_end(ap);
  }
  return code;
}

/* ------------------------------------------------------------------ small helpers */
static inline i128 lit_i128(const llkv_literal* l) { return (i128)(((u128)l->hi << 64) | (u128)l->lo); }
static inline double lit_f64(const llkv_literal* l) {
  double d;
  memcpy(&d, &l->lo, 8);
  return d;
}

static i128 pow10_i128(int k) {
  i128 v = 1;
  for (int i = 0; i < k; ++i) v *= 10;
  return v;
}
/* compiler-rt __powidf2: what Rust's f64::powi lowers to (llkv-aggregate/src/lib.rs:424, llkv-types/src/decimal.rs:104) */
static double powi_f64(double a, int b) {
  const int recip = b < 0;
  double r = 1;
  while (1) {
    if (b & 1) r *= a;
    b /= 2;
    if (b == 0) break;
    a *= a;
  }
  return recip ? 1 / r : r;
}
static int digits_i128(i128 v) { /* digit_count of |v|, 0 -> 1 (llkv-types/src/decimal.rs digit_count_i256) */
  u128 a = v < 0 ? (u128)0 - (u128)v : (u128)v;
  int d = 1;
  while (a >= 10) {
    a /= 10;
    ++d;
  }
  return d;
}
static int fits_precision(i128 v, int p) { /* arrow is_valid_decimal_precision: |v| <= 10^p - 1 */
  if (p >= 39) return 1;
  i128 lim = pow10_i128(p);
  return v < lim && v > -lim;
}

static const oracle_column* find_col(const oracle_column* cols, int32_t n, uint64_t fid) {
  for (int32_t i = 0; i < n; ++i)
    if (cols[i].field_id == fid) return &cols[i];
  return NULL;
}
static inline int col_valid(const oracle_column* c, uint64_t i) {
  return !c->validity || ((c->validity[i >> 3] >> (i & 7)) & 1);
}
static int type_is_signed_int(int t) {
  return t == LLKV_PT_INT8 || t == LLKV_PT_INT16 || t == LLKV_PT_INT32 || t == LLKV_PT_INT64 || t == LLKV_PT_DATE32 ||
         t == LLKV_PT_DATE64;
}
static int type_is_unsigned_int(int t) {
  return t == LLKV_PT_UINT8 || t == LLKV_PT_UINT16 || t == LLKV_PT_UINT32 || t == LLKV_PT_UINT64;
}
static int type_bits(int t) {
  switch (t) {
    case LLKV_PT_INT8: case LLKV_PT_UINT8: return 8;
    case LLKV_PT_INT16: case LLKV_PT_UINT16: return 16;
    case LLKV_PT_INT32: case LLKV_PT_UINT32: case LLKV_PT_DATE32: return 32;
    default: return 64;
  }
}
static inline int64_t load_sint(const oracle_column* c, uint64_t i) {
  switch (c->type) {
    case LLKV_PT_INT8: return ((const int8_t*)c->values)[i];
    case LLKV_PT_INT16: return ((const int16_t*)c->values)[i];
    case LLKV_PT_INT32: case LLKV_PT_DATE32: return ((const int32_t*)c->values)[i];
    default: return ((const int64_t*)c->values)[i];
  }
}
static inline uint64_t load_uint(const oracle_column* c, uint64_t i) {
  switch (c->type) {
    case LLKV_PT_UINT8: case LLKV_PT_BOOLEAN: return ((const uint8_t*)c->values)[i];
    case LLKV_PT_UINT16: return ((const uint16_t*)c->values)[i];
    case LLKV_PT_UINT32: return ((const uint32_t*)c->values)[i];
    default: return ((const uint64_t*)c->values)[i];
  }
}
static inline i128 load_dec(const oracle_column* c, uint64_t i) {
  i128 v;
  memcpy(&v, (const char*)c->values + 16 * i, 16);
  return v;
}
/* short string -> order-preserving u64: bytes big-endian from the top byte, length in the low byte */
static int pack_short_string(const uint8_t* p, uint32_t len, uint64_t* out) {
  if (len > 7) return 0;
  uint64_t k = 0;
  for (uint32_t i = 0; i < len; ++i) k |= (uint64_t)p[i] << (56 - 8 * i);
  *out = k | len;
  return 1;
}
static int load_str(const oracle_column* c, uint64_t i, uint64_t* out) {
  const int32_t* off = (const int32_t*)c->values;
  return pack_short_string((const uint8_t*)c->aux + off[i], (uint32_t)(off[i + 1] - off[i]), out);
}

/* ------------------------------------------------------------------ bitmaps (stand in for croaring::Treemap) */
typedef struct {
  uint64_t* w;
  uint64_t nbits, nwords;
} Bits;
static Bits bits_new(uint64_t nbits) {
  Bits b;
  b.nbits = nbits;
  b.nwords = (nbits + 63) / 64;
  b.w = (uint64_t*)calloc(b.nwords ? b.nwords : 1, 8);
  return b;
}
static void bits_free(Bits* b) {
  free(b->w);
  b->w = NULL;
}
static void bits_fill(Bits* b) {
  memset(b->w, 0xff, b->nwords * 8);
  if (b->nbits & 63) b->w[b->nwords - 1] = (~0ull) >> (64 - (b->nbits & 63));
}
static inline void bits_set(Bits* b, uint64_t i) { b->w[i >> 6] |= 1ull << (i & 63); }
static inline int bits_get(const Bits* b, uint64_t i) { return (b->w[i >> 6] >> (i & 63)) & 1; }
static uint64_t bits_count(const Bits* b) {
  uint64_t c = 0;
  for (uint64_t i = 0; i < b->nwords; ++i) c += (uint64_t)__builtin_popcountll(b->w[i]);
  return c;
}
static void bits_and(Bits* a, const Bits* b) { for (uint64_t i = 0; i < a->nwords; ++i) a->w[i] &= b->w[i]; }
static void bits_or(Bits* a, const Bits* b) { for (uint64_t i = 0; i < a->nwords; ++i) a->w[i] |= b->w[i]; }
static void bits_andnot_from(Bits* dst, const Bits* dom, const Bits* t) { /* dst = dom - t */
  for (uint64_t i = 0; i < dst->nwords; ++i) dst->w[i] = dom->w[i] & ~t->w[i];
}
static Bits bits_clone(const Bits* a) {
  Bits b = bits_new(a->nbits);
  memcpy(b.w, a->w, a->nwords * 8);
  return b;
}

/* ------------------------------------------------------------------ typed predicates
 * llkv-expr/src/typed_predicate.rs:41-167 (matches), :252-315 (build), llkv-types/src/literal.rs:368-519 (casts) */
enum { DOM_I64, DOM_U64, DOM_F64, DOM_F32, DOM_DEC, DOM_STR, DOM_BOOL };
typedef union {
  int64_t i;
  uint64_t u;
  double f;
  i128 d;
  struct { const uint8_t* p; uint64_t n; } s; /* DOM_STR: the string's bytes (any length) */
} PVal;
typedef struct {
  int op; /* LLKV_OP_*; RANGE with both unbounded is Predicate::All */
  int dom;
  int lower_kind, upper_kind;
  PVal a, b;
  PVal* in;
  int n_in;
  int dec_scale_col;  /* DOM_DEC: scale of the column */
  int dec_scale_lit;  /* DOM_DEC: scale the literals were aligned to (>= col scale) */
  uint8_t pat[16];    /* STARTS_WITH / ENDS_WITH / CONTAINS: the pattern's inline bytes */
  const uint8_t* pat_long;
  uint8_t* pat_ptr;   /* the pattern (owned copy; lower-cased when ci) */
  int pat_len, ci;
  uint8_t* strbuf;    /* DOM_STR: the bytes of the string literals a / b / in[] point into */
  int n_str;
} TPred;

static int col_domain(const oracle_column* c) {
  if (c->type == LLKV_PT_FLOAT64) return DOM_F64;
  if (c->type == LLKV_PT_FLOAT32) return DOM_F32;
  if (c->type == LLKV_PT_DECIMAL128) return DOM_DEC;
  if (c->type == LLKV_PT_UTF8) return DOM_STR;
  if (c->type == LLKV_PT_BOOLEAN) return DOM_BOOL;
  if (type_is_unsigned_int(c->type)) return DOM_U64;
  return DOM_I64;
}

/* FromLiteral for the column's native type. Returns 0 or LLKV_ERR_PREDICATE_BUILD. */
static int32_t lit_to_native(const oracle_column* c, const llkv_literal* l, PVal* out, int* dec_scale, uint8_t* str_store, Err* e) {
  int dom = col_domain(c);
  switch (dom) {
    case DOM_I64:
    case DOM_U64: {
      i128 v;
      if (l->kind == LLKV_LIT_INT128) v = lit_i128(l);
      else if (l->kind == LLKV_LIT_DECIMAL128 && l->scale == 0) v = lit_i128(l);
      else if (l->kind == LLKV_LIT_DATE32 && (c->type == LLKV_PT_DATE32))
        v = (int64_t)l->lo; /* extension D4: Date32 literal against a Date32 column (reference: TypeMismatch) */
      else
        return fail(e, LLKV_ERR_PREDICATE_BUILD, "literal type mismatch: expected integer");
      int bits = type_bits(c->type);
      if (dom == DOM_I64) {
        i128 lo = -((i128)1 << (bits - 1)), hi = ((i128)1 << (bits - 1)) - 1;
        if (v < lo || v > hi) return fail(e, LLKV_ERR_PREDICATE_BUILD, "literal out of range for %d-bit integer", bits);
        out->i = (int64_t)v;
      } else {
        i128 hi = bits == 64 ? (i128)UINT64_MAX : (((i128)1 << bits) - 1);
        if (v < 0 || v > hi) return fail(e, LLKV_ERR_PREDICATE_BUILD, "literal out of range for unsigned %d-bit", bits);
        out->u = (uint64_t)v;
      }
      return 0;
    }
    case DOM_F64:
    case DOM_F32: {
      double v;
      if (l->kind == LLKV_LIT_FLOAT64) v = lit_f64(l);
      else if (l->kind == LLKV_LIT_INT128) v = (double)lit_i128(l);
      else if (l->kind == LLKV_LIT_DECIMAL128) {
        i128 raw = lit_i128(l);
        v = raw == 0 ? 0.0 : (double)raw / powi_f64(10.0, l->scale);
      } else
        return fail(e, LLKV_ERR_PREDICATE_BUILD, "literal type mismatch: expected float");
      if (dom == DOM_F32) {
        float f = (float)v;
        if (!isfinite(f)) return fail(e, LLKV_ERR_PREDICATE_BUILD, "float literal out of range for f32");
        v = (double)f;
      }
      out->f = v;
      return 0;
    }
    case DOM_DEC: { /* extension D4: i128 compare after aligning scales (llkv-types/src/decimal.rs:171-191) */
      i128 raw;
      int ls;
      if (l->kind == LLKV_LIT_INT128) { raw = lit_i128(l); ls = 0; }
      else if (l->kind == LLKV_LIT_DECIMAL128) { raw = lit_i128(l); ls = l->scale; }
      else return fail(e, LLKV_ERR_PREDICATE_BUILD, "literal type mismatch: expected decimal");
      int target = ls > c->scale ? ls : c->scale;
      if (*dec_scale >= 0 && *dec_scale != target) {
        /* all literals of one predicate are aligned to one scale: rescale to the larger */
        if (target < *dec_scale) target = *dec_scale;
      }
      i128 f = pow10_i128(target - ls), r;
      if (__builtin_mul_overflow(raw, f, &r)) return fail(e, LLKV_ERR_PREDICATE_BUILD, "decimal literal overflow");
      out->d = r;
      *dec_scale = target;
      return 0;
    }
    case DOM_BOOL: {
      if (l->kind == LLKV_LIT_BOOLEAN) out->u = l->lo != 0;
      else if (l->kind == LLKV_LIT_INT128 && (lit_i128(l) == 0 || lit_i128(l) == 1)) out->u = (uint64_t)lit_i128(l);
      else return fail(e, LLKV_ERR_PREDICATE_BUILD, "literal type mismatch: expected boolean");
      return 0;
    }
    case DOM_STR: {
      /* String::from_literal: the literal's bytes; the column side compares whole strings (str: Ord is byte-wise) */
      if (l->kind != LLKV_LIT_STRING) return fail(e, LLKV_ERR_PREDICATE_BUILD, "literal type mismatch: expected string");
      if (l->precision == LLKV_LIT_STRING_BY_REF) { /* by reference: the caller's bytes outlive this call */
        out->s.p = (const uint8_t*)(uintptr_t)l->lo;
        out->s.n = l->hi;
        return 0;
      }
      memcpy(str_store, &l->lo, 8);
      memcpy(str_store + 8, &l->hi, 8);
      out->s.p = str_store;
      out->s.n = l->precision > 15 ? 15 : l->precision;
      return 0;
    }
  }
  return fail(e, LLKV_ERR_INTERNAL, "bad domain");
}

static int32_t build_pred(const oracle_column* c, const llkv_eval_op* op, const llkv_literal* lits, TPred* p, Err* e) {
  memset(p, 0, sizeof(*p));
  p->op = op->operator_tag;
  p->dom = col_domain(c);
  p->lower_kind = p->upper_kind = LLKV_BOUND_UNBOUNDED;
  p->dec_scale_col = c->scale;
  int ds = -1;
  int32_t rc;
  const llkv_literal* l = lits + op->lit_begin;
  p->strbuf = (uint8_t*)calloc((size_t)(op->lit_count > 2 ? op->lit_count : 2), 16);
  switch (op->operator_tag) {
    case LLKV_OP_EQUALS: case LLKV_OP_GT: case LLKV_OP_GTE: case LLKV_OP_LT: case LLKV_OP_LTE:
      if (op->lit_count != 1) return fail(e, LLKV_ERR_INTERNAL, "operator needs one literal");
      if ((rc = lit_to_native(c, l, &p->a, &ds, p->strbuf, e))) return rc;
      break;
    case LLKV_OP_RANGE: {
      int k = 0;
      p->lower_kind = op->lower_kind;
      p->upper_kind = op->upper_kind;
      /* align both decimal bounds to one scale */
      if (p->dom == DOM_DEC) {
        for (int i = 0; i < op->lit_count; ++i) {
          int s = l[i].kind == LLKV_LIT_DECIMAL128 ? l[i].scale : 0;
          if (s > ds) ds = s;
        }
        if (c->scale > ds) ds = c->scale;
      }
      if (op->lower_kind != LLKV_BOUND_UNBOUNDED) { if ((rc = lit_to_native(c, l + k, &p->a, &ds, p->strbuf, e))) return rc; ++k; }
      if (op->upper_kind != LLKV_BOUND_UNBOUNDED) { if ((rc = lit_to_native(c, l + k, &p->b, &ds, p->strbuf + 16, e))) return rc; ++k; }
      break;
    }
    case LLKV_OP_IN:
      p->n_in = op->lit_count;
      p->in = (PVal*)calloc((size_t)(op->lit_count ? op->lit_count : 1), sizeof(PVal));
      if (p->dom == DOM_DEC) {
        for (int i = 0; i < op->lit_count; ++i) {
          int s = l[i].kind == LLKV_LIT_DECIMAL128 ? l[i].scale : 0;
          if (s > ds) ds = s;
        }
        if (c->scale > ds) ds = c->scale;
      }
      for (int i = 0; i < op->lit_count; ++i)
        if ((rc = lit_to_native(c, l + i, &p->in[i], &ds, p->strbuf + 16 * i, e))) { free(p->in); p->in = NULL; return rc; }
      break;
    case LLKV_OP_STARTS_WITH: case LLKV_OP_ENDS_WITH: case LLKV_OP_CONTAINS:
      /* typed_predicate.rs:25-36: every native type but String answers false; :187-209: str::starts_with / ends_with /
       * contains, through to_lowercase() on both sides when !case_sensitive (exact here for ASCII-only data) */
      if (op->lit_count != 1) return fail(e, LLKV_ERR_INTERNAL, "operator needs one literal");
      if (p->dom != DOM_STR) { p->op = LLKV_OP_IN; p->n_in = 0; break; }
      if (l->kind != LLKV_LIT_STRING) return fail(e, LLKV_ERR_PREDICATE_BUILD, "literal type mismatch: expected string");
      if (l->precision == LLKV_LIT_STRING_BY_REF) {
        p->pat_long = (const uint8_t*)(uintptr_t)l->lo;
        p->pat_len = (int)l->hi;
        p->pat_ptr = (uint8_t*)malloc((size_t)p->pat_len + 1);
        memcpy(p->pat_ptr, p->pat_long, (size_t)p->pat_len);
      } else {
        memcpy(p->pat, &l->lo, 8);
        memcpy(p->pat + 8, &l->hi, 8);
        p->pat_len = (int)(l->precision > 15 ? 15 : l->precision);
        p->pat_ptr = (uint8_t*)malloc(16);
        memcpy(p->pat_ptr, p->pat, 16);
      }
      p->ci = op->literal_bool != 0;
      if (p->ci) {
        for (int i = 0; i < p->pat_len; ++i) {
          if (p->pat_ptr[i] >= 0x80) return fail(e, LLKV_ERR_PREDICATE_BUILD, "case-insensitive patterns are ASCII-only on this path");
          if (p->pat_ptr[i] >= 'A' && p->pat_ptr[i] <= 'Z') p->pat_ptr[i] = (uint8_t)(p->pat_ptr[i] + 32);
        }
        const int32_t* off = (const int32_t*)c->values;
        const uint8_t* data = (const uint8_t*)c->aux;
        for (int64_t i = off[0]; i < off[c->n_rows]; ++i)
          if (data[i] >= 0x80)
            return fail(e, LLKV_ERR_PREDICATE_BUILD, "case-insensitive match over a column with non-ASCII strings is not on this path");
      }
      break;
    default:
      return fail(e, LLKV_ERR_PREDICATE_BUILD, "operator lacks typed literal support");
  }
  p->dec_scale_lit = ds < 0 ? c->scale : ds;
  return 0;
}

/* partial_cmp: -1/0/1, 2 = unordered (NaN) */
static inline int pv_cmp(int dom, PVal v, PVal t) {
  switch (dom) {
    case DOM_I64: return v.i < t.i ? -1 : v.i > t.i;
    case DOM_U64: case DOM_BOOL: return v.u < t.u ? -1 : v.u > t.u;
    case DOM_STR: { /* str::cmp: byte-wise, then by length */
      int c = memcmp(v.s.p, t.s.p, (size_t)(v.s.n < t.s.n ? v.s.n : t.s.n));
      if (c) return c < 0 ? -1 : 1;
      return v.s.n < t.s.n ? -1 : v.s.n > t.s.n;
    }
    case DOM_DEC: return v.d < t.d ? -1 : v.d > t.d;
    default: if (v.f != v.f || t.f != t.f) return 2; return v.f < t.f ? -1 : v.f > t.f;
  }
}
static inline int pv_eq(int dom, PVal v, PVal t) {
  switch (dom) {
    case DOM_I64: return v.i == t.i;
    case DOM_U64: case DOM_BOOL: return v.u == t.u;
    case DOM_STR: return v.s.n == t.s.n && memcmp(v.s.p, t.s.p, (size_t)v.s.n) == 0;
    case DOM_DEC: return v.d == t.d;
    default: return v.f == t.f;
  }
}
static inline int pred_matches(const TPred* p, PVal v) {
  int c;
  switch (p->op) {
    case LLKV_OP_EQUALS: return pv_eq(p->dom, v, p->a);
    case LLKV_OP_GT: return pv_cmp(p->dom, v, p->a) == 1;
    case LLKV_OP_GTE: c = pv_cmp(p->dom, v, p->a); return c == 1 || c == 0;
    case LLKV_OP_LT: return pv_cmp(p->dom, v, p->a) == -1;
    case LLKV_OP_LTE: c = pv_cmp(p->dom, v, p->a); return c == -1 || c == 0;
    case LLKV_OP_RANGE:
      if (p->lower_kind != LLKV_BOUND_UNBOUNDED) {
        c = pv_cmp(p->dom, v, p->a);
        if (!(c == 1 || (c == 0 && p->lower_kind == LLKV_BOUND_INCLUDED))) return 0;
      }
      if (p->upper_kind != LLKV_BOUND_UNBOUNDED) {
        c = pv_cmp(p->dom, v, p->b);
        if (!(c == -1 || (c == 0 && p->upper_kind == LLKV_BOUND_INCLUDED))) return 0;
      }
      return 1;
    case LLKV_OP_IN:
      for (int i = 0; i < p->n_in; ++i)
        if (pv_eq(p->dom, v, p->in[i])) return 1;
      return 0;
    case LLKV_OP_STARTS_WITH: case LLKV_OP_ENDS_WITH: case LLKV_OP_CONTAINS: {
      const int64_t len = (int64_t)v.s.n, L = p->pat_len;
      if (L > len) return 0;
      const int64_t first = p->op == LLKV_OP_ENDS_WITH ? len - L : 0;
      const int64_t last = p->op == LLKV_OP_STARTS_WITH ? 0 : len - L;
      for (int64_t at = first; at <= last; ++at) {
        int64_t k = 0;
        for (; k < L; ++k) {
          uint8_t ch = v.s.p[at + k];
          if (p->ci && ch >= 'A' && ch <= 'Z') ch = (uint8_t)(ch + 32);
          if (ch != p->pat_ptr[k]) break;
        }
        if (k == L) return 1;
      }
      return 0;
    }
  }
  return 0;
}
static inline int load_pval(const oracle_column* c, const TPred* p, uint64_t i, PVal* v) {
  switch (p->dom) {
    case DOM_I64: v->i = load_sint(c, i); return 1;
    case DOM_U64: case DOM_BOOL: v->u = load_uint(c, i); return 1;
    case DOM_F64: v->f = ((const double*)c->values)[i]; return 1;
    case DOM_F32: v->f = (double)((const float*)c->values)[i]; return 1;
    case DOM_DEC: {
      i128 x = load_dec(c, i);
      int k = p->dec_scale_lit - p->dec_scale_col;
      if (k > 0) {
        i128 r;
        if (__builtin_mul_overflow(x, pow10_i128(k), &r)) r = x < 0 ? ((i128)1 << 127) : ~((i128)1 << 127);
        x = r;
      }
      v->d = x;
      return 1;
    }
    case DOM_STR: {
      const int32_t* off = (const int32_t*)c->values;
      v->s.p = (const uint8_t*)c->aux + off[i];
      v->s.n = (uint64_t)(off[i + 1] - off[i]);
      return 1;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ row-id vector (Vec<u64>) */
typedef struct {
  uint64_t* v;
  size_t n, cap;
} RowVec;
static inline void rv_push(RowVec* r, uint64_t x) {
  if (r->n == r->cap) {
    r->cap = r->cap ? r->cap * 2 : 4096;
    r->v = (uint64_t*)realloc(r->v, r->cap * 8);
  }
  r->v[r->n++] = x;
}

/* HOT LOOP #1: RowIdNullableFilterVisitor — per chunk `for i { if pred(v[i]) out.push(rid[i]) }`
 * (llkv-column-map/src/store/scan/filter.rs:931-958); one full column scan per leaf. */
typedef struct {
  const oracle_column* c;
  const TPred* p;
  uint64_t lo, hi;
  RowVec out;
  int bad_string;
} LeafJob;
static void* leaf_scan_range(void* arg) {
  LeafJob* j = (LeafJob*)arg;
  const oracle_column* c = j->c;
  const TPred* p = j->p;
  /* specialised inner loops for the two dtypes the TPC-H configs hit hardest */
  if (p->dom == DOM_I64 && c->type == LLKV_PT_INT64 && !c->validity) {
    const int64_t* v = (const int64_t*)c->values;
    for (uint64_t i = j->lo; i < j->hi; ++i) {
      PVal x;
      x.i = v[i];
      if (pred_matches(p, x)) rv_push(&j->out, i);
    }
    return NULL;
  }
  for (uint64_t i = j->lo; i < j->hi; ++i) {
    if (!col_valid(c, i)) continue;
    PVal x;
    if (!load_pval(c, p, i, &x)) { j->bad_string = 1; continue; }
    if (pred_matches(p, x)) rv_push(&j->out, i);
  }
  return NULL;
}

static int32_t leaf_scan(const oracle_column* c, const TPred* p, uint64_t rb, uint64_t re, int n_threads, Bits* out,
                         Err* e) {
  if (n_threads < 1) n_threads = 1;
  uint64_t n = re - rb;
  if ((uint64_t)n_threads > n / 65536 + 1) n_threads = (int)(n / 65536 + 1);
  LeafJob* jobs = (LeafJob*)calloc((size_t)n_threads, sizeof(LeafJob));
  pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
  for (int t = 0; t < n_threads; ++t) {
    jobs[t].c = c;
    jobs[t].p = p;
    jobs[t].lo = rb + n * (uint64_t)t / (uint64_t)n_threads;
    jobs[t].hi = rb + n * (uint64_t)(t + 1) / (uint64_t)n_threads;
    if (n_threads > 1) pthread_create(&th[t], NULL, leaf_scan_range, &jobs[t]);
    else leaf_scan_range(&jobs[t]);
  }
  int bad = 0;
  for (int t = 0; t < n_threads; ++t) {
    if (n_threads > 1) pthread_join(th[t], NULL);
    /* Treemap::from_iter(row_ids) (llkv-table/src/table.rs:1246) */
    for (size_t k = 0; k < jobs[t].out.n; ++k) bits_set(out, jobs[t].out.v[k] - rb);
    bad |= jobs[t].bad_string;
    free(jobs[t].out.v);
  }
  free(jobs);
  free(th);
  if (bad) return fail(e, LLKV_ERR_INVALID_ARGUMENT, "string longer than 7 bytes in short-string column");
  return 0;
}

static void field_present_rows(const oracle_column* c, uint64_t rb, uint64_t re, Bits* out) {
  if (!c->validity) { bits_fill(out); return; }
  for (uint64_t i = rb; i < re; ++i)
    if (col_valid(c, i)) bits_set(out, i - rb);
}

/* ------------------------------------------------------------------ arrow-like arrays for expression evaluation */
enum { K_NULL, K_I64, K_U64, K_F64, K_DEC, K_BOOL, K_DATE32, K_STR };
typedef struct {
  int kind;    /* storage class */
  int type;    /* LLKV_PT_* logical type */
  int p, s;    /* Decimal128 precision/scale */
  size_t n;
  int scalar;  /* VectorizedExpr::Scalar: n == 1, broadcast */
  void* data;  /* int64_t / uint64_t / double / i128 / uint8_t */
  uint8_t* valid; /* byte per element, NULL = all valid */
} Arr;

static int kind_of_type(int t) {
  switch (t) {
    case LLKV_PT_NULL: return K_NULL;
    case LLKV_PT_FLOAT64: case LLKV_PT_FLOAT32: return K_F64;
    case LLKV_PT_DECIMAL128: return K_DEC;
    case LLKV_PT_BOOLEAN: return K_BOOL;
    case LLKV_PT_UTF8: return K_STR;
    case LLKV_PT_DATE32: return K_DATE32;
    default: return type_is_unsigned_int(t) ? K_U64 : K_I64;
  }
}
static size_t kind_size(int k) { return k == K_DEC ? 16 : (k == K_BOOL ? 1 : 8); }
static Arr arr_new(int type, int p, int s, size_t n) {
  Arr a;
  memset(&a, 0, sizeof(a));
  a.type = type;
  a.kind = kind_of_type(type);
  a.p = p;
  a.s = s;
  a.n = n;
  a.data = a.kind == K_NULL ? NULL : calloc(n ? n : 1, kind_size(a.kind));
  return a;
}
static void arr_free(Arr* a) {
  free(a->data);
  free(a->valid);
  a->data = NULL;
  a->valid = NULL;
}
static inline int arr_is_valid(const Arr* a, size_t i) { return a->kind != K_NULL && (!a->valid || a->valid[i]); }
static void arr_set_null(Arr* a, size_t i) {
  if (!a->valid) {
    a->valid = (uint8_t*)malloc(a->n ? a->n : 1);
    memset(a->valid, 1, a->n ? a->n : 1);
  }
  a->valid[i] = 0;
}
static Arr arr_clone(const Arr* a) {
  Arr b = *a;
  if (a->data) {
    b.data = malloc((a->n ? a->n : 1) * kind_size(a->kind));
    memcpy(b.data, a->data, a->n * kind_size(a->kind));
  }
  if (a->valid) {
    b.valid = (uint8_t*)malloc(a->n ? a->n : 1);
    memcpy(b.valid, a->valid, a->n);
  }
  return b;
}
/* Scalar -> Array of length len keeping the scalar's own type (eval.rs:41-53 `take`) */
static Arr arr_expand(const Arr* a, size_t len) {
  Arr b = arr_new(a->type, a->p, a->s, len);
  if (a->kind == K_NULL) return b;
  int valid = arr_is_valid(a, 0);
  size_t sz = kind_size(a->kind);
  for (size_t i = 0; i < len; ++i) memcpy((char*)b.data + i * sz, a->data, sz);
  if (!valid) {
    b.valid = (uint8_t*)calloc(len ? len : 1, 1);
  }
  return b;
}

/* ---- DataType algebra: get_common_type / coerce_decimals (llkv-compute/src/kernels.rs:38-45,179-242) */
typedef struct {
  int type, p, s;
} DT;
static DT dt(int t, int p, int s) { DT d = {t, p, s}; return d; }
static int dt_eq(DT a, DT b) { return a.type == b.type && (a.type != LLKV_PT_DECIMAL128 || (a.p == b.p && a.s == b.s)); }
static DT coerce_decimals(int lp, int ls, int rp, int rs) {
  int scale = ls > rs ? ls : rs;
  int li = lp - ls, ri = rp - rs;
  int id = li > ri ? li : ri;
  int prec = id + scale;
  if (prec < 1) prec = 1;
  if (prec > 38) prec = 38;
  return dt(LLKV_PT_DECIMAL128, prec, scale);
}
static int is_float_t(int t) { return t == LLKV_PT_FLOAT64 || t == LLKV_PT_FLOAT32; }
static int is_int_t(int t) { return (type_is_signed_int(t) && t != LLKV_PT_DATE32 && t != LLKV_PT_DATE64) || type_is_unsigned_int(t); }
static DT common_type(DT l, DT r) {
  if (dt_eq(l, r)) return l;
  if (l.type == LLKV_PT_NULL) return r;
  if (r.type == LLKV_PT_NULL) return l;
  if (l.type == LLKV_PT_DECIMAL128 && r.type == LLKV_PT_DECIMAL128) return coerce_decimals(l.p, l.s, r.p, r.s);
  if (l.type == LLKV_PT_DECIMAL128 || r.type == LLKV_PT_DECIMAL128) {
    DT d = l.type == LLKV_PT_DECIMAL128 ? l : r, o = l.type == LLKV_PT_DECIMAL128 ? r : l;
    if (is_float_t(o.type)) return dt(LLKV_PT_FLOAT64, 0, 0);
    if (is_int_t(o.type)) return coerce_decimals(d.p, d.s, 38, 0);
    return dt(LLKV_PT_FLOAT64, 0, 0);
  }
  if (is_float_t(l.type) || is_float_t(r.type)) return dt(LLKV_PT_FLOAT64, 0, 0);
  if (is_int_t(l.type) && is_int_t(r.type)) {
    int ls = type_is_signed_int(l.type), rs = type_is_signed_int(r.type);
    int lb = type_bits(l.type), rb = type_bits(r.type);
    int mx = lb > rb ? lb : rb;
    if (ls != rs) return mx >= 64 ? dt(LLKV_PT_FLOAT64, 0, 0) : dt(LLKV_PT_INT64, 0, 0);
    if (ls) return dt(mx >= 64 ? LLKV_PT_INT64 : mx >= 32 ? LLKV_PT_INT32 : mx >= 16 ? LLKV_PT_INT16 : LLKV_PT_INT8, 0, 0);
    return dt(mx >= 64 ? LLKV_PT_UINT64 : mx >= 32 ? LLKV_PT_UINT32 : mx >= 16 ? LLKV_PT_UINT16 : LLKV_PT_UINT8, 0, 0);
  }
  return dt(LLKV_PT_FLOAT64, 0, 0);
}

/* i128 -> f64, round to nearest even (Rust `as f64`) */
static inline double i128_to_f64(i128 v) { return (double)v; }

/* ---- arrow-cast restatement (safe mode: failures become NULL).  The crate is not vendored under /root/reference; the
 * decimal -> decimal rescale (half away from zero) and the precision-overflow -> NULL rule are pinned by
 * tests/golden/arrow_known_answers.json (decimal_mul_rescale_rounds_half_away_from_zero,
 * decimal_mul_precision_overflow_is_null_under_safe_cast, decimal_add_aligns_scales_and_carries_one_digit,
 * decimal_sub_aligns_scales): values derived from the published rule and cross-checked with Python decimal and pyarrow
 * (tests/golden/make_arrow_golden.py).  The integer / float casts below have no vector of their own. */
static int32_t arr_cast(const Arr* in, DT to, Arr* out, Err* e) {
  DT from = dt(in->type, in->p, in->s);
  if (dt_eq(from, to)) { *out = arr_clone(in); return 0; }
  Arr o = arr_new(to.type, to.p, to.s, in->n);
  o.scalar = in->scalar;
  int ik = in->kind, ok = o.kind;
  if (ik == K_NULL) { /* Null -> anything: all NULL */
    if (in->n) { o.valid = (uint8_t*)calloc(in->n, 1); }
    *out = o;
    return 0;
  }
  for (size_t i = 0; i < in->n; ++i) {
    if (!arr_is_valid(in, i)) { arr_set_null(&o, i); continue; }
    if ((ik == K_I64 || ik == K_DATE32) && ok == K_DEC) { /* cast_integer_to_decimal: v * 10^scale, precision checked */
      i128 v = ((int64_t*)in->data)[i], r;
      if (__builtin_mul_overflow(v, pow10_i128(to.s), &r) || !fits_precision(r, to.p)) arr_set_null(&o, i);
      else ((i128*)o.data)[i] = r;
    } else if (ik == K_U64 && ok == K_DEC) {
      i128 v = ((uint64_t*)in->data)[i], r;
      if (__builtin_mul_overflow(v, pow10_i128(to.s), &r) || !fits_precision(r, to.p)) arr_set_null(&o, i);
      else ((i128*)o.data)[i] = r;
    } else if (ik == K_DEC && ok == K_DEC) { /* cast_decimal_to_decimal */
      i128 x = ((i128*)in->data)[i], r;
      if (to.s >= in->s) {
        if (__builtin_mul_overflow(x, pow10_i128(to.s - in->s), &r) || !fits_precision(r, to.p)) { arr_set_null(&o, i); continue; }
      } else { /* round half away from zero */
        i128 div = pow10_i128(in->s - to.s), half = div / 2;
        i128 d = x / div, rem = x % div;
        if (x >= 0) { if (rem >= half) d += 1; } else { if (rem <= -half) d -= 1; }
        r = d;
        if (!fits_precision(r, to.p)) { arr_set_null(&o, i); continue; }
      }
      ((i128*)o.data)[i] = r;
    } else if (ik == K_DEC && ok == K_F64) { /* (x as f64) / 10^scale */
      ((double*)o.data)[i] = i128_to_f64(((i128*)in->data)[i]) / powi_f64(10.0, in->s);
    } else if ((ik == K_I64 || ik == K_DATE32) && ok == K_F64) {
      ((double*)o.data)[i] = (double)((int64_t*)in->data)[i];
    } else if (ik == K_U64 && ok == K_F64) {
      ((double*)o.data)[i] = (double)((uint64_t*)in->data)[i];
    } else if (ik == K_F64 && ok == K_I64) { /* float -> int: truncate, out of range / NaN -> NULL */
      double f = ((double*)in->data)[i];
      if (!(f > -9223372036854777856.0 && f < 9223372036854775808.0)) arr_set_null(&o, i);
      else ((int64_t*)o.data)[i] = (int64_t)f;
    } else if ((ik == K_I64 || ik == K_DATE32) && ok == K_I64) {
      int64_t v = ((int64_t*)in->data)[i];
      int bits = type_bits(to.type);
      if (bits < 64 && (v < -((int64_t)1 << (bits - 1)) || v > ((int64_t)1 << (bits - 1)) - 1)) arr_set_null(&o, i);
      else ((int64_t*)o.data)[i] = v;
    } else if (ik == K_U64 && ok == K_I64) {
      uint64_t v = ((uint64_t*)in->data)[i];
      if (v > (uint64_t)INT64_MAX) arr_set_null(&o, i);
      else ((int64_t*)o.data)[i] = (int64_t)v;
    } else if (ik == K_I64 && ok == K_U64) {
      int64_t v = ((int64_t*)in->data)[i];
      if (v < 0) arr_set_null(&o, i);
      else ((uint64_t*)o.data)[i] = (uint64_t)v;
    } else if (ik == K_BOOL && ok == K_I64) {
      ((int64_t*)o.data)[i] = ((uint8_t*)in->data)[i];
    } else if (ik == K_I64 && ok == K_BOOL) {
      ((uint8_t*)o.data)[i] = ((int64_t*)in->data)[i] != 0;
    } else if (ik == K_F64 && ok == K_F64) {
      ((double*)o.data)[i] = ((double*)in->data)[i];
    } else {
      arr_free(&o);
      return fail(e, LLKV_ERR_INTERNAL, "oracle: cast %d -> %d not restated", in->type, to.type);
    }
  }
  *out = o;
  return 0;
}

/* ---- arrow-arith numeric::{add,sub,mul,div,rem} on equal-typed inputs (kernels.rs:112-137).  Pinned by
 * tests/golden/arrow_known_answers.json: Decimal128 result types of mul / add / sub (the "intermediate_type" of the decimal
 * cases, asserted against pyarrow's identical rule by the generator), checked i64 add / sub / mul / div (int64_*_overflow_*,
 * int64_*_at_the_edge_of_the_range, int64_min_div_minus_one_overflows), truncating division and remainder
 * (int64_div_truncates_and_zero_divisors_give_null, int64_rem_*). */
static int32_t arr_arith(const Arr* l, const Arr* r, int op, Arr* out, Err* e) {
  size_t n = l->n;
  if (l->kind == K_NULL) { *out = arr_new(LLKV_PT_NULL, 0, 0, n); return 0; }
  DT rt = dt(l->type, l->p, l->s);
  if (l->kind == K_DEC) { /* arrow decimal result types */
    int p1 = l->p, s1 = l->s, p2 = r->p, s2 = r->s;
    if (op == LLKV_BIN_ADD || op == LLKV_BIN_SUB) {
      int s = s1 > s2 ? s1 : s2;
      int a = p1 - s1 > p2 - s2 ? p1 - s1 : p2 - s2;
      int p = a + s + 1;
      rt = dt(LLKV_PT_DECIMAL128, p > 38 ? 38 : p, s);
    } else if (op == LLKV_BIN_MUL) {
      int p = p1 + p2 + 1;
      rt = dt(LLKV_PT_DECIMAL128, p > 38 ? 38 : p, s1 + s2);
      if (s1 + s2 > 38) return fail(e, LLKV_ERR_INTERNAL, "Invalid argument error: Output scale of mul exceeds 38");
    } else {
      return fail(e, LLKV_ERR_INTERNAL, "oracle: Decimal128 div/rem not restated");
    }
  }
  Arr o = arr_new(rt.type, rt.p, rt.s, n);
  for (size_t i = 0; i < n; ++i) {
    if (!arr_is_valid(l, i) || !arr_is_valid(r, i)) { arr_set_null(&o, i); continue; }
    switch (l->kind) {
      case K_I64: {
        int64_t a = ((int64_t*)l->data)[i], b = ((int64_t*)r->data)[i], c = 0;
        int ov = 0;
        if (op == LLKV_BIN_ADD) ov = __builtin_add_overflow(a, b, &c);
        else if (op == LLKV_BIN_SUB) ov = __builtin_sub_overflow(a, b, &c);
        else if (op == LLKV_BIN_MUL) ov = __builtin_mul_overflow(a, b, &c);
        else if (op == LLKV_BIN_DIV) {
          if (b == 0) { arr_free(&o); return fail(e, LLKV_ERR_INTERNAL, "Divide by zero error"); }
          if (a == INT64_MIN && b == -1) ov = 1; else c = a / b;
        } else if (op == LLKV_BIN_MOD) {
          if (b == 0) { arr_free(&o); return fail(e, LLKV_ERR_INTERNAL, "Divide by zero error"); }
          c = (a == INT64_MIN && b == -1) ? 0 : a % b;
        }
        if (ov) { arr_free(&o); return fail(e, LLKV_ERR_INTERNAL, "Arithmetic overflow: Overflow happened on: %lld op %lld", (long long)a, (long long)b); }
        ((int64_t*)o.data)[i] = c;
        break;
      }
      case K_F64: {
        double a = ((double*)l->data)[i], b = ((double*)r->data)[i], c = 0;
        if (op == LLKV_BIN_ADD) c = a + b;
        else if (op == LLKV_BIN_SUB) c = a - b;
        else if (op == LLKV_BIN_MUL) c = a * b;
        else if (op == LLKV_BIN_DIV) c = a / b;
        else c = fmod(a, b);
        ((double*)o.data)[i] = c;
        break;
      }
      case K_DEC: {
        i128 a = ((i128*)l->data)[i], b = ((i128*)r->data)[i], c = 0;
        int ov = 0;
        if (op == LLKV_BIN_ADD) ov = __builtin_add_overflow(a, b, &c);
        else if (op == LLKV_BIN_SUB) ov = __builtin_sub_overflow(a, b, &c);
        else ov = __builtin_mul_overflow(a, b, &c);
        if (ov) { arr_free(&o); return fail(e, LLKV_ERR_INTERNAL, "Arithmetic overflow: Overflow happened on decimal op"); }
        ((i128*)o.data)[i] = c;
        break;
      }
      default:
        arr_free(&o);
        return fail(e, LLKV_ERR_INTERNAL, "oracle: arithmetic on type %d not restated", l->type);
    }
  }
  *out = o;
  return 0;
}

/* compute_binary (kernels.rs:99-177): coerce to the common type, zeros -> NULL before div
 * (vector: int64_div_truncates_and_zero_divisors_give_null; rem keeps its zeros: int64_rem_by_zero_is_an_error) */
static int32_t compute_binary(const Arr* l, const Arr* r, int op, Arr* out, Err* e) {
  DT ct = common_type(dt(l->type, l->p, l->s), dt(r->type, r->p, r->s));
  Arr lc, rc;
  int32_t rcode;
  if ((rcode = arr_cast(l, ct, &lc, e))) return rcode;
  if ((rcode = arr_cast(r, ct, &rc, e))) { arr_free(&lc); return rcode; }
  if (op == LLKV_BIN_DIV) {
    for (size_t i = 0; i < rc.n; ++i) {
      if (!arr_is_valid(&rc, i)) continue;
      int z = rc.kind == K_I64 ? ((int64_t*)rc.data)[i] == 0 : rc.kind == K_F64 ? ((double*)rc.data)[i] == 0.0 :
              rc.kind == K_DEC ? ((i128*)rc.data)[i] == 0 : 0;
      if (z) arr_set_null(&rc, i);
    }
  }
  rcode = arr_arith(&lc, &rc, op, out, e);
  arr_free(&lc);
  arr_free(&rc);
  return rcode;
}

/* arrow-ord cmp::{eq,neq,lt,lt_eq,gt,gt_eq}: floats use IEEE totalOrder.  Integer / decimal comparisons are pinned through the
 * reference's own filter fixtures (tests/golden/reference_known_answers.json, llkv-table/src/table.rs:2206-2355,2829-2906);
 * the float total order (NaN handling) has no reference-held vector: UNPINNED for that part only. */
static inline int64_t f64_total_key(double d) {
  int64_t b;
  memcpy(&b, &d, 8);
  return b ^ (int64_t)((uint64_t)(b >> 63) >> 1);
}
static int32_t compute_compare(const Arr* l, int op, const Arr* r, Arr* out, Err* e) {
  DT ct = common_type(dt(l->type, l->p, l->s), dt(r->type, r->p, r->s));
  Arr lc, rc;
  int32_t rcode;
  if ((rcode = arr_cast(l, ct, &lc, e))) return rcode;
  if ((rcode = arr_cast(r, ct, &rc, e))) { arr_free(&lc); return rcode; }
  size_t n = lc.n;
  Arr o = arr_new(LLKV_PT_BOOLEAN, 0, 0, n);
  for (size_t i = 0; i < n; ++i) {
    if (!arr_is_valid(&lc, i) || !arr_is_valid(&rc, i)) { arr_set_null(&o, i); continue; }
    int c;
    switch (lc.kind) {
      case K_I64: case K_DATE32: { int64_t a = ((int64_t*)lc.data)[i], b = ((int64_t*)rc.data)[i]; c = a < b ? -1 : a > b; break; }
      case K_U64: case K_STR: { uint64_t a = ((uint64_t*)lc.data)[i], b = ((uint64_t*)rc.data)[i]; c = a < b ? -1 : a > b; break; }
      case K_F64: { int64_t a = f64_total_key(((double*)lc.data)[i]), b = f64_total_key(((double*)rc.data)[i]); c = a < b ? -1 : a > b; break; }
      case K_DEC: { i128 a = ((i128*)lc.data)[i], b = ((i128*)rc.data)[i]; c = a < b ? -1 : a > b; break; }
      case K_BOOL: { int a = ((uint8_t*)lc.data)[i], b = ((uint8_t*)rc.data)[i]; c = a < b ? -1 : a > b; break; }
      default: arr_free(&lc); arr_free(&rc); arr_free(&o); return fail(e, LLKV_ERR_INTERNAL, "oracle: compare on type %d not restated", lc.type);
    }
    int res = op == LLKV_CMP_EQ ? c == 0 : op == LLKV_CMP_NE ? c != 0 : op == LLKV_CMP_LT ? c < 0 :
              op == LLKV_CMP_LE ? c <= 0 : op == LLKV_CMP_GT ? c > 0 : c >= 0;
    ((uint8_t*)o.data)[i] = (uint8_t)res;
  }
  arr_free(&lc);
  arr_free(&rc);
  *out = o;
  return 0;
}

/* ---- evaluation context: the gathered window (RecordBatch) */
typedef struct {
  const oracle_column* cols;
  int32_t n_cols;
  const uint64_t* rows; /* row ids of this window */
  size_t n;
  const llkv_scalar_node* nodes;
  int32_t n_nodes;
} EvalCtx;

/* HOT LOOP #3: gather_row_window (llkv-column-map/src/store/projection.rs:929-1352) — row ids -> array */
static int32_t gather_column(const EvalCtx* cx, uint64_t fid, Arr* out, Err* e) {
  const oracle_column* c = find_col(cx->cols, cx->n_cols, fid);
  if (!c) return fail(e, LLKV_ERR_INTERNAL, "missing column for field %llu", (unsigned long long)fid);
  Arr a = arr_new(c->type, c->precision, c->scale, cx->n);
  for (size_t k = 0; k < cx->n; ++k) {
    uint64_t i = cx->rows[k];
    if (!col_valid(c, i)) { arr_set_null(&a, k); continue; }
    switch (a.kind) {
      case K_I64: case K_DATE32: ((int64_t*)a.data)[k] = load_sint(c, i); break;
      case K_U64: ((uint64_t*)a.data)[k] = load_uint(c, i); break;
      case K_BOOL: ((uint8_t*)a.data)[k] = (uint8_t)load_uint(c, i); break;
      case K_F64: ((double*)a.data)[k] = c->type == LLKV_PT_FLOAT32 ? (double)((const float*)c->values)[i] : ((const double*)c->values)[i]; break;
      case K_DEC: ((i128*)a.data)[k] = load_dec(c, i); break;
      case K_STR: if (!load_str(c, i, &((uint64_t*)a.data)[k])) { arr_free(&a); return fail(e, LLKV_ERR_INVALID_ARGUMENT, "string longer than 7 bytes"); } break;
      default: break;
    }
  }
  *out = a;
  return 0;
}

static Arr literal_to_array(const llkv_literal* l) { /* eval.rs:521-543 */
  Arr a;
  switch (l->kind) {
    case LLKV_LIT_BOOLEAN: a = arr_new(LLKV_PT_BOOLEAN, 0, 0, 1); ((uint8_t*)a.data)[0] = l->lo != 0; break;
    case LLKV_LIT_INT128: a = arr_new(LLKV_PT_INT64, 0, 0, 1); ((int64_t*)a.data)[0] = (int64_t)lit_i128(l); break;
    case LLKV_LIT_FLOAT64: a = arr_new(LLKV_PT_FLOAT64, 0, 0, 1); ((double*)a.data)[0] = lit_f64(l); break;
    case LLKV_LIT_DECIMAL128: a = arr_new(LLKV_PT_DECIMAL128, digits_i128(lit_i128(l)), l->scale, 1); ((i128*)a.data)[0] = lit_i128(l); break;
    case LLKV_LIT_DATE32: a = arr_new(LLKV_PT_DATE32, 0, 0, 1); ((int64_t*)a.data)[0] = (int64_t)l->lo; break;
    case LLKV_LIT_STRING: {
      a = arr_new(LLKV_PT_UTF8, 0, 0, 1);
      uint8_t bytes[16];
      memcpy(bytes, &l->lo, 8);
      memcpy(bytes + 8, &l->hi, 8);
      if (l->precision == LLKV_LIT_STRING_BY_REF) memcpy(bytes, (const void*)(uintptr_t)l->lo, l->hi < 7 ? (size_t)l->hi : 7);
      uint32_t n = l->precision == LLKV_LIT_STRING_BY_REF ? (uint32_t)(l->hi < 7 ? l->hi : 7) : (uint32_t)(l->precision > 7 ? 7 : l->precision);
      pack_short_string(bytes, n, &((uint64_t*)a.data)[0]);
      break;
    }
    default: a = arr_new(LLKV_PT_NULL, 0, 0, 1); break;
  }
  a.scalar = 1;
  return a;
}

static DT literal_type(const llkv_literal* l) { /* eval.rs:166-186 */
  switch (l->kind) {
    case LLKV_LIT_BOOLEAN: return dt(LLKV_PT_BOOLEAN, 0, 0);
    case LLKV_LIT_INT128: return dt(LLKV_PT_INT64, 0, 0);
    case LLKV_LIT_FLOAT64: return dt(LLKV_PT_FLOAT64, 0, 0);
    case LLKV_LIT_DECIMAL128: return dt(LLKV_PT_DECIMAL128, digits_i128(lit_i128(l)), l->scale);
    case LLKV_LIT_DATE32: return dt(LLKV_PT_DATE32, 0, 0);
    case LLKV_LIT_STRING: return dt(LLKV_PT_UTF8, 0, 0);
    default: return dt(LLKV_PT_NULL, 0, 0);
  }
}
/* infer_result_type (eval.rs:71-148): binary = get_common_type ignoring the operator (:223-225) */
static int32_t infer_type(const EvalCtx* cx, int idx, DT* out, Err* e) {
  const llkv_scalar_node* nd = &cx->nodes[idx];
  switch (nd->tag) {
    case LLKV_SE_COLUMN: {
      const oracle_column* c = find_col(cx->cols, cx->n_cols, nd->field_id);
      if (!c) return fail(e, LLKV_ERR_INTERNAL, "missing column for field %llu", (unsigned long long)nd->field_id);
      *out = dt(c->type, c->precision, c->scale);
      return 0;
    }
    case LLKV_SE_LITERAL: *out = literal_type(&nd->literal); return 0;
    case LLKV_SE_BINARY: {
      DT l, r;
      int32_t rc;
      if ((rc = infer_type(cx, nd->left, &l, e)) || (rc = infer_type(cx, nd->right, &r, e))) return rc;
      *out = common_type(l, r);
      return 0;
    }
    case LLKV_SE_COMPARE: case LLKV_SE_NOT: case LLKV_SE_IS_NULL: *out = dt(LLKV_PT_BOOLEAN, 0, 0); return 0;
    case LLKV_SE_CAST: *out = dt(nd->cast_type, nd->cast_precision, nd->cast_scale); return 0;
    default: return fail(e, LLKV_ERR_INTERNAL, "oracle: scalar node tag %d not restated", nd->tag);
  }
}

/* NumericFastPath eligibility (fast_numeric.rs:40-60,250-300): int/float columns+literals, no Divide */
static int fast_numeric_ok(const EvalCtx* cx, int idx, DT* out) {
  const llkv_scalar_node* nd = &cx->nodes[idx];
  switch (nd->tag) {
    case LLKV_SE_COLUMN: {
      const oracle_column* c = find_col(cx->cols, cx->n_cols, nd->field_id);
      if (!c || !(is_int_t(c->type) || is_float_t(c->type))) return 0;
      *out = dt(c->type, 0, 0);
      return 1;
    }
    case LLKV_SE_LITERAL:
      if (nd->literal.kind == LLKV_LIT_INT128 || nd->literal.kind == LLKV_LIT_NULL || nd->literal.kind == LLKV_LIT_DECIMAL128) { *out = dt(LLKV_PT_INT64, 0, 0); return 1; }
      if (nd->literal.kind == LLKV_LIT_FLOAT64) { *out = dt(LLKV_PT_FLOAT64, 0, 0); return 1; }
      return 0;
    case LLKV_SE_BINARY: {
      if (nd->op == LLKV_BIN_DIV || nd->op > LLKV_BIN_MOD) return 0;
      DT l, r;
      if (!fast_numeric_ok(cx, nd->left, &l) || !fast_numeric_ok(cx, nd->right, &r)) return 0;
      *out = common_type(l, r);
      return is_int_t(out->type) || is_float_t(out->type);
    }
    default: return 0;
  }
}
/* NumericFastPath::execute: every column cast to the target type first, all ops in the target type */
static int32_t eval_fast(const EvalCtx* cx, int idx, DT target, Arr* out, Err* e) {
  const llkv_scalar_node* nd = &cx->nodes[idx];
  int32_t rc;
  switch (nd->tag) {
    case LLKV_SE_COLUMN: {
      Arr g;
      if ((rc = gather_column(cx, nd->field_id, &g, e))) return rc;
      rc = arr_cast(&g, target, out, e);
      arr_free(&g);
      return rc;
    }
    case LLKV_SE_LITERAL: { /* make_literal_array (fast_numeric.rs:131-190) */
      Arr a = arr_new(target.type, 0, 0, cx->n);
      const llkv_literal* l = &nd->literal;
      for (size_t i = 0; i < cx->n; ++i) {
        if (l->kind == LLKV_LIT_NULL) { arr_set_null(&a, i); continue; }
        if (a.kind == K_F64) ((double*)a.data)[i] = l->kind == LLKV_LIT_FLOAT64 ? lit_f64(l) : (double)lit_i128(l);
        else if (a.kind == K_I64) {
          if (l->kind == LLKV_LIT_FLOAT64) { arr_free(&a); return fail(e, LLKV_ERR_INTERNAL, "oracle: float literal in integer fast path"); }
          i128 v = lit_i128(l);
          if (v < INT64_MIN || v > INT64_MAX) { arr_free(&a); return fail(e, LLKV_ERR_INVALID_ARGUMENT, "literal out of range for Int64"); }
          ((int64_t*)a.data)[i] = (int64_t)v;
        } else { arr_free(&a); return fail(e, LLKV_ERR_INTERNAL, "oracle: fast path target %d not restated", target.type); }
      }
      *out = a;
      return 0;
    }
    default: {
      Arr l, r;
      if ((rc = eval_fast(cx, nd->left, target, &l, e))) return rc;
      if ((rc = eval_fast(cx, nd->right, target, &r, e))) { arr_free(&l); return rc; }
      rc = arr_arith(&l, &r, nd->op, out, e);
      arr_free(&l);
      arr_free(&r);
      return rc;
    }
  }
}

/* try_evaluate_vectorized (eval.rs:616-750) */
static int32_t eval_vec(const EvalCtx* cx, int idx, Arr* out, Err* e) {
  const llkv_scalar_node* nd = &cx->nodes[idx];
  int32_t rc;
  switch (nd->tag) {
    case LLKV_SE_COLUMN: return gather_column(cx, nd->field_id, out, e);
    case LLKV_SE_LITERAL: *out = literal_to_array(&nd->literal); return 0;
    case LLKV_SE_BINARY: {
      Arr l, r;
      if ((rc = eval_vec(cx, nd->left, &l, e))) return rc;
      if ((rc = eval_vec(cx, nd->right, &r, e))) { arr_free(&l); return rc; }
      int both_scalar = l.scalar && r.scalar;
      if (l.scalar && !r.scalar) { Arr x = arr_expand(&l, r.n); arr_free(&l); l = x; }
      if (r.scalar && !l.scalar) { Arr x = arr_expand(&r, l.n); arr_free(&r); r = x; }
      if (nd->op > LLKV_BIN_MOD) { arr_free(&l); arr_free(&r); return fail(e, LLKV_ERR_INTERNAL, "oracle: binary op %d not restated", nd->op); }
      rc = compute_binary(&l, &r, nd->op, out, e);
      if (!rc) out->scalar = both_scalar;
      arr_free(&l);
      arr_free(&r);
      return rc;
    }
    case LLKV_SE_CAST: {
      Arr in;
      if ((rc = eval_vec(cx, nd->left, &in, e))) return rc;
      rc = arr_cast(&in, dt(nd->cast_type, nd->cast_precision, nd->cast_scale), out, e);
      if (!rc) out->scalar = in.scalar;
      arr_free(&in);
      return rc;
    }
    case LLKV_SE_COMPARE: {
      Arr l, r;
      if ((rc = eval_vec(cx, nd->left, &l, e))) return rc;
      if ((rc = eval_vec(cx, nd->right, &r, e))) { arr_free(&l); return rc; }
      int both_scalar = l.scalar && r.scalar;
      if (l.scalar && !r.scalar) { Arr x = arr_expand(&l, r.n); arr_free(&l); l = x; }
      if (r.scalar && !l.scalar) { Arr x = arr_expand(&r, l.n); arr_free(&r); r = x; }
      rc = compute_compare(&l, nd->op, &r, out, e);
      if (!rc) out->scalar = both_scalar;
      arr_free(&l);
      arr_free(&r);
      return rc;
    }
    default: return fail(e, LLKV_ERR_INTERNAL, "oracle: scalar node tag %d not restated", nd->tag);
  }
}

/* ScalarEvaluator::evaluate_batch_simplified (eval.rs:565-614): result cast to the inferred "preferred" type */
static int32_t eval_batch_arrow(const EvalCtx* cx, int root, Arr* out, Err* e) {
  DT pref;
  int32_t rc;
  if ((rc = infer_type(cx, root, &pref, e))) return rc;
  DT fo;
  if ((is_int_t(pref.type) || is_float_t(pref.type)) && fast_numeric_ok(cx, root, &fo) && dt_eq(fo, pref))
    return eval_fast(cx, root, pref, out, e);
  Arr v;
  if ((rc = eval_vec(cx, root, &v, e))) return rc;
  if (v.scalar) { /* materialize(Scalar): expanded with the scalar's own type */
    *out = arr_expand(&v, cx->n);
    arr_free(&v);
    return 0;
  }
  if (dt_eq(dt(v.type, v.p, v.s), pref)) { *out = v; return 0; }
  Arr c;
  Err ignore = {NULL, 0};
  if (arr_cast(&v, pref, &c, &ignore) == 0) { arr_free(&v); *out = c; } /* cast(..).unwrap_or(array) */
  else *out = v;
  return 0;
}

/* ---- exact mode: PlanValue interpreter used by GROUP BY aggregates (llkv-executor/src/lib.rs:7008-7440) */
typedef struct {
  int kind; /* 0 Null, 1 Integer(i64), 2 Float(f64), 3 Decimal(i128, scale) */
  int64_t i;
  double f;
  i128 d;
  int s;
} PlanValue;

static int dec_new_ok(i128 v) { return digits_i128(v) <= 38; } /* DecimalValue::new (decimal.rs:67-76) */
static int32_t dec_rescale(i128 v, int s, int target, i128* out) { /* scalar::decimal::rescale (decimal.rs:30-64), upscale only */
  if (target == s) { *out = v; return 0; }
  if (target < s) return 1;
  i128 r;
  if (__builtin_mul_overflow(v, pow10_i128(target - s), &r)) return 1;
  if (!dec_new_ok(r)) return 1;
  *out = r;
  return 0;
}
static int32_t dec_binary(int op, i128 a, int sa, i128 b, int sb, i128* out, int* so) {
  if (op == LLKV_BIN_ADD || op == LLKV_BIN_SUB) {
    int t = sa > sb ? sa : sb;
    i128 l, r, c;
    if (dec_rescale(a, sa, t, &l) || dec_rescale(b, sb, t, &r)) return 1;
    if (op == LLKV_BIN_ADD ? __builtin_add_overflow(l, r, &c) : __builtin_sub_overflow(l, r, &c)) return 1;
    if (!dec_new_ok(c)) return 1;
    *out = c;
    *so = t;
    return 0;
  }
  if (op == LLKV_BIN_MUL) {
    int t = sa + sb;
    if (t > 38 || t < -38) return 2;
    i128 c;
    if (__builtin_mul_overflow(a, b, &c)) return 1;
    if (!dec_new_ok(c)) return 1;
    *out = c;
    *so = t;
    return 0;
  }
  return 3;
}
int32_t llkv_oracle_decimal_binary(int32_t op, const llkv_literal* a, const llkv_literal* b, llkv_literal* out) {
  i128 r;
  int s;
  int32_t rc = dec_binary(op, lit_i128(a), a->scale, lit_i128(b), b->scale, &r, &s);
  if (rc) return rc;
  memset(out, 0, sizeof(*out));
  out->kind = LLKV_LIT_DECIMAL128;
  out->scale = (int8_t)s;
  out->precision = (uint8_t)digits_i128(r);
  out->lo = (uint64_t)(u128)r;
  out->hi = (uint64_t)((u128)r >> 64);
  return 0;
}

static int32_t eval_row_exact(const EvalCtx* cx, int idx, uint64_t row, PlanValue* out, Err* e) {
  const llkv_scalar_node* nd = &cx->nodes[idx];
  memset(out, 0, sizeof(*out));
  switch (nd->tag) {
    case LLKV_SE_COLUMN: {
      const oracle_column* c = find_col(cx->cols, cx->n_cols, nd->field_id);
      if (!c) return fail(e, LLKV_ERR_INTERNAL, "missing column for field %llu", (unsigned long long)nd->field_id);
      if (!col_valid(c, row)) return 0;
      if (c->type == LLKV_PT_DECIMAL128) { out->kind = 3; out->d = load_dec(c, row); out->s = c->scale; }
      else if (is_float_t(c->type)) { out->kind = 2; out->f = c->type == LLKV_PT_FLOAT32 ? ((const float*)c->values)[row] : ((const double*)c->values)[row]; }
      else if (type_is_unsigned_int(c->type)) { out->kind = 1; out->i = (int64_t)load_uint(c, row); }
      else { out->kind = 1; out->i = load_sint(c, row); }
      return 0;
    }
    case LLKV_SE_LITERAL: {
      const llkv_literal* l = &nd->literal;
      if (l->kind == LLKV_LIT_INT128) { out->kind = 1; out->i = (int64_t)lit_i128(l); }
      else if (l->kind == LLKV_LIT_FLOAT64) { out->kind = 2; out->f = lit_f64(l); }
      else if (l->kind == LLKV_LIT_DECIMAL128) { out->kind = 3; out->d = lit_i128(l); out->s = l->scale; }
      else if (l->kind != LLKV_LIT_NULL) return fail(e, LLKV_ERR_INVALID_ARGUMENT, "oracle: literal kind %d not restated in exact mode", l->kind);
      return 0;
    }
    case LLKV_SE_BINARY: {
      PlanValue l, r;
      int32_t rc;
      if ((rc = eval_row_exact(cx, nd->left, row, &l, e)) || (rc = eval_row_exact(cx, nd->right, row, &r, e))) return rc;
      if (l.kind == 0 || r.kind == 0) return 0; /* NULL propagates */
      if (l.kind == 3 || r.kind == 3) {
        if (l.kind == 2 || r.kind == 2) return fail(e, LLKV_ERR_INVALID_ARGUMENT, "Cannot perform exact decimal arithmetic with Float operands");
        i128 a = l.kind == 3 ? l.d : (i128)l.i, b = r.kind == 3 ? r.d : (i128)r.i;
        int sa = l.kind == 3 ? l.s : 0, sb = r.kind == 3 ? r.s : 0;
        if (nd->op > LLKV_BIN_MUL) return fail(e, LLKV_ERR_INVALID_ARGUMENT, "oracle: decimal op %d not restated in exact mode", nd->op);
        if (dec_binary(nd->op, a, sa, b, sb, &out->d, &out->s))
          return fail(e, LLKV_ERR_INVALID_ARGUMENT, "Decimal %s overflow", nd->op == LLKV_BIN_ADD ? "addition" : nd->op == LLKV_BIN_SUB ? "subtraction" : "multiplication");
        out->kind = 3;
        return 0;
      }
      if (l.kind == 1 && r.kind == 1) { /* integer arithmetic: checked (lib.rs:7150-7226) */
        int64_t c = 0;
        int ov = 0;
        if (nd->op == LLKV_BIN_ADD) ov = __builtin_add_overflow(l.i, r.i, &c);
        else if (nd->op == LLKV_BIN_SUB) ov = __builtin_sub_overflow(l.i, r.i, &c);
        else if (nd->op == LLKV_BIN_MUL) ov = __builtin_mul_overflow(l.i, r.i, &c);
        else return fail(e, LLKV_ERR_INVALID_ARGUMENT, "oracle: integer op %d not restated in exact mode", nd->op);
        if (ov) return fail(e, LLKV_ERR_INVALID_ARGUMENT, "Integer overflow in %s", nd->op == LLKV_BIN_ADD ? "addition" : nd->op == LLKV_BIN_SUB ? "subtraction" : "multiplication");
        out->kind = 1;
        out->i = c;
        return 0;
      }
      double a = l.kind == 2 ? l.f : (double)l.i, b = r.kind == 2 ? r.f : (double)r.i;
      out->kind = 2;
      if (nd->op == LLKV_BIN_ADD) out->f = a + b;
      else if (nd->op == LLKV_BIN_SUB) out->f = a - b;
      else if (nd->op == LLKV_BIN_MUL) out->f = a * b;
      else return fail(e, LLKV_ERR_INVALID_ARGUMENT, "oracle: float op %d not restated in exact mode", nd->op);
      return 0;
    }
    default: return fail(e, LLKV_ERR_INTERNAL, "oracle: scalar node tag %d not restated in exact mode", nd->tag);
  }
}
/* plan_values_to_arrow_array (llkv-executor/src/lib.rs:298-415): per-group temp column */
static int32_t eval_batch_exact(const EvalCtx* cx, int root, Arr* out, Err* e) {
  PlanValue* vals = (PlanValue*)calloc(cx->n ? cx->n : 1, sizeof(PlanValue));
  int kind = 0, scale = 0;
  int32_t rc = 0;
  for (size_t k = 0; k < cx->n && !rc; ++k) {
    rc = eval_row_exact(cx, root, cx->rows[k], &vals[k], e);
    if (!rc && vals[k].kind) {
      if (vals[k].kind > kind) kind = vals[k].kind == 3 ? 3 : (kind == 3 ? 3 : vals[k].kind);
      if (vals[k].kind == 3 && vals[k].s > scale) scale = vals[k].s;
    }
  }
  if (rc) { free(vals); return rc; }
  Arr a = kind == 3 ? arr_new(LLKV_PT_DECIMAL128, 38, scale, cx->n) : kind == 2 ? arr_new(LLKV_PT_FLOAT64, 0, 0, cx->n) : kind == 1 ? arr_new(LLKV_PT_INT64, 0, 0, cx->n) : arr_new(LLKV_PT_NULL, 0, 0, cx->n);
  for (size_t k = 0; k < cx->n; ++k) {
    if (!vals[k].kind) { if (a.kind != K_NULL) arr_set_null(&a, k); continue; }
    if (kind == 3) {
      i128 v = vals[k].kind == 3 ? vals[k].d : (i128)vals[k].i;
      int s = vals[k].kind == 3 ? vals[k].s : 0;
      if (dec_rescale(v, s, scale, &v)) { arr_free(&a); free(vals); return fail(e, LLKV_ERR_INVALID_ARGUMENT, "Decimal rescale overflow"); }
      ((i128*)a.data)[k] = v;
    } else if (kind == 2) ((double*)a.data)[k] = vals[k].kind == 2 ? vals[k].f : (double)vals[k].i;
    else ((int64_t*)a.data)[k] = vals[k].i;
  }
  free(vals);
  *out = a;
  return 0;
}

/* ------------------------------------------------------------------ predicate program interpreter
 * collect_row_ids_for_program (llkv-scan/src/predicate.rs:32-193) with (rows, domain) pairs so that
 * Not{domain} = domain - rows follows compile_domain (llkv-compute/src/program.rs:447-520). */
typedef struct {
  Bits t, d;
  int has_d;
} StackEnt;

typedef struct {
  const oracle_column* cols;
  int32_t n_cols;
  const oracle_program* prog;
  uint64_t rb, re, n_table_rows;
  int n_threads;
  int need_domain;
} ProgCtx;

static int32_t eval_leaf(const ProgCtx* pc, const llkv_eval_op* op, StackEnt* out, Err* e) {
  uint64_t n = pc->re - pc->rb;
  const oracle_column* c = find_col(pc->cols, pc->n_cols, op->field_id);
  if (!c) return fail(e, LLKV_ERR_NOT_FOUND, "unknown field %llu", (unsigned long long)op->field_id);
  out->t = bits_new(n);
  out->has_d = 0;
  int32_t rc = 0;
  if (op->operator_tag == LLKV_OP_IS_NOT_NULL) field_present_rows(c, pc->rb, pc->re, &out->t);
  else if (op->operator_tag == LLKV_OP_IS_NULL) { /* all table rows - present rows (table.rs:1133-1142) */
    Bits p = bits_new(n);
    field_present_rows(c, pc->rb, pc->re, &p);
    bits_fill(&out->t);
    Bits tmp = bits_clone(&out->t);
    bits_andnot_from(&out->t, &tmp, &p);
    bits_free(&tmp);
    bits_free(&p);
  } else if (op->operator_tag == LLKV_OP_RANGE && op->lower_kind == LLKV_BOUND_UNBOUNDED && op->upper_kind == LLKV_BOUND_UNBOUNDED) {
    bits_fill(&out->t); /* all table rows (table.rs:1146-1154) */
  } else {
    TPred p;
    if ((rc = build_pred(c, op, pc->prog->literals, &p, e))) { free(p.strbuf); free(p.pat_ptr); bits_free(&out->t); return rc; }
    rc = leaf_scan(c, &p, pc->rb, pc->re, pc->n_threads, &out->t, e);
    free(p.in);
    free(p.strbuf);
    free(p.pat_ptr);
    if (rc) { bits_free(&out->t); return rc; }
  }
  if (pc->need_domain) { /* DomainOp::PushFieldAll(field) */
    out->d = bits_new(n);
    field_present_rows(c, pc->rb, pc->re, &out->d);
    out->has_d = 1;
  }
  return 0;
}

/* Expr::Compare leaf (llkv-scan/src/predicate.rs:333-396,562-663): evaluate both sides in 4096-row chunks */
static int32_t eval_compare_leaf(const ProgCtx* pc, int left, int cmp, int right, StackEnt* out, Err* e) {
  uint64_t n = pc->re - pc->rb;
  out->t = bits_new(n);
  out->d = bits_new(n);
  out->has_d = 1;
  const size_t CH = 4096; /* llkv-scan/src/predicate.rs:29 */
  uint64_t* rows = (uint64_t*)malloc(CH * 8);
  int32_t rc = 0;
  for (uint64_t base = pc->rb; base < pc->re && !rc; base += CH) {
    size_t m = (size_t)((pc->re - base) < CH ? (pc->re - base) : CH);
    for (size_t k = 0; k < m; ++k) rows[k] = base + k;
    EvalCtx cx = {pc->cols, pc->n_cols, rows, m, pc->prog->nodes, pc->prog->n_nodes};
    Arr l, r, res;
    if ((rc = eval_batch_arrow(&cx, left, &l, e))) break;
    if ((rc = eval_batch_arrow(&cx, right, &r, e))) { arr_free(&l); break; }
    rc = compute_compare(&l, cmp, &r, &res, e);
    arr_free(&l);
    arr_free(&r);
    if (rc) break;
    for (size_t k = 0; k < m; ++k) {
      if (!arr_is_valid(&res, k)) continue;
      bits_set(&out->d, base + k - pc->rb);
      if (((uint8_t*)res.data)[k]) bits_set(&out->t, base + k - pc->rb);
    }
    arr_free(&res);
  }
  free(rows);
  if (rc) { bits_free(&out->t); bits_free(&out->d); }
  return rc;
}

static int32_t eval_in_list_leaf(const ProgCtx* pc, const llkv_eval_op* op, StackEnt* out, Err* e) {
  /* evaluate_in_list_over_rows (predicate.rs:442-560): SQL IN with NULL semantics */
  uint64_t n = pc->re - pc->rb;
  out->t = bits_new(n);
  out->d = bits_new(n);
  out->has_d = 1;
  const size_t CH = 4096;
  uint64_t* rows = (uint64_t*)malloc(CH * 8);
  int32_t rc = 0;
  for (uint64_t base = pc->rb; base < pc->re && !rc; base += CH) {
    size_t m = (size_t)((pc->re - base) < CH ? (pc->re - base) : CH);
    for (size_t k = 0; k < m; ++k) rows[k] = base + k;
    EvalCtx cx = {pc->cols, pc->n_cols, rows, m, pc->prog->nodes, pc->prog->n_nodes};
    Arr target;
    if ((rc = eval_batch_arrow(&cx, op->expr_left, &target, e))) break;
    uint8_t* matched = (uint8_t*)calloc(m, 1);
    uint8_t* saw_null = (uint8_t*)calloc(m, 1);
    for (int li = 0; li < op->child_count && !rc; ++li) {
      Arr item, res;
      if ((rc = eval_batch_arrow(&cx, pc->prog->list_roots[op->expr_right + li], &item, e))) break;
      rc = compute_compare(&target, LLKV_CMP_EQ, &item, &res, e);
      arr_free(&item);
      if (rc) break;
      for (size_t k = 0; k < m; ++k) {
        if (!arr_is_valid(&res, k)) saw_null[k] = 1;
        else if (((uint8_t*)res.data)[k]) matched[k] = 1;
      }
      arr_free(&res);
    }
    if (!rc)
      for (size_t k = 0; k < m; ++k) {
        if (!arr_is_valid(&target, k)) continue; /* NULL IN (...) -> NULL */
        int known = matched[k] || !saw_null[k];
        if (!known) continue;
        bits_set(&out->d, base + k - pc->rb);
        int val = op->negated ? !matched[k] : matched[k];
        if (val) bits_set(&out->t, base + k - pc->rb);
      }
    free(matched);
    free(saw_null);
    arr_free(&target);
  }
  free(rows);
  if (rc) { bits_free(&out->t); bits_free(&out->d); }
  return rc;
}

static int32_t run_program(const ProgCtx* pc, Bits* result, Err* e) {
  const oracle_program* pg = pc->prog;
  uint64_t n = pc->re - pc->rb;
  StackEnt* st = (StackEnt*)calloc((size_t)pg->n_ops + 1, sizeof(StackEnt));
  int sp = 0;
  int32_t rc = 0;
  for (int i = 0; i < pg->n_ops && !rc; ++i) {
    const llkv_eval_op* op = &pg->ops[i];
    switch (op->tag) {
      case LLKV_EV_PUSH_PREDICATE: rc = eval_leaf(pc, op, &st[sp], e); if (!rc) ++sp; break;
      case LLKV_EV_FUSED_AND: { /* N separate scans + AND (llkv-table/src/table.rs:1173-1200) */
        StackEnt acc;
        memset(&acc, 0, sizeof(acc));
        for (int k = 0; k < op->child_count && !rc; ++k) {
          StackEnt x;
          rc = eval_leaf(pc, &pg->ops[i + 1 + k], &x, e);
          if (rc) break;
          if (k == 0) acc = x;
          else {
            bits_and(&acc.t, &x.t);
            if (acc.has_d) bits_and(&acc.d, &x.d);
            bits_free(&x.t);
            if (x.has_d) bits_free(&x.d);
          }
        }
        i += op->child_count;
        if (!rc) st[sp++] = acc;
        else if (acc.t.w) { bits_free(&acc.t); if (acc.has_d) bits_free(&acc.d); }
        break;
      }
      case LLKV_EV_PUSH_COMPARE: rc = eval_compare_leaf(pc, op->expr_left, op->cmp_op, op->expr_right, &st[sp], e); if (!rc) ++sp; break;
      case LLKV_EV_PUSH_IN_LIST: rc = eval_in_list_leaf(pc, op, &st[sp], e); if (!rc) ++sp; break;
      case LLKV_EV_PUSH_IS_NULL: {
        /* collect_row_ids_for_is_null: rows where expr IS [NOT] NULL; domain = rows where all fields present */
        StackEnt x;
        x.t = bits_new(n);
        x.d = bits_new(n);
        x.has_d = 1;
        const size_t CH = 4096;
        uint64_t* rows = (uint64_t*)malloc(CH * 8);
        for (uint64_t base = pc->rb; base < pc->re && !rc; base += CH) {
          size_t m = (size_t)((pc->re - base) < CH ? (pc->re - base) : CH);
          for (size_t k = 0; k < m; ++k) rows[k] = base + k;
          EvalCtx cx = {pc->cols, pc->n_cols, rows, m, pg->nodes, pg->n_nodes};
          Arr v;
          if ((rc = eval_batch_arrow(&cx, op->expr_left, &v, e))) break;
          for (size_t k = 0; k < m; ++k) {
            int isnull = !arr_is_valid(&v, k);
            bits_set(&x.d, base + k - pc->rb);
            if (op->negated ? !isnull : isnull) bits_set(&x.t, base + k - pc->rb);
          }
          arr_free(&v);
        }
        free(rows);
        if (rc) { bits_free(&x.t); bits_free(&x.d); } else st[sp++] = x;
        break;
      }
      case LLKV_EV_PUSH_LITERAL: {
        StackEnt x;
        x.t = bits_new(n);
        if (op->literal_bool) bits_fill(&x.t);
        x.d = bits_new(n);
        bits_fill(&x.d);
        x.has_d = 1;
        st[sp++] = x;
        break;
      }
      case LLKV_EV_AND:
      case LLKV_EV_OR: {
        if (op->child_count <= 0 || sp < op->child_count) { rc = fail(e, LLKV_ERR_INTERNAL, "%s opcode underflow", op->tag == LLKV_EV_AND ? "AND" : "OR"); break; }
        StackEnt acc = st[--sp];
        for (int k = 1; k < op->child_count; ++k) {
          StackEnt x = st[--sp];
          if (op->tag == LLKV_EV_AND) bits_and(&acc.t, &x.t); else bits_or(&acc.t, &x.t);
          if (pc->need_domain) {
            if (!acc.has_d || !x.has_d) { rc = fail(e, LLKV_ERR_INTERNAL, "domain missing"); }
            else if (op->tag == LLKV_EV_AND) bits_and(&acc.d, &x.d); /* DomainOp::Intersect */
            else bits_or(&acc.d, &x.d);                              /* DomainOp::Union */
          }
          bits_free(&x.t);
          if (x.has_d) bits_free(&x.d);
        }
        st[sp++] = acc;
        break;
      }
      case LLKV_EV_NOT: {
        if (sp < 1) { rc = fail(e, LLKV_ERR_INTERNAL, "NOT opcode underflow"); break; }
        StackEnt* x = &st[sp - 1];
        if (!x->has_d) { rc = fail(e, LLKV_ERR_INTERNAL, "domain missing"); break; }
        Bits r = bits_new(n);
        bits_andnot_from(&r, &x->d, &x->t); /* domain_rows - operand */
        bits_free(&x->t);
        x->t = r;
        break;
      }
      default: rc = fail(e, LLKV_ERR_INTERNAL, "unknown eval op tag %d", op->tag);
    }
  }
  if (!rc && sp != 1) rc = fail(e, LLKV_ERR_INTERNAL, "Program stack empty after evaluation");
  if (!rc) { *result = st[0].t; if (st[0].has_d) bits_free(&st[0].d); sp = 0; }
  for (int i = 0; i < sp; ++i) { bits_free(&st[i].t); if (st[i].has_d) bits_free(&st[i].d); }
  free(st);
  return rc;
}

/* ------------------------------------------------------------------ MVCC
 * TxnIdManager::status (mvcc.rs:157-171) + RowVersion::is_visible_for (mvcc.rs:282-334) */
static int txn_committed(uint64_t id, const uint64_t* nc, int32_t n) {
  if (id == TXN_ID_NONE) return 0; /* TxnStatus::None is not committed */
  if (id == TXN_ID_AUTO_COMMIT) return 1;
  for (int32_t i = 0; i < n; ++i)
    if (nc[i] == id) return 0;
  return 1; /* unknown ids are Committed */
}
int32_t llkv_oracle_mvcc_visible(uint64_t created_by, uint64_t deleted_by, uint64_t txn_id, uint64_t snapshot_id,
                                 const uint64_t* nc, int32_t n) {
  if (created_by == txn_id && txn_id != TXN_ID_AUTO_COMMIT) return deleted_by != txn_id;
  if (!txn_committed(created_by, nc, n)) return 0;
  if (created_by > snapshot_id) return 0;
  if (deleted_by == TXN_ID_NONE) return 1;
  if (deleted_by == txn_id && txn_id != TXN_ID_AUTO_COMMIT) return 0;
  if (!txn_committed(deleted_by, nc, n)) return 1;
  return deleted_by > snapshot_id;
}
/* HOT LOOP #2: filter_row_ids_impl (llkv-transaction/src/helpers.rs:112-253) */
static void mvcc_filter(const oracle_mvcc* m, uint64_t rb, Bits* rows) {
  if (!m || !m->created_by || !m->deleted_by) return; /* missing MVCC columns => all visible (helpers.rs:141-152) */
  const uint64_t* cb = (const uint64_t*)m->created_by->values;
  const uint64_t* db = (const uint64_t*)m->deleted_by->values;
  for (uint64_t w = 0; w < rows->nwords; ++w) {
    uint64_t bitsw = rows->w[w];
    while (bitsw) {
      int b = __builtin_ctzll(bitsw);
      bitsw &= bitsw - 1;
      uint64_t i = rb + w * 64 + (uint64_t)b;
      uint64_t c = col_valid(m->created_by, i) ? cb[i] : TXN_ID_AUTO_COMMIT; /* NULL created_by -> 1 (helpers.rs:214-223) */
      uint64_t d = col_valid(m->deleted_by, i) ? db[i] : TXN_ID_NONE;       /* NULL deleted_by -> MAX */
      if (!llkv_oracle_mvcc_visible(c, d, m->txn_id, m->snapshot_id, m->noncommitted, m->n_noncommitted))
        rows->w[w] &= ~(1ull << b);
    }
  }
}

static int32_t select_rows(const oracle_column* cols, int32_t n_cols, const oracle_program* prog, const oracle_mvcc* mvcc,
                           uint64_t rb, uint64_t re, int n_threads, Bits* out, Err* e) {
  uint64_t n = re - rb;
  if (!prog || prog->n_ops == 0) {
    *out = bits_new(n);
    bits_fill(out);
  } else {
    ProgCtx pc = {cols, n_cols, prog, rb, re, 0, n_threads, 0};
    for (int i = 0; i < prog->n_ops; ++i)
      if (prog->ops[i].tag == LLKV_EV_NOT) pc.need_domain = 1;
    int32_t rc = run_program(&pc, out, e);
    if (rc) return rc;
  }
  mvcc_filter(mvcc, rb, out);
  return 0;
}

int32_t llkv_oracle_filter(const oracle_column* cols, int32_t n_cols, const oracle_program* prog, const oracle_mvcc* mvcc,
                           uint64_t row_begin, uint64_t row_end, int32_t n_threads, uint64_t* out_words, uint64_t n_words,
                           uint64_t* out_count, char* err, size_t errcap) {
  Err e = {err, errcap};
  if (row_end < row_begin) return fail(&e, LLKV_ERR_INVALID_ARGUMENT, "row_end < row_begin");
  Bits b;
  int32_t rc = select_rows(cols, n_cols, prog, mvcc, row_begin, row_end, n_threads, &b, &e);
  if (rc) return rc;
  if (out_words) {
    if (n_words < b.nwords) { bits_free(&b); return fail(&e, LLKV_ERR_INVALID_ARGUMENT, "bitmap buffer too small"); }
    memcpy(out_words, b.w, b.nwords * 8);
  }
  if (out_count) *out_count = bits_count(&b);
  bits_free(&b);
  return 0;
}

/* ------------------------------------------------------------------ accumulators
 * AggregateAccumulator (llkv-aggregate/src/lib.rs:95-249), ctor :463-748, update :759-1477, finalize :1488-1939 */
enum { ACC_COUNT_STAR, ACC_COUNT_COL, ACC_SUM_I64, ACC_SUM_F64, ACC_SUM_DEC, ACC_TOTAL_I64, ACC_TOTAL_F64, ACC_TOTAL_DEC,
       ACC_AVG_I64, ACC_AVG_F64, ACC_AVG_DEC, ACC_MIN_I64, ACC_MIN_F64, ACC_MIN_DEC, ACC_MAX_I64, ACC_MAX_F64, ACC_MAX_DEC,
       ACC_COUNT_NULLS };
typedef struct {
  int kind;
  int p, s;
  int64_t i;     /* count / i64 sum / i64 min/max */
  int i_some;    /* SumInt64 value: Option<i64> */
  int has;       /* has_values / saw_value / Some(min) */
  double f;
  i128 d;
  int64_t count; /* AVG count; CountNulls: total rows */
  int64_t non_null;
} Acc;

static int32_t acc_new(const llkv_agg_spec* sp, Acc* a, Err* e) {
  memset(a, 0, sizeof(*a));
  a->p = sp->precision;
  a->s = sp->scale;
  if (sp->distinct) return fail(e, LLKV_ERR_INVALID_ARGUMENT, "DISTINCT aggregates are outside this path");
  int t = sp->data_type;
  switch (sp->kind) {
    case LLKV_AGG_COUNT: a->kind = sp->expr_root < 0 ? ACC_COUNT_STAR : ACC_COUNT_COL; return 0;
    case LLKV_AGG_COUNT_NULLS: a->kind = ACC_COUNT_NULLS; return 0;
    case LLKV_AGG_SUM: case LLKV_AGG_TOTAL: case LLKV_AGG_AVG: {
      int base = sp->kind == LLKV_AGG_SUM ? ACC_SUM_I64 : sp->kind == LLKV_AGG_TOTAL ? ACC_TOTAL_I64 : ACC_AVG_I64;
      if (t == LLKV_PT_INT64) { a->kind = base; a->i_some = 1; return 0; }
      if (t == LLKV_PT_DECIMAL128) { a->kind = base + 2; return 0; }
      if (t == LLKV_PT_FLOAT64 || t == LLKV_PT_UTF8) { a->kind = base + 1; return 0; }
      return fail(e, LLKV_ERR_INVALID_ARGUMENT, "%s aggregate not supported for column type %d", sp->kind == LLKV_AGG_SUM ? "SUM" : sp->kind == LLKV_AGG_TOTAL ? "TOTAL" : "AVG", t);
    }
    case LLKV_AGG_MIN: case LLKV_AGG_MAX: {
      int base = sp->kind == LLKV_AGG_MIN ? ACC_MIN_I64 : ACC_MAX_I64;
      if (t == LLKV_PT_INT64) { a->kind = base; return 0; }
      if (t == LLKV_PT_DECIMAL128) { a->kind = base + 2; return 0; }
      if (t == LLKV_PT_FLOAT64 || t == LLKV_PT_UTF8) { a->kind = base + 1; return 0; }
      return fail(e, LLKV_ERR_INVALID_ARGUMENT, "%s aggregate not supported for column type %d", sp->kind == LLKV_AGG_MIN ? "MIN" : "MAX", t);
    }
  }
  return fail(e, LLKV_ERR_INVALID_ARGUMENT, "unknown aggregate kind %d", sp->kind);
}

/* array_value_to_numeric (lib.rs:400-449) */
static int32_t to_numeric(const Arr* a, size_t i, double* out, Err* e) {
  switch (a->kind) {
    case K_I64: if (a->type != LLKV_PT_INT64) break; *out = (double)((int64_t*)a->data)[i]; return 0;
    case K_F64: if (a->type != LLKV_PT_FLOAT64) break; *out = ((double*)a->data)[i]; return 0;
    case K_DEC: *out = i128_to_f64(((i128*)a->data)[i]) / powi_f64(10.0, a->s); return 0;
    case K_BOOL: *out = ((uint8_t*)a->data)[i] ? 1.0 : 0.0; return 0;
    case K_NULL: *out = 0.0; return 0;
    default: break;
  }
  return fail(e, LLKV_ERR_INVALID_ARGUMENT, "Numeric coercion not supported for column type %d", a->type);
}

/* HOT LOOP #4: AggregateAccumulator::update for one batch column */
static int32_t acc_update(Acc* a, const Arr* col, size_t n_rows, Err* e) {
  int32_t rc;
  switch (a->kind) {
    case ACC_COUNT_STAR:
      if (__builtin_add_overflow(a->i, (int64_t)n_rows, &a->i)) return fail(e, LLKV_ERR_INVALID_ARGUMENT, "COUNT result exceeds i64 range");
      return 0;
    case ACC_COUNT_COL: {
      if (col->kind == K_NULL) return 0;
      int64_t c = 0;
      for (size_t i = 0; i < col->n; ++i) c += arr_is_valid(col, i);
      if (__builtin_add_overflow(a->i, c, &a->i)) return fail(e, LLKV_ERR_INVALID_ARGUMENT, "COUNT result exceeds i64 range");
      return 0;
    }
    case ACC_COUNT_NULLS: {
      a->count += (int64_t)n_rows;
      if (col->kind != K_NULL) for (size_t i = 0; i < col->n; ++i) a->non_null += arr_is_valid(col, i);
      return 0;
    }
    case ACC_SUM_I64: case ACC_TOTAL_I64: case ACC_AVG_I64: case ACC_MIN_I64: case ACC_MAX_I64: {
      if (col->kind == K_NULL) return 0;
      if (col->type != LLKV_PT_INT64) return fail(e, LLKV_ERR_INVALID_ARGUMENT, "aggregate expected an INT column in execution");
      const int64_t* v = (const int64_t*)col->data;
      for (size_t i = 0; i < col->n; ++i) {
        if (!arr_is_valid(col, i)) continue;
        if (a->kind == ACC_SUM_I64) {
          a->has = 1;
          if (__builtin_add_overflow(a->i, v[i], &a->i)) return fail(e, LLKV_ERR_INVALID_ARGUMENT, "integer overflow");
        } else if (a->kind == ACC_TOTAL_I64) {
          a->f += (double)v[i]; /* TotalInt64 accumulates as f64 (lib.rs:1012-1040) */
          a->has = 1;
        } else if (a->kind == ACC_AVG_I64) {
          if (__builtin_add_overflow(a->i, v[i], &a->i)) return fail(e, LLKV_ERR_INVALID_ARGUMENT, "AVG aggregate sum exceeds i64 range");
          a->count += 1;
        } else if (a->kind == ACC_MIN_I64) { a->i = a->has ? (v[i] < a->i ? v[i] : a->i) : v[i]; a->has = 1; }
        else { a->i = a->has ? (v[i] > a->i ? v[i] : a->i) : v[i]; a->has = 1; }
      }
      return 0;
    }
    case ACC_SUM_F64: case ACC_TOTAL_F64: case ACC_AVG_F64: case ACC_MIN_F64: case ACC_MAX_F64: {
      if (col->kind == K_NULL) return 0;
      for (size_t i = 0; i < col->n; ++i) {
        if (!arr_is_valid(col, i)) continue;
        double v = 0.0;
        if ((rc = to_numeric(col, i, &v, e))) return rc;
        if (a->kind == ACC_SUM_F64 || a->kind == ACC_TOTAL_F64) { a->f += v; a->has = 1; }
        else if (a->kind == ACC_AVG_F64) { a->f += v; a->count += 1; }
        else if (a->kind == ACC_MIN_F64) { if (!a->has) a->f = v; else if (v < a->f) a->f = v; a->has = 1; } /* partial_cmp == Less */
        else { if (!a->has) a->f = v; else if (v > a->f) a->f = v; a->has = 1; }
      }
      return 0;
    }
    default: { /* Decimal128 variants */
      if (col->kind != K_DEC) return fail(e, LLKV_ERR_INVALID_ARGUMENT, "Expected Decimal128 array");
      const i128* v = (const i128*)col->data;
      for (size_t i = 0; i < col->n; ++i) {
        if (!arr_is_valid(col, i)) continue;
        if (a->kind == ACC_SUM_DEC || a->kind == ACC_TOTAL_DEC || a->kind == ACC_AVG_DEC) {
          if (__builtin_add_overflow(a->d, v[i], &a->d))
            return fail(e, LLKV_ERR_INVALID_ARGUMENT, a->kind == ACC_TOTAL_DEC ? "Decimal128 total overflow" : "Decimal128 sum overflow");
          if (a->kind == ACC_AVG_DEC) a->count += 1;
        } else if (a->kind == ACC_MIN_DEC) { a->d = a->has ? (v[i] < a->d ? v[i] : a->d) : v[i]; a->has = 1; }
        else { a->d = a->has ? (v[i] > a->d ? v[i] : a->d) : v[i]; a->has = 1; }
      }
      return 0;
    }
  }
}

static void val_i64(llkv_agg_value* o, int64_t v, int valid) { memset(o, 0, sizeof(*o)); o->type = LLKV_PT_INT64; o->lo = (uint64_t)v; o->valid = (uint8_t)valid; }
static void val_f64(llkv_agg_value* o, double v, int valid) { memset(o, 0, sizeof(*o)); o->type = LLKV_PT_FLOAT64; memcpy(&o->lo, &v, 8); o->valid = (uint8_t)valid; if (!valid) o->lo = 0; }
static void val_dec(llkv_agg_value* o, i128 v, int p, int s, int valid) {
  memset(o, 0, sizeof(*o));
  o->type = LLKV_PT_DECIMAL128;
  o->precision = (uint8_t)p;
  o->scale = (int8_t)s;
  o->valid = (uint8_t)valid;
  if (valid) { o->lo = (uint64_t)(u128)v; o->hi = (uint64_t)((u128)v >> 64); }
}
static int32_t acc_finalize(const Acc* a, llkv_agg_value* o, Err* e) {
  (void)e;
  switch (a->kind) {
    case ACC_COUNT_STAR: case ACC_COUNT_COL: val_i64(o, a->i, 1); return 0;
    case ACC_COUNT_NULLS: val_i64(o, a->count - a->non_null, 1); return 0;
    case ACC_SUM_I64: val_i64(o, a->i, a->has); return 0;
    case ACC_SUM_F64: val_f64(o, a->f, a->has); return 0;
    case ACC_SUM_DEC: val_dec(o, a->d, a->p, a->s, 1); return 0; /* always a value: 0 when no rows (lib.rs:1567-1582) */
    case ACC_TOTAL_I64: case ACC_TOTAL_F64: val_f64(o, a->f, 1); return 0;
    case ACC_TOTAL_DEC: val_dec(o, a->d, a->p, a->s, 1); return 0;
    case ACC_AVG_I64: val_f64(o, a->count > 0 ? (double)a->i / (double)a->count : 0, a->count > 0); return 0;
    case ACC_AVG_F64: val_f64(o, a->count > 0 ? a->f / (double)a->count : 0, a->count > 0); return 0;
    case ACC_AVG_DEC: {
      if (a->count <= 0) { val_dec(o, 0, a->p, a->s, 0); return 0; }
      i128 c = a->count, avg = a->d / c, rem = a->d % c;
      i128 ar = rem < 0 ? -rem : rem;
      if (ar * 2 >= c) { if ((a->d > 0) == (c > 0) && a->d != 0) avg += 1; else avg -= 1; } /* round half away from zero (lib.rs:1731-1742) */
      val_dec(o, avg, a->p, a->s, 1);
      return 0;
    }
    case ACC_MIN_I64: case ACC_MAX_I64: val_i64(o, a->i, a->has); return 0;
    case ACC_MIN_F64: case ACC_MAX_F64: val_f64(o, a->f, a->has); return 0;
    default: val_dec(o, a->d, a->p, a->s, a->has); return 0;
  }
}

/* ------------------------------------------------------------------ aggregate drivers */
static int32_t eval_agg_arg(const EvalCtx* cx, const llkv_agg_spec* sp, int expr_mode, Arr* out, Err* e) {
  if (sp->expr_root < 0) { *out = arr_new(LLKV_PT_NULL, 0, 0, cx->n); return 0; }
  if (cx->nodes[sp->expr_root].tag == LLKV_SE_COLUMN) return gather_column(cx, cx->nodes[sp->expr_root].field_id, out, e);
  return expr_mode == LLKV_EXPR_EXACT ? eval_batch_exact(cx, sp->expr_root, out, e) : eval_batch_arrow(cx, sp->expr_root, out, e);
}

/* executor pre-normalisation of the argument for Float64 accumulators is array_value_to_numeric; nothing to do here */

static int32_t run_ungrouped(const EvalCtx* base, const Bits* sel, uint64_t rb, const llkv_agg_spec* specs, int32_t n_aggs,
                             int expr_mode, llkv_agg_value* out, Err* e) {
  Acc* accs = (Acc*)calloc((size_t)n_aggs, sizeof(Acc));
  int32_t rc = 0;
  for (int a = 0; a < n_aggs && !rc; ++a) rc = acc_new(&specs[a], &accs[a], e);
  uint64_t* rows = (uint64_t*)malloc(ROW_STREAM_CHUNK_SIZE * 8);
  size_t m = 0;
  uint64_t w = 0, cur = sel->nwords ? sel->w[0] : 0;
  int done = sel->nwords == 0;
  while (!rc) {
    /* RowStreamBuilder: next window of <= 65 536 selected row ids (llkv-scan/src/row_stream.rs:369-437) */
    while (!done && m < ROW_STREAM_CHUNK_SIZE) {
      if (!cur) {
        if (++w >= sel->nwords) { done = 1; break; }
        cur = sel->w[w];
        continue;
      }
      int b = __builtin_ctzll(cur);
      cur &= cur - 1;
      rows[m++] = rb + w * 64 + (uint64_t)b;
    }
    if (m == 0) break;
    EvalCtx cx = *base;
    cx.rows = rows;
    cx.n = m;
    for (int a = 0; a < n_aggs && !rc; ++a) {
      Arr col;
      if ((rc = eval_agg_arg(&cx, &specs[a], expr_mode, &col, e))) break;
      rc = acc_update(&accs[a], &col, m, e);
      arr_free(&col);
    }
    m = 0;
    if (done) break;
  }
  for (int a = 0; a < n_aggs && !rc; ++a) rc = acc_finalize(&accs[a], &out[a], e);
  free(rows);
  free(accs);
  return rc;
}

/* GROUP BY: FxHashMap<Vec<GroupKeyValue>, usize> in first-appearance order (llkv-executor/src/lib.rs:5064-5089) */
typedef struct {
  uint64_t* keys; /* n_keys bits per group */
  uint8_t* kvalid;
  RowVec rows;
} Group;
typedef struct {
  int n_keys;
  Group* g;
  size_t n, cap;
  int64_t* table; /* open addressing: index into g or -1 */
  size_t tcap;
} GroupMap;
static uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
static uint64_t key_hash(const uint64_t* k, const uint8_t* v, int n) {
  uint64_t h = 0x9e3779b97f4a7c15ULL;
  for (int i = 0; i < n; ++i) h = mix64(h ^ (v[i] ? k[i] : 0x5bd1e995ULL) ^ ((uint64_t)v[i] << 63));
  return h;
}
static void gm_rehash(GroupMap* m) {
  size_t nc = m->tcap ? m->tcap * 2 : 1024;
  int64_t* t = (int64_t*)malloc(nc * 8);
  for (size_t i = 0; i < nc; ++i) t[i] = -1;
  for (size_t gi = 0; gi < m->n; ++gi) {
    size_t h = key_hash(m->g[gi].keys, m->g[gi].kvalid, m->n_keys) & (nc - 1);
    while (t[h] >= 0) h = (h + 1) & (nc - 1);
    t[h] = (int64_t)gi;
  }
  free(m->table);
  m->table = t;
  m->tcap = nc;
}
static size_t gm_find_or_add(GroupMap* m, const uint64_t* k, const uint8_t* v) {
  if (m->n * 2 >= m->tcap) gm_rehash(m);
  size_t h = key_hash(k, v, m->n_keys) & (m->tcap - 1);
  while (m->table[h] >= 0) {
    Group* g = &m->g[m->table[h]];
    int same = 1;
    for (int i = 0; i < m->n_keys && same; ++i) same = g->kvalid[i] == v[i] && (!v[i] || g->keys[i] == k[i]);
    if (same) return (size_t)m->table[h];
    h = (h + 1) & (m->tcap - 1);
  }
  if (m->n == m->cap) { m->cap = m->cap ? m->cap * 2 : 256; m->g = (Group*)realloc(m->g, m->cap * sizeof(Group)); }
  Group* g = &m->g[m->n];
  memset(g, 0, sizeof(*g));
  g->keys = (uint64_t*)malloc((size_t)m->n_keys * 8);
  g->kvalid = (uint8_t*)malloc((size_t)m->n_keys);
  memcpy(g->keys, k, (size_t)m->n_keys * 8);
  memcpy(g->kvalid, v, (size_t)m->n_keys);
  m->table[h] = (int64_t)m->n;
  return m->n++;
}

/* GroupKeyValue::String for columns that hold strings longer than the 7 bytes a packed key carries: the key is the first
 * row (in scan order) holding the same bytes, found through a hash set of rows compared by content.  The caller reads the
 * string from its own column (llkv_group_key.dict = 2). */
typedef struct {
  const oracle_column* c;
  uint64_t* slot; /* row + 1, 0 = empty */
  size_t cap, n;
} StrIntern;
static uint64_t str_hash(const oracle_column* c, uint64_t row) {
  const int32_t* off = (const int32_t*)c->values;
  const uint8_t* p = (const uint8_t*)c->aux + off[row];
  uint64_t h = 1469598103934665603ull;
  for (int32_t i = 0; i < off[row + 1] - off[row]; ++i) h = (h ^ p[i]) * 1099511628211ull;
  return h;
}
static int str_same(const oracle_column* c, uint64_t a, uint64_t b) {
  const int32_t* off = (const int32_t*)c->values;
  const int32_t la = off[a + 1] - off[a], lb = off[b + 1] - off[b];
  return la == lb && memcmp((const uint8_t*)c->aux + off[a], (const uint8_t*)c->aux + off[b], (size_t)la) == 0;
}
static uint64_t intern_row(StrIntern* t, uint64_t row) {
  if ((t->n + 1) * 2 > t->cap) {
    size_t ncap = t->cap ? t->cap * 2 : 1024;
    uint64_t* ns = (uint64_t*)calloc(ncap, 8);
    for (size_t i = 0; i < t->cap; ++i)
      if (t->slot[i]) {
        size_t h = (size_t)str_hash(t->c, t->slot[i] - 1) & (ncap - 1);
        while (ns[h]) h = (h + 1) & (ncap - 1);
        ns[h] = t->slot[i];
      }
    free(t->slot);
    t->slot = ns;
    t->cap = ncap;
  }
  size_t h = (size_t)str_hash(t->c, row) & (t->cap - 1);
  while (t->slot[h]) {
    if (str_same(t->c, t->slot[h] - 1, row)) return t->slot[h] - 1;
    h = (h + 1) & (t->cap - 1);
  }
  t->slot[h] = row + 1;
  ++t->n;
  return row;
}
static int col_has_long_string(const oracle_column* c) {
  if (c->type != LLKV_PT_UTF8) return 0;
  const int32_t* off = (const int32_t*)c->values;
  for (uint64_t i = 0; i < c->n_rows; ++i)
    if (off[i + 1] - off[i] > 7) return 1;
  return 0;
}

/* group_key_value (llkv-executor/src/lib.rs:9362-9456): ints/Date32 -> i64, bool, Utf8; others unsupported */
static int32_t key_value(const oracle_column* c, uint64_t row, uint64_t* bits, uint8_t* valid, Err* e) {
  *valid = (uint8_t)col_valid(c, row);
  *bits = 0;
  if (!*valid) return 0;
  if (type_is_signed_int(c->type)) *bits = (uint64_t)load_sint(c, row);
  else if (type_is_unsigned_int(c->type)) *bits = load_uint(c, row);
  else if (c->type == LLKV_PT_BOOLEAN) *bits = load_uint(c, row) != 0;
  else if (c->type == LLKV_PT_UTF8) { if (!load_str(c, row, bits)) return fail(e, LLKV_ERR_INVALID_ARGUMENT, "GROUP BY string key longer than 7 bytes"); }
  else return fail(e, LLKV_ERR_INVALID_ARGUMENT, "GROUP BY does not support column type %d", c->type);
  return 0;
}

static int32_t run_grouped(const EvalCtx* base, const Bits* sel, uint64_t rb, const llkv_agg_spec* specs, int32_t n_aggs,
                           const uint64_t* key_fields, int32_t n_keys, int expr_mode, llkv_agg_value* out_values,
                           llkv_group_key* out_keys, uint64_t cap, uint64_t* out_groups, Err* e) {
  const oracle_column** kc = (const oracle_column**)calloc((size_t)n_keys, sizeof(void*));
  for (int k = 0; k < n_keys; ++k) {
    kc[k] = find_col(base->cols, base->n_cols, key_fields[k]);
    if (!kc[k]) { free(kc); return fail(e, LLKV_ERR_NOT_FOUND, "unknown GROUP BY field %llu", (unsigned long long)key_fields[k]); }
  }
  GroupMap gm;
  memset(&gm, 0, sizeof(gm));
  gm.n_keys = n_keys;
  uint64_t kb[16];
  uint8_t kv[16];
  int32_t rc = 0;
  StrIntern* interns = (StrIntern*)calloc((size_t)n_keys, sizeof(StrIntern));
  for (int k = 0; k < n_keys; ++k)
    if (col_has_long_string(kc[k])) interns[k].c = kc[k];
  /* pass 1: key -> group index, (row) lists */
  for (uint64_t w = 0; w < sel->nwords && !rc; ++w) {
    uint64_t cur = sel->w[w];
    while (cur && !rc) {
      int b = __builtin_ctzll(cur);
      cur &= cur - 1;
      uint64_t row = rb + w * 64 + (uint64_t)b;
      for (int k = 0; k < n_keys && !rc; ++k) {
        if (interns[k].c) {
          kv[k] = (uint8_t)col_valid(kc[k], row);
          kb[k] = kv[k] ? intern_row(&interns[k], row) : 0;
        } else {
          rc = key_value(kc[k], row, &kb[k], &kv[k], e);
        }
      }
      if (rc) break;
      size_t gi = gm_find_or_add(&gm, kb, kv);
      rv_push(&gm.g[gi].rows, row);
    }
  }
  if (!rc && gm.n > cap) rc = fail(e, LLKV_ERR_INVALID_ARGUMENT, "group capacity %llu < %llu groups", (unsigned long long)cap, (unsigned long long)gm.n);
  /* pass 2: per group a fresh accumulator fed the group's mini batch once (lib.rs:5101-5246) */
  for (size_t gi = 0; gi < gm.n && !rc; ++gi) {
    Group* g = &gm.g[gi];
    EvalCtx cx = *base;
    cx.rows = g->rows.v;
    cx.n = g->rows.n;
    for (int a = 0; a < n_aggs && !rc; ++a) {
      Acc acc;
      if ((rc = acc_new(&specs[a], &acc, e))) break;
      Arr col;
      if ((rc = eval_agg_arg(&cx, &specs[a], expr_mode, &col, e))) break;
      rc = acc_update(&acc, &col, cx.n, e);
      arr_free(&col);
      if (!rc) rc = acc_finalize(&acc, &out_values[gi * (size_t)n_aggs + (size_t)a], e);
    }
    for (int k = 0; k < n_keys && !rc; ++k) {
      llkv_group_key* ok = &out_keys[gi * (size_t)n_keys + (size_t)k];
      memset(ok, 0, sizeof(*ok));
      ok->bits = g->keys[k];
      ok->valid = g->kvalid[k];
      ok->type = kc[k]->type;
      ok->dict = interns[k].c ? 2 : 0;
    }
  }
  if (!rc) *out_groups = gm.n;
  for (int k = 0; k < n_keys; ++k) free(interns[k].slot);
  free(interns);
  for (size_t gi = 0; gi < gm.n; ++gi) { free(gm.g[gi].keys); free(gm.g[gi].kvalid); free(gm.g[gi].rows.v); }
  free(gm.g);
  free(gm.table);
  free(kc);
  return rc;
}

int32_t llkv_oracle_aggregate(const oracle_column* cols, int32_t n_cols, const oracle_program* prog, const oracle_mvcc* mvcc,
                              const llkv_agg_spec* specs, int32_t n_aggs, const llkv_scalar_node* nodes, int32_t n_nodes,
                              const uint64_t* key_fields, int32_t n_keys, int32_t expr_mode, uint64_t row_begin,
                              uint64_t row_end, int32_t n_threads, llkv_agg_value* out_values, llkv_group_key* out_keys,
                              uint64_t group_capacity, uint64_t* out_groups, char* err, size_t errcap) {
  Err e = {err, errcap};
  if (row_end < row_begin) return fail(&e, LLKV_ERR_INVALID_ARGUMENT, "row_end < row_begin");
  if (n_keys > 16) return fail(&e, LLKV_ERR_INVALID_ARGUMENT, "too many GROUP BY keys");
  Bits sel;
  int32_t rc = select_rows(cols, n_cols, prog, mvcc, row_begin, row_end, n_threads, &sel, &e);
  if (rc) return rc;
  EvalCtx base = {cols, n_cols, NULL, 0, nodes, n_nodes};
  if (n_keys == 0) {
    if (group_capacity < 1) { bits_free(&sel); return fail(&e, LLKV_ERR_INVALID_ARGUMENT, "group capacity 0"); }
    rc = run_ungrouped(&base, &sel, row_begin, specs, n_aggs, expr_mode, out_values, &e);
    if (!rc && out_groups) *out_groups = 1;
  } else {
    rc = run_grouped(&base, &sel, row_begin, specs, n_aggs, key_fields, n_keys, expr_mode, out_values, out_keys, group_capacity,
                     out_groups, &e);
  }
  bits_free(&sel);
  return rc;
}

/* ------------------------------------------------------------------ chunk blob codec */
int64_t llkv_oracle_serialize_primitive(int32_t prim_type, uint8_t precision, int8_t scale, const void* values, uint64_t n_rows,
                                        uint8_t* out, uint64_t out_cap) {
  uint64_t w = prim_type == LLKV_PT_DECIMAL128 ? 16 : prim_type == LLKV_PT_BOOLEAN ? 1 : (uint64_t)type_bits(prim_type) / 8;
  if (prim_type == LLKV_PT_FLOAT32) w = 4;
  if (prim_type == LLKV_PT_FLOAT64) w = 8;
  uint64_t bytes = w * n_rows;
  if (bytes > UINT32_MAX) return -LLKV_ERR_INTERNAL; /* "values too large" */
  if (out_cap < 24 + bytes) return -LLKV_ERR_INVALID_ARGUMENT;
  memcpy(out, "ARR0", 4);
  out[4] = 0; /* Layout::Primitive */
  out[5] = (uint8_t)prim_type;
  out[6] = prim_type == LLKV_PT_DECIMAL128 ? precision : 0;
  out[7] = prim_type == LLKV_PT_DECIMAL128 ? (uint8_t)scale : 0;
  memcpy(out + 8, &n_rows, 8);
  uint32_t vb = (uint32_t)bytes, zero = 0;
  memcpy(out + 16, &vb, 4);
  memcpy(out + 20, &zero, 4);
  memcpy(out + 24, values, bytes);
  return (int64_t)(24 + bytes);
}
