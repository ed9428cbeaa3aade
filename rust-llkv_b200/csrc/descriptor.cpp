// descriptor.cpp — host-side readers for the column-store metadata that sits in front of the scan: the column descriptor
// and its linked descriptor pages of ChunkMetadata (llkv-column-map/src/store/descriptor.rs), the order-preserving u64
// images of values (src/codecs.rs:33-65) and the chunk-pruning rule IntRanges::matches (store/pruning.rs:104-258).
// Pure C++ (no CUDA): the Rust wrapper walks the page chain with Pager::batch_get and hands every page to
// llkv_gpu_descriptor_page_parse; the library never calls back into the host language.
#include <stdint.h>
#include <string.h>

#include <unordered_set>

#include "../../include/llkv_gpu.h"

namespace {
inline uint64_t rd64(const unsigned char* p) {
  uint64_t v;
  memcpy(&v, p, 8);  // the on-disk format is little endian, and so is every host this library builds for
  return v;
}
inline uint32_t rd32(const unsigned char* p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
}
}  // namespace

int32_t llkv_set_error_message(int32_t code, const char* msg);  // llkv_gpu.cu

// ColumnDescriptor::from_le_bytes (descriptor.rs:263-299): 40 fixed bytes, then (newer files) data_type_code, padding and
// the length of the index metadata that follows
extern "C" int32_t llkv_gpu_descriptor_parse(const void* bytes, uint64_t len, llkv_column_descriptor* out) {
  if (!bytes || !out) return llkv_set_error_message(LLKV_ERR_INVALID_ARGUMENT, "NULL argument");
  if (len < 40) return llkv_set_error_message(LLKV_ERR_IO, "column descriptor shorter than 40 bytes");
  const unsigned char* b = static_cast<const unsigned char*>(bytes);
  memset(out, 0, sizeof(*out));
  out->field_id = rd64(b);
  out->head_page_pk = rd64(b + 8);
  out->tail_page_pk = rd64(b + 16);
  out->total_row_count = rd64(b + 24);
  out->total_chunk_count = rd64(b + 32);
  if (len >= 40 + 12) {
    out->data_type_code = rd32(b + 40);
    const uint32_t index_meta_len = rd32(b + 48);
    // (the reference drops index metadata that does not fit in the blob instead of failing)
    out->index_meta_len = (index_meta_len > 0 && len >= 52 + (uint64_t)index_meta_len) ? index_meta_len : 0;
  }
  return LLKV_OK;
}

// DescriptorPageHeader::from_le_bytes + the packed ChunkMetadata entries behind it (descriptor.rs:344-378, 419-434)
extern "C" int32_t llkv_gpu_descriptor_page_parse(const void* bytes, uint64_t len, uint64_t* next_page_pk, llkv_chunk_metadata* out,
                                                   uint64_t capacity, uint64_t* n_entries) {
  if (!bytes || !next_page_pk || !n_entries) return llkv_set_error_message(LLKV_ERR_INVALID_ARGUMENT, "NULL argument");
  if (len < 16) return llkv_set_error_message(LLKV_ERR_IO, "descriptor page shorter than its 16-byte header");
  const unsigned char* b = static_cast<const unsigned char*>(bytes);
  *next_page_pk = rd64(b);
  const uint64_t n = rd32(b + 8);
  *n_entries = n;
  if (16 + n * 64 > len) return llkv_set_error_message(LLKV_ERR_IO, "descriptor page is shorter than its entry count says");
  if (n > capacity) return llkv_set_error_message(LLKV_ERR_INVALID_ARGUMENT, "descriptor page holds more entries than the output buffer");
  if (n && !out) return llkv_set_error_message(LLKV_ERR_INVALID_ARGUMENT, "output buffer is NULL");
  for (uint64_t i = 0; i < n; ++i) {
    const unsigned char* e = b + 16 + i * 64;
    llkv_chunk_metadata& m = out[i];
    m.chunk_pk = rd64(e);
    m.value_order_perm_pk = rd64(e + 8);
    m.row_count = rd64(e + 16);
    m.serialized_bytes = rd64(e + 24);
    m.min_val_u64 = rd64(e + 32);
    m.max_val_u64 = rd64(e + 40);
    m.null_count = rd64(e + 48);
    m.distinct_count = rd64(e + 56);
  }
  return LLKV_OK;
}

// codecs.rs:33-65; unsigned types map to themselves (pruning.rs:123-170), f32 goes through f64
extern "C" uint64_t llkv_gpu_sortable_u64(int32_t prim_type, uint64_t value_bits) {
  switch (prim_type) {
    case LLKV_PT_INT8: return (uint64_t)(uint8_t)((uint8_t)value_bits ^ 0x80u);
    case LLKV_PT_INT16: return (uint64_t)(uint16_t)((uint16_t)value_bits ^ 0x8000u);
    case LLKV_PT_INT32: case LLKV_PT_DATE32: return (uint64_t)((uint32_t)value_bits ^ 0x80000000u);
    case LLKV_PT_INT64: case LLKV_PT_DATE64: return value_bits ^ 0x8000000000000000ull;
    case LLKV_PT_FLOAT32: {
      float f;
      const uint32_t w = (uint32_t)value_bits;
      memcpy(&f, &w, 4);
      const double d = (double)f;
      uint64_t bits;
      memcpy(&bits, &d, 8);
      return (bits & 0x8000000000000000ull) ? ~bits : (bits | 0x8000000000000000ull);
    }
    case LLKV_PT_FLOAT64: return (value_bits & 0x8000000000000000ull) ? ~value_bits : (value_bits | 0x8000000000000000ull);
    case LLKV_PT_UINT8: return value_bits & 0xffull;
    case LLKV_PT_UINT16: return value_bits & 0xffffull;
    case LLKV_PT_UINT32: return value_bits & 0xffffffffull;
    default: return value_bits;
  }
}

// IntRanges::matches for one typed range + check_overlap (pruning.rs:104-258): the chunk [min, max] (inclusive, sortable
// u64) against the range (lower, upper), each Unbounded / Included / Excluded.  1 = the chunk may hold matching rows.
extern "C" int32_t llkv_gpu_chunk_overlaps(int32_t prim_type, uint64_t chunk_min_u64, uint64_t chunk_max_u64, const llkv_range_bound* lower,
                                            const llkv_range_bound* upper) {
  if (upper && upper->kind != LLKV_BOUND_UNBOUNDED) {
    const uint64_t u = llkv_gpu_sortable_u64(prim_type, upper->value_bits);
    if (upper->kind == LLKV_BOUND_INCLUDED ? u < chunk_min_u64 : u <= chunk_min_u64) return 0;
  }
  if (lower && lower->kind != LLKV_BOUND_UNBOUNDED) {
    const uint64_t l = llkv_gpu_sortable_u64(prim_type, lower->value_bits);
    if (lower->kind == LLKV_BOUND_INCLUDED ? l > chunk_max_u64 : l >= chunk_max_u64) return 0;
  }
  return 1;
}

// compute_chunk_stats (pruning.rs:272-470), primitive arrays
namespace {
template <typename T>
int32_t chunk_stats_int(int32_t prim_type, const T* v, uint64_t n, const uint8_t* validity, llkv_chunk_metadata* out) {
  bool have = false;
  T mn = 0, mx = 0;
  uint64_t nulls = 0;
  std::unordered_set<T> distinct;
  distinct.reserve((size_t)n);
  for (uint64_t i = 0; i < n; ++i) {
    if (validity && !((validity[i >> 3] >> (i & 7)) & 1)) {
      ++nulls;
      continue;
    }
    const T x = v[i];
    distinct.insert(x);
    if (!have || x < mn) mn = x;
    if (!have || x > mx) mx = x;
    have = true;
  }
  out->null_count = nulls;
  if (!have) {
    out->min_val_u64 = out->max_val_u64 = out->distinct_count = 0;
    return LLKV_OK;
  }
  uint64_t bmn = 0, bmx = 0;  // raw bits in the low bytes (the sign extension beyond the type's width is masked by the codec)
  memcpy(&bmn, &mn, sizeof(T));
  memcpy(&bmx, &mx, sizeof(T));
  out->min_val_u64 = llkv_gpu_sortable_u64(prim_type, bmn);
  out->max_val_u64 = llkv_gpu_sortable_u64(prim_type, bmx);
  out->distinct_count = distinct.size();
  return LLKV_OK;
}
template <typename F, typename Bits>
int32_t chunk_stats_float(int32_t prim_type, const F* v, uint64_t n, const uint8_t* validity, llkv_chunk_metadata* out) {
  bool have = false;
  F mn = (F)(1.0 / 0.0), mx = (F)(-1.0 / 0.0);
  uint64_t nulls = 0;
  std::unordered_set<Bits> distinct;
  distinct.reserve((size_t)n);
  for (uint64_t i = 0; i < n; ++i) {
    if (validity && !((validity[i >> 3] >> (i & 7)) & 1)) {
      ++nulls;
      continue;
    }
    have = true;
    Bits b;
    memcpy(&b, &v[i], sizeof(F));
    distinct.insert(b);
    if (v[i] < mn) mn = v[i];  // NaN never passes a strict comparison
    if (v[i] > mx) mx = v[i];
  }
  out->null_count = nulls;
  if (!have) {
    out->min_val_u64 = out->max_val_u64 = out->distinct_count = 0;
    return LLKV_OK;
  }
  Bits bmn, bmx;
  memcpy(&bmn, &mn, sizeof(F));
  memcpy(&bmx, &mx, sizeof(F));
  out->min_val_u64 = llkv_gpu_sortable_u64(prim_type, (uint64_t)bmn);
  out->max_val_u64 = llkv_gpu_sortable_u64(prim_type, (uint64_t)bmx);
  out->distinct_count = distinct.size();
  return LLKV_OK;
}
}  // namespace

extern "C" int32_t llkv_gpu_chunk_stats(int32_t prim_type, const void* values, uint64_t n_rows, const uint8_t* validity, llkv_chunk_metadata* out) {
  if (!out || (n_rows && !values)) return llkv_set_error_message(LLKV_ERR_INVALID_ARGUMENT, "NULL argument");
  if (n_rows == 0) return llkv_set_error_message(LLKV_ERR_NOT_FOUND, "an empty chunk has no statistics");
  switch (prim_type) {
    case LLKV_PT_INT8: return chunk_stats_int(prim_type, static_cast<const int8_t*>(values), n_rows, validity, out);
    case LLKV_PT_INT16: return chunk_stats_int(prim_type, static_cast<const int16_t*>(values), n_rows, validity, out);
    case LLKV_PT_INT32: case LLKV_PT_DATE32: return chunk_stats_int(prim_type, static_cast<const int32_t*>(values), n_rows, validity, out);
    case LLKV_PT_INT64: case LLKV_PT_DATE64: return chunk_stats_int(prim_type, static_cast<const int64_t*>(values), n_rows, validity, out);
    case LLKV_PT_UINT8: return chunk_stats_int(prim_type, static_cast<const uint8_t*>(values), n_rows, validity, out);
    case LLKV_PT_UINT16: return chunk_stats_int(prim_type, static_cast<const uint16_t*>(values), n_rows, validity, out);
    case LLKV_PT_UINT32: return chunk_stats_int(prim_type, static_cast<const uint32_t*>(values), n_rows, validity, out);
    case LLKV_PT_UINT64: return chunk_stats_int(prim_type, static_cast<const uint64_t*>(values), n_rows, validity, out);
    case LLKV_PT_FLOAT32: return chunk_stats_float<float, uint32_t>(prim_type, static_cast<const float*>(values), n_rows, validity, out);
    case LLKV_PT_FLOAT64: return chunk_stats_float<double, uint64_t>(prim_type, static_cast<const double*>(values), n_rows, validity, out);
    default: return llkv_set_error_message(LLKV_ERR_INVALID_ARGUMENT, "no chunk statistics for this column type (compute_chunk_stats returns None)");
  }
}
