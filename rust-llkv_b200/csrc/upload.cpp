// upload.cpp — see upload.h.
#include "upload.h"

#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace llkv {

// The loops below are the host side of the PCIe path: 16 bytes read and 8 / 4 written per value, with a branch-free fit
// check (an OR-reduction over "the upper words repeat the sign").  One thread of the SSE2 form moves ~5.5 GB/s of Arrow
// bytes on the hosts of this pool — the end-to-end step was bound by exactly that (15 workers x 6 GB/s) — so the loop is
// dispatched once per process on what the CPU has: AVX-512 (two 64-byte loads and one cross-lane permute per eight
// values) or AVX2, both with a software prefetch 2 KB ahead (a single core's demand misses alone do not keep enough
// lines in flight); ~12 GB/s per thread.  Every form gives bit-identical output and the same verdict
// (tests/test_host_logic.py runs them against each other).
namespace {

enum { ISA_SSE2 = 1, ISA_AVX2 = 2, ISA_AVX512 = 3 };
constexpr int kPrefetchAhead = 2048;

inline bool tail_i64(const int64_t* in, int64_t* out, uint64_t i, uint64_t n) {
  int64_t bad = 0;
  for (; i < n; ++i) {
    const int64_t lo = in[2 * i], hi = in[2 * i + 1];
    out[i] = lo;
    bad |= hi ^ (lo >> 63);
  }
  return bad == 0;
}
inline bool tail_i32(const int64_t* in, int32_t* out, uint64_t i, uint64_t n) {
  int64_t bad = 0;
  for (; i < n; ++i) {
    const int64_t lo = in[2 * i], hi = in[2 * i + 1];
    out[i] = (int32_t)lo;
    bad |= (hi ^ (lo >> 63)) | (lo ^ (int64_t)(int32_t)lo);
  }
  return bad == 0;
}

bool narrow_i64_sse2(const void* src, void* dst, uint64_t n) {
  const int64_t* in = static_cast<const int64_t*>(src);
  int64_t* out = static_cast<int64_t*>(dst);
  uint64_t i = 0;
  bool ok = true;
#if defined(__SSE2__)
  __m128i vbad = _mm_setzero_si128();
  for (; i + 4 <= n; i += 4) {
    const __m128i r0 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(in + 2 * i));
    const __m128i r1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(in + 2 * i + 2));
    const __m128i r2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(in + 2 * i + 4));
    const __m128i r3 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(in + 2 * i + 6));
    const __m128i lo01 = _mm_unpacklo_epi64(r0, r1), hi01 = _mm_unpackhi_epi64(r0, r1);
    const __m128i lo23 = _mm_unpacklo_epi64(r2, r3), hi23 = _mm_unpackhi_epi64(r2, r3);
    _mm_storeu_si128(reinterpret_cast<__m128i*>(out + i), lo01);
    _mm_storeu_si128(reinterpret_cast<__m128i*>(out + i + 2), lo23);
    const __m128i s01 = _mm_shuffle_epi32(_mm_srai_epi32(lo01, 31), _MM_SHUFFLE(3, 3, 1, 1));
    const __m128i s23 = _mm_shuffle_epi32(_mm_srai_epi32(lo23, 31), _MM_SHUFFLE(3, 3, 1, 1));
    vbad = _mm_or_si128(vbad, _mm_or_si128(_mm_xor_si128(hi01, s01), _mm_xor_si128(hi23, s23)));
  }
  int64_t t[2];
  _mm_storeu_si128(reinterpret_cast<__m128i*>(t), vbad);
  ok = (t[0] | t[1]) == 0;
#endif
  return tail_i64(in, out, i, n) && ok;
}

bool narrow_i32_sse2(const void* src, void* dst, uint64_t n) {
  const int64_t* in = static_cast<const int64_t*>(src);
  int32_t* out = static_cast<int32_t*>(dst);
  uint64_t i = 0;
  bool ok = true;
#if defined(__SSE2__)
  __m128i vbad = _mm_setzero_si128();
  const __m128i upper = _mm_set_epi32(-1, -1, -1, 0);  // dwords 1..3 of a value
  for (; i + 4 <= n; i += 4) {
    const __m128i r0 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(in + 2 * i));
    const __m128i r1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(in + 2 * i + 2));
    const __m128i r2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(in + 2 * i + 4));
    const __m128i r3 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(in + 2 * i + 6));
    // dword 0 of the four values
    const __m128i a = _mm_unpacklo_epi32(r0, r1), b = _mm_unpacklo_epi32(r2, r3);
    _mm_storeu_si128(reinterpret_cast<__m128i*>(out + i), _mm_unpacklo_epi64(a, b));
    // dwords 1..3 must repeat the sign of dword 0
    const __m128i x0 = _mm_xor_si128(r0, _mm_shuffle_epi32(_mm_srai_epi32(r0, 31), 0));
    const __m128i x1 = _mm_xor_si128(r1, _mm_shuffle_epi32(_mm_srai_epi32(r1, 31), 0));
    const __m128i x2 = _mm_xor_si128(r2, _mm_shuffle_epi32(_mm_srai_epi32(r2, 31), 0));
    const __m128i x3 = _mm_xor_si128(r3, _mm_shuffle_epi32(_mm_srai_epi32(r3, 31), 0));
    vbad = _mm_or_si128(vbad, _mm_and_si128(upper, _mm_or_si128(_mm_or_si128(x0, x1), _mm_or_si128(x2, x3))));
  }
  int64_t t[2];
  _mm_storeu_si128(reinterpret_cast<__m128i*>(t), vbad);
  ok = (t[0] | t[1]) == 0;
#endif
  return tail_i32(in, out, i, n) && ok;
}

#if defined(__x86_64__) && defined(__GNUC__)
#define LLKV_HAVE_WIDE_NARROW 1
__attribute__((target("avx2"))) bool narrow_i32_avx2(const void* src, void* dst, uint64_t n) {
  const char* in = static_cast<const char*>(src);
  int32_t* out = static_cast<int32_t*>(dst);
  uint64_t i = 0;
  __m256i bad = _mm256_setzero_si256();
  const __m256i upper = _mm256_setr_epi32(0, -1, -1, -1, 0, -1, -1, -1);
  const __m256i order = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
  for (; i + 8 <= n; i += 8) {
    _mm_prefetch(in + i * 16 + kPrefetchAhead, _MM_HINT_T0);
    _mm_prefetch(in + i * 16 + kPrefetchAhead + 64, _MM_HINT_T0);
    const __m256i y0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i * 16));        // values 0, 1
    const __m256i y1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i * 16 + 32));   // values 2, 3
    const __m256i y2 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i * 16 + 64));   // values 4, 5
    const __m256i y3 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i * 16 + 96));   // values 6, 7
    const __m256i a = _mm256_unpacklo_epi32(y0, y1), b = _mm256_unpacklo_epi32(y2, y3);  // dword 0 of (0,2 | 1,3), (4,6 | 5,7)
    const __m256i c = _mm256_unpacklo_epi64(a, b);                                        // values 0 2 4 6 | 1 3 5 7
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(out + i), _mm256_permutevar8x32_epi32(c, order));
    const __m256i x0 = _mm256_xor_si256(y0, _mm256_shuffle_epi32(_mm256_srai_epi32(y0, 31), 0));
    const __m256i x1 = _mm256_xor_si256(y1, _mm256_shuffle_epi32(_mm256_srai_epi32(y1, 31), 0));
    const __m256i x2 = _mm256_xor_si256(y2, _mm256_shuffle_epi32(_mm256_srai_epi32(y2, 31), 0));
    const __m256i x3 = _mm256_xor_si256(y3, _mm256_shuffle_epi32(_mm256_srai_epi32(y3, 31), 0));
    bad = _mm256_or_si256(bad, _mm256_or_si256(_mm256_or_si256(x0, x1), _mm256_or_si256(x2, x3)));
  }
  const bool ok = _mm256_testz_si256(bad, upper) != 0;
  return tail_i32(static_cast<const int64_t*>(src), out, i, n) && ok;
}

__attribute__((target("avx2"))) bool narrow_i64_avx2(const void* src, void* dst, uint64_t n) {
  const char* in = static_cast<const char*>(src);
  int64_t* out = static_cast<int64_t*>(dst);
  uint64_t i = 0;
  __m256i bad = _mm256_setzero_si256();
  const __m256i upper = _mm256_setr_epi32(0, 0, -1, -1, 0, 0, -1, -1);  // the high qword of a value
  for (; i + 8 <= n; i += 8) {
    _mm_prefetch(in + i * 16 + kPrefetchAhead, _MM_HINT_T0);
    _mm_prefetch(in + i * 16 + kPrefetchAhead + 64, _MM_HINT_T0);
    const __m256i y0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i * 16));
    const __m256i y1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i * 16 + 32));
    const __m256i y2 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i * 16 + 64));
    const __m256i y3 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i * 16 + 96));
    // low qwords: (0, 2 | 1, 3) -> 0 1 2 3
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(out + i), _mm256_permute4x64_epi64(_mm256_unpacklo_epi64(y0, y1), 0xD8));
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(out + i + 4), _mm256_permute4x64_epi64(_mm256_unpacklo_epi64(y2, y3), 0xD8));
    // the high qword must repeat the sign of the low one (the sign sits in dword 1)
    const __m256i x0 = _mm256_xor_si256(y0, _mm256_shuffle_epi32(_mm256_srai_epi32(y0, 31), _MM_SHUFFLE(1, 1, 1, 1)));
    const __m256i x1 = _mm256_xor_si256(y1, _mm256_shuffle_epi32(_mm256_srai_epi32(y1, 31), _MM_SHUFFLE(1, 1, 1, 1)));
    const __m256i x2 = _mm256_xor_si256(y2, _mm256_shuffle_epi32(_mm256_srai_epi32(y2, 31), _MM_SHUFFLE(1, 1, 1, 1)));
    const __m256i x3 = _mm256_xor_si256(y3, _mm256_shuffle_epi32(_mm256_srai_epi32(y3, 31), _MM_SHUFFLE(1, 1, 1, 1)));
    bad = _mm256_or_si256(bad, _mm256_or_si256(_mm256_or_si256(x0, x1), _mm256_or_si256(x2, x3)));
  }
  const bool ok = _mm256_testz_si256(bad, upper) != 0;
  return tail_i64(static_cast<const int64_t*>(src), out, i, n) && ok;
}

__attribute__((target("avx512f,avx512bw,avx512vl"))) bool narrow_i32_avx512(const void* src, void* dst, uint64_t n) {
  const char* in = static_cast<const char*>(src);
  int32_t* out = static_cast<int32_t*>(dst);
  uint64_t i = 0;
  const __m512i first = _mm512_setr_epi32(0, 4, 8, 12, 16, 20, 24, 28, 0, 0, 0, 0, 0, 0, 0, 0);  // dword 0 of eight values in two registers
  __m512i bad0 = _mm512_setzero_si512(), bad1 = bad0;
  for (; i + 16 <= n; i += 16) {
    _mm_prefetch(in + i * 16 + kPrefetchAhead, _MM_HINT_T0);
    _mm_prefetch(in + i * 16 + kPrefetchAhead + 64, _MM_HINT_T0);
    _mm_prefetch(in + i * 16 + kPrefetchAhead + 128, _MM_HINT_T0);
    _mm_prefetch(in + i * 16 + kPrefetchAhead + 192, _MM_HINT_T0);
    const __m512i z0 = _mm512_loadu_si512(in + i * 16), z1 = _mm512_loadu_si512(in + i * 16 + 64);
    const __m512i z2 = _mm512_loadu_si512(in + i * 16 + 128), z3 = _mm512_loadu_si512(in + i * 16 + 192);
    const __m512i v0 = _mm512_permutex2var_epi32(z0, first, z1), v1 = _mm512_permutex2var_epi32(z2, first, z3);
    _mm512_storeu_si512(out + i, _mm512_inserti64x4(v0, _mm512_castsi512_si256(v1), 1));
    // dwords 1..3 must repeat the sign of dword 0
    const __m512i x0 = _mm512_xor_si512(z0, _mm512_shuffle_epi32(_mm512_srai_epi32(z0, 31), (_MM_PERM_ENUM)0x00));
    const __m512i x1 = _mm512_xor_si512(z1, _mm512_shuffle_epi32(_mm512_srai_epi32(z1, 31), (_MM_PERM_ENUM)0x00));
    const __m512i x2 = _mm512_xor_si512(z2, _mm512_shuffle_epi32(_mm512_srai_epi32(z2, 31), (_MM_PERM_ENUM)0x00));
    const __m512i x3 = _mm512_xor_si512(z3, _mm512_shuffle_epi32(_mm512_srai_epi32(z3, 31), (_MM_PERM_ENUM)0x00));
    bad0 = _mm512_or_si512(bad0, _mm512_or_si512(x0, x1));
    bad1 = _mm512_or_si512(bad1, _mm512_or_si512(x2, x3));
  }
  bad0 = _mm512_or_si512(bad0, bad1);
  const bool ok = _mm512_mask_test_epi32_mask((__mmask16)0xEEEE, bad0, bad0) == 0;
  return tail_i32(static_cast<const int64_t*>(src), out, i, n) && ok;
}

__attribute__((target("avx512f,avx512bw,avx512vl"))) bool narrow_i64_avx512(const void* src, void* dst, uint64_t n) {
  const char* in = static_cast<const char*>(src);
  int64_t* out = static_cast<int64_t*>(dst);
  uint64_t i = 0;
  const __m512i low = _mm512_setr_epi64(0, 2, 4, 6, 8, 10, 12, 14);  // the low qword of eight values in two registers
  __m512i bad0 = _mm512_setzero_si512(), bad1 = bad0;
  for (; i + 16 <= n; i += 16) {
    _mm_prefetch(in + i * 16 + kPrefetchAhead, _MM_HINT_T0);
    _mm_prefetch(in + i * 16 + kPrefetchAhead + 64, _MM_HINT_T0);
    _mm_prefetch(in + i * 16 + kPrefetchAhead + 128, _MM_HINT_T0);
    _mm_prefetch(in + i * 16 + kPrefetchAhead + 192, _MM_HINT_T0);
    const __m512i z0 = _mm512_loadu_si512(in + i * 16), z1 = _mm512_loadu_si512(in + i * 16 + 64);
    const __m512i z2 = _mm512_loadu_si512(in + i * 16 + 128), z3 = _mm512_loadu_si512(in + i * 16 + 192);
    _mm512_storeu_si512(out + i, _mm512_permutex2var_epi64(z0, low, z1));
    _mm512_storeu_si512(out + i + 8, _mm512_permutex2var_epi64(z2, low, z3));
    // the high qword must repeat the sign of the low one: (sign, sign) of qword 0 copied over qword 1
    const __m512i x0 = _mm512_xor_si512(z0, _mm512_shuffle_epi32(_mm512_srai_epi64(z0, 63), (_MM_PERM_ENUM)0x44));
    const __m512i x1 = _mm512_xor_si512(z1, _mm512_shuffle_epi32(_mm512_srai_epi64(z1, 63), (_MM_PERM_ENUM)0x44));
    const __m512i x2 = _mm512_xor_si512(z2, _mm512_shuffle_epi32(_mm512_srai_epi64(z2, 63), (_MM_PERM_ENUM)0x44));
    const __m512i x3 = _mm512_xor_si512(z3, _mm512_shuffle_epi32(_mm512_srai_epi64(z3, 63), (_MM_PERM_ENUM)0x44));
    bad0 = _mm512_or_si512(bad0, _mm512_or_si512(x0, x1));
    bad1 = _mm512_or_si512(bad1, _mm512_or_si512(x2, x3));
  }
  bad0 = _mm512_or_si512(bad0, bad1);
  const bool ok = _mm512_mask_test_epi64_mask((__mmask8)0xAA, bad0, bad0) == 0;
  return tail_i64(static_cast<const int64_t*>(src), out, i, n) && ok;
}
#endif

// best form this CPU (and the OS: the wide register state must be enabled) runs; LLKV_GPU_HOST_ISA=sse2|avx2|avx512 caps it
int host_isa() {
  static const int level = [] {
    int have = ISA_SSE2;
#ifdef LLKV_HAVE_WIDE_NARROW
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx2")) have = ISA_AVX2;
    if (have == ISA_AVX2 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl"))
      have = ISA_AVX512;
#endif
    if (const char* e = getenv("LLKV_GPU_HOST_ISA")) {
      const int cap = !strcmp(e, "sse2") ? ISA_SSE2 : !strcmp(e, "avx2") ? ISA_AVX2 : ISA_AVX512;
      if (cap < have) have = cap;
    }
    return have;
  }();
  return level;
}

bool narrow_with(int isa, int out_width, const void* src, void* dst, uint64_t n) {
#ifdef LLKV_HAVE_WIDE_NARROW
  if (isa == ISA_AVX512) return out_width == 4 ? narrow_i32_avx512(src, dst, n) : narrow_i64_avx512(src, dst, n);
  if (isa == ISA_AVX2) return out_width == 4 ? narrow_i32_avx2(src, dst, n) : narrow_i64_avx2(src, dst, n);
#endif
  return out_width == 4 ? narrow_i32_sse2(src, dst, n) : narrow_i64_sse2(src, dst, n);
}

}  // namespace

int narrow_isa() { return host_isa(); }
bool narrow_d128_i64(const void* src, void* dst, uint64_t n) { return narrow_with(host_isa(), 8, src, dst, n); }
bool narrow_d128_i32(const void* src, void* dst, uint64_t n) { return narrow_with(host_isa(), 4, src, dst, n); }

UploadPool::UploadPool(int device, int n_threads) : device_(device) {
  if (n_threads < 1) n_threads = 1;
  streams_.assign((size_t)n_threads + 1, nullptr);  // (the last one belongs to the DMA thread)
  for (int i = 0; i < n_threads; ++i) workers_.emplace_back([this, i] { run(i); });
  dma_thread_ = std::thread([this] { run_dma(); });
  // the streams exist before the first wait()
  std::unique_lock<std::mutex> lk(mu_);
  cv_done_.wait(lk, [&] { return ready_.load() == (int)workers_.size() + 1; });
}

UploadPool::~UploadPool() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_job_.notify_all();
  cv_dma_.notify_all();
  for (std::thread& t : workers_) t.join();
  dma_thread_.join();
}

void UploadPool::submit_dma(UploadTicket* ticket, const void* src, void* dst, uint64_t n_rows, int kind, unsigned int* d_flag) {
  if (!n_rows) return;
  {
    std::lock_guard<std::mutex> lk(mu_);
    ticket->outstanding.fetch_add(1);
    UploadJob j{ticket, src, dst, n_rows, kind};
    j.d_flag = d_flag;
    dma_queue_.push_back(j);
  }
  cv_dma_.notify_one();
}

void UploadPool::run_dma() {
  cudaError_t e = cudaSetDevice(device_);
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool used[2] = {false, false};
  void* slot[2] = {nullptr, nullptr};
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
  const cudaError_t init_error = e;
  {
    std::lock_guard<std::mutex> lk(mu_);
    streams_.back() = stream;
    ready_.fetch_add(1);
  }
  cv_done_.notify_all();
  unsigned turn = 0;
  std::vector<UploadJob> run;  // jobs that travel as one copy
  for (;;) {
    run.clear();
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_dma_.wait(lk, [&] { return stop_ || !dma_queue_.empty(); });
      if (dma_queue_.empty()) break;  // stop_
      run.push_back(dma_queue_.front());
      dma_queue_.pop_front();
      uint64_t rows = run[0].n_rows;
      const uint64_t out_w = run[0].kind == UP_NARROW_D128_I32 ? 4 : 8;
      while (!dma_queue_.empty()) {  // the next job continues this one on both sides: one copy
        const UploadJob& n = dma_queue_.front();
        const UploadJob& l = run.back();
        if (n.kind != l.kind || n.d_flag != l.d_flag || static_cast<const char*>(l.src) + l.n_rows * 16 != n.src ||
            static_cast<char*>(l.dst) + l.n_rows * out_w != n.dst || rows + n.n_rows > kDmaSlotRows)
          break;
        rows += n.n_rows;
        run.push_back(n);
        dma_queue_.pop_front();
      }
    }
    cudaError_t je = init_error;
    uint64_t rows = 0;
    for (const UploadJob& j : run) rows += j.n_rows;
    const UploadJob& first = run[0];
    uint64_t done = 0;
    while (je == cudaSuccess && done < rows) {  // (a single job larger than a slot goes in slices)
      const uint64_t take = rows - done < kDmaSlotRows ? rows - done : kDmaSlotRows;
      const unsigned s = turn++ & 1u;
      if (!slot[s]) je = cudaMalloc(&slot[s], kDmaSlotRows * 16);
      if (je == cudaSuccess && used[s]) je = cudaEventSynchronize(ev[s]);  // the kernel that read the slot last is done
      const uint64_t out_w = first.kind == UP_NARROW_D128_I32 ? 4 : 8;
      if (je == cudaSuccess) je = cudaMemcpyAsync(slot[s], static_cast<const char*>(first.src) + done * 16, take * 16, cudaMemcpyHostToDevice, stream);
      if (je == cudaSuccess) je = narrow_launch_ ? narrow_launch_(first.kind, slot[s], static_cast<char*>(first.dst) + done * out_w, take, first.d_flag, stream) : cudaErrorNotSupported;
      if (je == cudaSuccess) je = cudaEventRecord(ev[s], stream);
      used[s] = true;
      done += take;
    }
    for (const UploadJob& j : run) {
      if (je != cudaSuccess) {
        uint32_t expected = 0;
        j.ticket->cuda_error.compare_exchange_strong(expected, (uint32_t)je);
      }
      std::lock_guard<std::mutex> lk(mu_);
      j.ticket->outstanding.fetch_sub(1);
    }
    cv_done_.notify_all();
  }
  if (stream) cudaStreamSynchronize(stream);
  for (int i = 0; i < 2; ++i) {
    if (slot[i]) cudaFree(slot[i]);
    if (ev[i]) cudaEventDestroy(ev[i]);
  }
  if (stream) cudaStreamDestroy(stream);
}

void UploadPool::submit(UploadTicket* ticket, const void* src, void* dst, uint64_t n_rows, int kind) {
  const uint64_t out_w = kind == UP_NARROW_D128_I32 ? 4 : 8;
  uint32_t pieces = 0;
  {
    std::lock_guard<std::mutex> lk(mu_);
    for (uint64_t lo = 0; lo < n_rows; lo += kPieceRows) {
      const uint64_t m = n_rows - lo < kPieceRows ? n_rows - lo : kPieceRows;
      ticket->outstanding.fetch_add(1);
      queue_.push_back(UploadJob{ticket, static_cast<const char*>(src) + lo * 16, static_cast<char*>(dst) + lo * out_w, m, kind});
      ++pieces;
    }
  }
  if (pieces == 1) cv_job_.notify_one();
  else if (pieces) cv_job_.notify_all();
}

cudaError_t UploadPool::wait(UploadTicket* ticket) {
  {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return ticket->outstanding.load() == 0; });
  }
  cudaError_t first = (cudaError_t)ticket->cuda_error.load();
  for (cudaStream_t s : streams_) {
    if (!s) continue;
    const cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess && first == cudaSuccess) first = e;
  }
  return first;
}

void UploadPool::run(int index) {
  cudaError_t e = cudaSetDevice(device_);
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  unsigned char* stage = nullptr;
  const size_t slot = (size_t)kPieceRows * 8;
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&stage, 2 * slot, cudaHostAllocDefault);
  const cudaError_t init_error = e;
  {
    std::lock_guard<std::mutex> lk(mu_);
    streams_[(size_t)index] = stream;
    ready_.fetch_add(1);
  }
  cv_done_.notify_all();
  unsigned turn = 0;
  for (;;) {
    UploadJob job;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_job_.wait(lk, [&] { return stop_ || !queue_.empty(); });
      if (queue_.empty()) break;  // stop_
      job = queue_.front();
      queue_.pop_front();
    }
    cudaError_t je = init_error;
    if (je == cudaSuccess) {
      const unsigned s = turn++ & 1u;
      je = cudaEventSynchronize(ev[s]);  // the slot's previous copy has drained
      unsigned char* out = stage + s * slot;
      const bool ok = job.kind == UP_NARROW_D128_I32 ? narrow_d128_i32(job.src, out, job.n_rows) : narrow_d128_i64(job.src, out, job.n_rows);
      if (!ok) job.ticket->failed.store(1);
      const size_t bytes = (size_t)job.n_rows * (job.kind == UP_NARROW_D128_I32 ? 4 : 8);
      if (je == cudaSuccess) je = cudaMemcpyAsync(job.dst, out, bytes, cudaMemcpyHostToDevice, stream);
      if (je == cudaSuccess) je = cudaEventRecord(ev[s], stream);
    }
    if (je != cudaSuccess) {
      uint32_t expected = 0;
      job.ticket->cuda_error.compare_exchange_strong(expected, (uint32_t)je);
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      job.ticket->outstanding.fetch_sub(1);
    }
    cv_done_.notify_all();
  }
  if (stream) cudaStreamSynchronize(stream);
  if (stage) cudaFreeHost(stage);
  if (ev[0]) cudaEventDestroy(ev[0]);
  if (ev[1]) cudaEventDestroy(ev[1]);
  if (stream) cudaStreamDestroy(stream);
}

}  // namespace llkv

// Test hook (not part of include/llkv_gpu.h): one narrowing loop in a chosen instruction-set form.  isa: 0 = what the
// process dispatches to, 1 = SSE2, 2 = AVX2, 3 = AVX-512.  Returns 1 = every value fits, 0 = some value does not fit
// (dst holds the truncated values), -1 = this CPU does not have that form.
extern "C" int llkv_internal_narrow_d128(int out_width, int isa, const void* src, void* dst, uint64_t n) {
  if (out_width != 4 && out_width != 8) return -1;
  int have = llkv::ISA_SSE2;
#ifdef LLKV_HAVE_WIDE_NARROW
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx2")) have = llkv::ISA_AVX2;
  if (have == llkv::ISA_AVX2 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl"))
    have = llkv::ISA_AVX512;
#endif
  if (isa == 0) isa = llkv::host_isa();
  if (isa < llkv::ISA_SSE2 || isa > have) return -1;
  return llkv::narrow_with(isa, out_width, src, dst, n) ? 1 : 0;
}
