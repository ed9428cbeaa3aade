// upload.cpp — see upload.h.
#include "upload.h"

#include <string.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace llkv {

// The loops below are the host side of the PCIe path: 16 bytes read and 8 / 4 written per value.  With SSE2 (every
// x86-64) four values are handled per iteration with unpack instructions; the fit check is an OR-reduction, so the loop
// has no branch.
bool narrow_d128_i64(const void* src, void* dst, uint64_t n) {
  const int64_t* in = static_cast<const int64_t*>(src);
  int64_t* out = static_cast<int64_t*>(dst);
  int64_t bad = 0;
  uint64_t i = 0;
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
  {
    int64_t t[2];
    _mm_storeu_si128(reinterpret_cast<__m128i*>(t), vbad);
    bad = t[0] | t[1];
  }
#endif
  for (; i < n; ++i) {
    const int64_t lo = in[2 * i], hi = in[2 * i + 1];
    out[i] = lo;
    bad |= hi ^ (lo >> 63);
  }
  return bad == 0;
}

bool narrow_d128_i32(const void* src, void* dst, uint64_t n) {
  const int64_t* in = static_cast<const int64_t*>(src);
  int32_t* out = static_cast<int32_t*>(dst);
  int64_t bad = 0;
  uint64_t i = 0;
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
  {
    int64_t t[2];
    _mm_storeu_si128(reinterpret_cast<__m128i*>(t), vbad);
    bad = t[0] | t[1];
  }
#endif
  for (; i < n; ++i) {
    const int64_t lo = in[2 * i], hi = in[2 * i + 1];
    out[i] = (int32_t)lo;
    bad |= (hi ^ (lo >> 63)) | (lo ^ (int64_t)(int32_t)lo);
  }
  return bad == 0;
}

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
