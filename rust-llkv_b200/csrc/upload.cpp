// upload.cpp — see upload.h.
#include "upload.h"

#include <string.h>

namespace llkv {

bool narrow_d128_i64(const void* src, void* dst, uint64_t n) {
  const int64_t* in = static_cast<const int64_t*>(src);
  int64_t* out = static_cast<int64_t*>(dst);
  int64_t bad = 0;
  for (uint64_t i = 0; i < n; ++i) {
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
  for (uint64_t i = 0; i < n; ++i) {
    const int64_t lo = in[2 * i], hi = in[2 * i + 1];
    out[i] = (int32_t)lo;
    bad |= (hi ^ (lo >> 63)) | (lo ^ (int64_t)(int32_t)lo);
  }
  return bad == 0;
}

UploadPool::UploadPool(int device, int n_threads) : device_(device) {
  if (n_threads < 1) n_threads = 1;
  streams_.assign((size_t)n_threads, nullptr);
  for (int i = 0; i < n_threads; ++i) workers_.emplace_back([this, i] { run(i); });
  // the streams exist before the first wait()
  std::unique_lock<std::mutex> lk(mu_);
  cv_done_.wait(lk, [&] { return ready_.load() == (int)workers_.size(); });
}

UploadPool::~UploadPool() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_job_.notify_all();
  for (std::thread& t : workers_) t.join();
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
