// upload.h — host worker pool of the chunk upload path (llkv_gpu_column_append_chunk / _append_blob).
//
// The pager hands out Decimal128 chunks as 16 bytes per value (llkv-column-map/src/serialization.rs:41-53); the
// resident image keeps 8 (or 4) bytes whenever every value of the column is a sign-extended i64 (i32).  When the source
// is page-locked the workers do that narrowing on the host, chunk by chunk, into their own page-locked staging slots and
// DMA the narrow bytes: half (a quarter) of the bytes cross PCIe and the device never sees the wide layout.  A chunk
// that does not fit raises the column's `failed` flag; the runtime then re-uploads the recorded chunks in the Arrow layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace llkv {

struct UploadTicket {  // one per column: completion + failure state of its outstanding jobs
  std::atomic<uint32_t> outstanding{0};
  std::atomic<uint32_t> failed{0};      // some value did not fit the narrow width
  std::atomic<uint32_t> cuda_error{0};  // first cudaError_t a worker met
};

enum UploadKind : int {
  UP_NARROW_D128_I64 = 0,  // 16-byte little-endian i128 -> i64, checked
  UP_NARROW_D128_I32 = 1   // 16-byte little-endian i128 -> i32, checked
};

struct UploadJob {
  UploadTicket* ticket;
  const void* src;
  void* dst;  // device
  uint64_t n_rows;
  int kind;
  unsigned int* d_flag = nullptr;  // DMA share: device word the narrowing kernel raises when a value does not fit
};

// Launches the device-side narrowing of `n_rows` values that crossed the link in the Arrow layout (defined next to the
// kernels, llkv_gpu.cu): wide -> dst, *d_flag |= 1 when a value does not fit.
typedef cudaError_t (*NarrowLaunchFn)(int kind, const void* wide, void* dst, uint64_t n_rows, unsigned int* d_flag, cudaStream_t s);

class UploadPool {
 public:
  UploadPool(int device, int n_threads);
  ~UploadPool();
  int threads() const { return (int)workers_.size(); }
  // splits [src, src + n_rows) into pieces of at most kPieceRows rows and queues them
  void submit(UploadTicket* ticket, const void* src, void* dst, uint64_t n_rows, int kind);
  // The DMA share of a hybrid upload: the rows travel as they lie (16 B/value), in copies of up to kDmaSlotRows rows
  // (consecutive jobs that continue each other are merged), into one of two device staging slots, and are narrowed there.
  // A thread of its own issues them, so the appender never waits for a slot and the workers never wait for the appender.
  void submit_dma(UploadTicket* ticket, const void* src, void* dst, uint64_t n_rows, int kind, unsigned int* d_flag);
  void set_narrow_launcher(NarrowLaunchFn fn) { narrow_launch_ = fn; }
  // blocks until every job of `ticket` has been issued to its worker's stream, then until those streams have drained
  cudaError_t wait(UploadTicket* ticket);
  static constexpr uint64_t kPieceRows = 65536;
  static constexpr uint64_t kDmaSlotRows = 512u << 10;  // 8 MiB of Arrow Decimal128

 private:
  void run(int index);
  void run_dma();
  int device_;
  NarrowLaunchFn narrow_launch_ = nullptr;
  std::thread dma_thread_;
  std::deque<UploadJob> dma_queue_;
  std::condition_variable cv_dma_;
  std::vector<std::thread> workers_;
  std::vector<cudaStream_t> streams_;
  std::mutex mu_;
  std::condition_variable cv_job_, cv_done_;
  std::deque<UploadJob> queue_;
  bool stop_ = false;
  std::atomic<int> ready_{0};
};

// checked narrowing loops (also used directly for sources that are not page-locked: they replace the staging memcpy)
bool narrow_d128_i64(const void* src, void* dst, uint64_t n);
bool narrow_d128_i32(const void* src, void* dst, uint64_t n);
// instruction-set form the two loops dispatch to in this process: 1 = SSE2, 2 = AVX2, 3 = AVX-512 (upload.cpp)
int narrow_isa();

}  // namespace llkv
