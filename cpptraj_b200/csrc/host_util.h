// host_util.h -- host-side helpers of the C ABI: a small copy pool and page hints.
//
// cpptraj hands the library PAGEABLE buffers: COORDS is a std::vector<float> (src/CompactFrameArray.h), the result
// a new float[] inside Matrix<float> (src/Matrix.h:196-233) that nobody has touched yet.  Measured on the B200 box
// (tools/microbench/host_path_probe.cu, profiles/r2_host_path_probe.txt): cudaHostRegister pins such memory at
// 3.6 GB/s (fresh) / 11.8 GB/s (touched), the driver's own pageable D2H copy reaches 2.2 GB/s into fresh pages and
// one memcpy thread 2.3 GB/s (page faults), 14.5 GB/s into touched ones -- against 55 GB/s of PCIe.  Sixteen memcpy
// threads between a pinned stage and pageable memory reach 29-38 GB/s into fresh pages (38 with transparent huge
// pages requested) and ~80 GB/s out of touched ones.  So the pageable paths stage through pinned ring slots and
// move the bytes with a pool of host threads, overlapped with the transfers and kernels of the neighbouring slots.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#ifdef __linux__
#include <sys/mman.h>
#endif

namespace b200 {

/// Fork-join pool: run(pieces, fn) calls fn(0..pieces-1) on the workers and the calling thread and returns when
/// all are done.  One pool per device (the per-device host threads of a multi-device call never share one).
class CopyPool {
 public:
  ~CopyPool() { stop(); }
  /// (Re)size to n threads in total, the caller included.
  void ensure(int n) {
    n = std::max(1, n);
    if ((int)workers_.size() == n - 1) return;
    stop();
    stop_ = false;
    for (int t = 0; t < n - 1; ++t) workers_.emplace_back([this] { loop(); });
  }
  void stop() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cvGo_.notify_all();
    for (auto& t : workers_) t.join();
    workers_.clear();
  }
  int threads() const { return (int)workers_.size() + 1; }
  void run(size_t pieces, const std::function<void(size_t)>& fn) {
    if (workers_.empty() || pieces <= 1) {
      for (size_t i = 0; i < pieces; ++i) fn(i);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn; pieces_ = pieces; next_.store(0); checkedIn_ = 0; ++gen_;
    }
    cvGo_.notify_all();
    size_t i;
    while ((i = next_.fetch_add(1)) < pieces) fn(i);
    std::unique_lock<std::mutex> lk(mu_);
    cvDone_.wait(lk, [&] { return checkedIn_ == workers_.size(); });
    fn_ = nullptr;
  }
  /// dst[r*dPitch .. +rowBytes) = src[r*sPitch .. +rowBytes) for r < rows (pitches in bytes), ~1 MiB per piece.
  void copy2d(void* dst, size_t dPitch, const void* src, size_t sPitch, size_t rowBytes, size_t rows) {
    if (!rows || !rowBytes) return;
    char* d = (char*)dst;
    const char* s = (const char*)src;
    if (dPitch == rowBytes && sPitch == rowBytes) {   // contiguous
      const size_t n = rowBytes * rows, piece = (size_t)1 << 20;
      const size_t np = (n + piece - 1) / piece;
      run(np, [=](size_t p) { const size_t a = p * piece; std::memcpy(d + a, s + a, std::min(piece, n - a)); });
      return;
    }
    const size_t rpp = std::max<size_t>(1, ((size_t)1 << 20) / rowBytes);   // rows per piece
    const size_t np = (rows + rpp - 1) / rpp;
    run(np, [=](size_t p) {
      const size_t r1 = std::min(rows, (p + 1) * rpp);
      for (size_t r = p * rpp; r < r1; ++r) std::memcpy(d + r * dPitch, s + r * sPitch, rowBytes);
    });
  }
  void copy(void* dst, const void* src, size_t bytes) { copy2d(dst, bytes, src, bytes, bytes, 1); }

 private:
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(size_t)>* fn;
      size_t pieces;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cvGo_.wait(lk, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_; fn = fn_; pieces = pieces_;
      }
      size_t i;
      while ((i = next_.fetch_add(1)) < pieces) (*fn)(i);
      {
        std::lock_guard<std::mutex> lk(mu_);
        ++checkedIn_;
      }
      cvDone_.notify_one();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cvGo_, cvDone_;
  const std::function<void(size_t)>* fn_ = nullptr;
  size_t pieces_ = 0, checkedIn_ = 0;
  std::atomic<size_t> next_{0};
  uint64_t gen_ = 0;
  bool stop_ = false;
};

/// Host threads per device for the staged copies: env B200_HOST_THREADS, else the hardware threads shared out over
/// the devices, at most 16 (the probe's memcpy rates flatten there).
inline int host_threads_per_device(int nDevices) {
  if (const char* e = getenv("B200_HOST_THREADS")) return std::max(1, atoi(e));
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  return (int)std::min<unsigned>(16u, std::max(1u, hw / (unsigned)std::max(1, nDevices)));
}

/// Ask for transparent huge pages on a range the library is about to fill for the first time (the caller's fresh
/// result buffer): first-touch faults then come 2 MiB at a time.  Harmless where THP is off or the range is in use.
inline void advise_hugepages(void* p, size_t bytes) {
#if defined(__linux__) && defined(MADV_HUGEPAGE)
  const uintptr_t two = (uintptr_t)2 << 20;
  const uintptr_t a = ((uintptr_t)p + two - 1) & ~(two - 1), b = ((uintptr_t)p + bytes) & ~(two - 1);
  if (b > a) (void)madvise((void*)a, b - a, MADV_HUGEPAGE);
#else
  (void)p; (void)bytes;
#endif
}

/// Populate (fault in, writable) a range the library is about to fill, in the kernel and in bulk
/// (MADV_POPULATE_WRITE, Linux >= 5.14): a fresh 4 KiB page costs ~1.7 us when it is faulted by the first store to it
/// (2.3 GB/s per thread on the B200 box), most of it trap overhead that a bulk populate does not pay.
/// \return false where unsupported (the first stores fault the pages in, as before).
inline bool populate_pages(void* p, size_t bytes) {
#if defined(__linux__)
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
  const uintptr_t pg = 4096;
  const uintptr_t a = (uintptr_t)p & ~(pg - 1), b = ((uintptr_t)p + bytes + pg - 1) & ~(pg - 1);
  return b > a && madvise((void*)a, b - a, MADV_POPULATE_WRITE) == 0;
#else
  (void)p; (void)bytes;
  return false;
#endif
}

}  // namespace b200
