// host_widen.hpp — u32 -> usize widening of downloaded CSR index arrays on host threads.
//
// The reference's data contract is `usize` indices (nalgebra_sparse::CsrMatrix, handed to faer as-is:
// formoniq/src/linalg/faer.rs:16-24); the device holds u32.  Widening on the device doubles the bytes that cross PCIe,
// which is what bounds a one-shot assembly end to end.  Here the u32 array is copied into the UPPER half of the
// caller's u64 buffer and widened in place by a small thread pool while the copy stream already moves the next array:
//   dst[i] (bytes [8i, 8i + 8)) <- src[i] (bytes [4n + 4i, 4n + 4i + 4))
// A write never lands on a source element that is still needed if the elements are processed in waves
// [lo, hi), hi = (n + lo) / 2: the bytes a wave writes, [8 lo, 8 hi), end at or before 4n + 4 lo, i.e. they hold only
// source elements < lo, which earlier waves have consumed.  Inside a wave every element is independent (threads split
// it); the waves halve, and the last few elements are done front to back by one thread.
#pragma once
#if defined(__SSE2__) && defined(__x86_64__)
#include <emmintrin.h>
#endif
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace fq {

class HostWidener {
 public:
  explicit HostWidener(int nthreads) : nthreads_(nthreads < 1 ? 1 : nthreads) {
    for (int t = 0; t < nthreads_; ++t) workers_.emplace_back([this, t] { worker(t); });
    driver_ = std::thread([this] { drive(); });
  }
  ~HostWidener() {
    wait_idle();  // queued jobs need the workers
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
    }
    cv_job_.notify_all();
    cv_wave_.notify_all();
    driver_.join();
    for (auto& w : workers_) w.join();
  }
  // Called from a CUDA host callback once the n u32 values sit at ((uint32_t*)buf) + n: no CUDA calls in here.
  void enqueue(uint64_t* buf, size_t n) {
    {
      std::lock_guard<std::mutex> g(m_);
      jobs_.push_back({buf, n});
      ++pending_;
    }
    cv_job_.notify_all();
  }
  void wait_idle() {
    std::unique_lock<std::mutex> g(m_);
    cv_idle_.wait(g, [this] { return pending_ == 0; });
  }
  int threads() const { return nthreads_; }

  // the serial definition (also the tail of the parallel one); exposed for the tests
  // (source and destination overlap here and have different types: may_alias keeps the loads and stores in order)
  typedef uint32_t __attribute__((may_alias)) u32_alias;
  typedef uint64_t __attribute__((may_alias)) u64_alias;
  static void widen_serial(uint64_t* buf, size_t n, size_t from = 0) {
    const volatile u32_alias* src = reinterpret_cast<const volatile u32_alias*>(buf) + n;
    volatile u64_alias* dst = reinterpret_cast<volatile u64_alias*>(buf);
    for (size_t i = from; i < n; ++i) {
      const uint64_t v = src[i];
      dst[i] = v;
    }
  }

 private:
  struct Job {
    uint64_t* buf;
    size_t n;
  };
  // The destination of a wave is written once and not read again by the pool: streaming stores skip the read-for-
  // ownership of 8 bytes per index that ordinary stores would add to the host memory traffic (which the DMA engines of
  // the running downloads are competing for).
  static void widen_range(uint64_t* __restrict__ dst, const uint32_t* __restrict__ src, size_t lo, size_t hi) {
#if defined(__SSE2__) && defined(__x86_64__)
    static const bool streaming = [] {
      const char* e = std::getenv("FQ_HOST_WIDEN_NT");
      return !(e && *e == '0');
    }();
    if (streaming) {
      for (size_t i = lo; i < hi; ++i) _mm_stream_si64(reinterpret_cast<long long*>(dst + i), static_cast<long long>(src[i]));
      _mm_sfence();
      return;
    }
#endif
    for (size_t i = lo; i < hi; ++i) dst[i] = src[i];
  }
  void drive() {
    for (;;) {
      Job job;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_job_.wait(g, [this] { return stop_ || !jobs_.empty(); });
        if (jobs_.empty()) return;
        job = jobs_.front();
        jobs_.pop_front();
      }
      size_t lo = 0;
      const size_t n = job.n;
      while (n - lo > 65536) {
        const size_t hi = (n + lo) / 2;
        run_wave(job.buf, n, lo, hi);
        lo = hi;
      }
      widen_serial(job.buf, n, lo);
      {
        std::lock_guard<std::mutex> g(m_);
        --pending_;
      }
      cv_idle_.notify_all();
    }
  }
  void run_wave(uint64_t* buf, size_t n, size_t lo, size_t hi) {
    {
      std::lock_guard<std::mutex> g(m_);
      wave_buf_ = buf, wave_n_ = n, wave_lo_ = lo, wave_hi_ = hi;
      wave_left_ = nthreads_;
      ++wave_id_;
    }
    cv_wave_.notify_all();
    std::unique_lock<std::mutex> g(m_);
    cv_done_.wait(g, [this] { return wave_left_ == 0; });
  }
  void worker(int t) {
    uint64_t seen = 0;
    for (;;) {
      uint64_t* buf;
      size_t n, lo, hi;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_wave_.wait(g, [&] { return stop_ || wave_id_ != seen; });
        if (stop_) return;
        seen = wave_id_;
        buf = wave_buf_, n = wave_n_, lo = wave_lo_, hi = wave_hi_;
      }
      const size_t len = hi - lo, per = (len + size_t(nthreads_) - 1) / size_t(nthreads_);
      const size_t a = lo + per * size_t(t), b = a + per < hi ? a + per : hi;
      if (a < b) widen_range(buf, reinterpret_cast<const uint32_t*>(buf) + n, a, b);
      {
        std::lock_guard<std::mutex> g(m_);
        --wave_left_;
      }
      cv_done_.notify_all();
    }
  }

  int nthreads_;
  std::vector<std::thread> workers_;
  std::thread driver_;
  std::mutex m_;
  std::condition_variable cv_job_, cv_wave_, cv_done_, cv_idle_;
  std::deque<Job> jobs_;
  size_t pending_ = 0;
  bool stop_ = false;
  uint64_t* wave_buf_ = nullptr;
  size_t wave_n_ = 0, wave_lo_ = 0, wave_hi_ = 0;
  int wave_left_ = 0;
  uint64_t wave_id_ = 0;
};

// threads of the pool: FQ_HOST_WIDEN_THREADS, else hardware concurrency - 1 capped at 16; 0 disables host widening
inline int host_widen_threads() {
  if (const char* e = std::getenv("FQ_HOST_WIDEN_THREADS")) return std::atoi(e);
  const unsigned hw = std::thread::hardware_concurrency();
  if (hw < 8) return 0;
  return int(hw - 1 < 16 ? hw - 1 : 16);
}

}  // namespace fq
