// Host side of the result transport (no kernels in this file; compiled by g++).
//
// The reference API returns dense float64 rows on the host (shot_parallelization.py:183): 288 MB for the 102k SHOT
// rows of the benchmark, 5 ms of PCIe — six times the compute. A SHOT row is ~86 % zeros (each neighbour touches at
// most five of the 352 bins), so the rows cross PCIe compacted (csrc/transport.cu: per-row offsets, uint16 columns,
// float32 values, ~30 MB) and the dense float64 array is rebuilt here by a small pool of host threads: the
// zero-fill starts when the call starts and runs under the upload and the kernels, the scatter of the non-zeros
// follows the copy. float64(float32 x) is exact, so the array is the one a dense float64 copy would deliver.
#include <immintrin.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#if defined(__linux__)
#include <sys/mman.h>
#endif

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/shotfpfh_b200.h"

namespace sf {
void set_error(const char* fmt, ...);  // grid.cu
}

namespace {

// Persistent workers (creating 16 threads per call would cost more than the work). A job is a function of the part
// index, part p runs on worker p. One job at a time; `start` returns at once, `wait` blocks until the job is done.
class Pool {
 public:
  explicit Pool(int workers) {
    for (int w = 0; w < workers; ++w) threads_.emplace_back([this, w] { loop(w); });
  }
  ~Pool() {
    wait();
    {
      std::lock_guard<std::mutex> lock(m_);
      stop_ = true;
      ++generation_;
    }
    wake_.notify_all();
    for (auto& t : threads_) t.join();
  }
  int workers() const { return int(threads_.size()); }

  void start(int parts, std::function<void(int)> job) {
    wait();
    parts = std::max(1, std::min(parts, workers()));
    {
      std::lock_guard<std::mutex> lock(m_);
      job_ = std::move(job);
      parts_ = parts;
      pending_ = parts;
      ++generation_;
    }
    wake_.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lock(m_);
    done_.wait(lock, [this] { return pending_ == 0; });
  }

 private:
  void loop(int p) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lock(m_);
        wake_.wait(lock, [&] { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
        if (p >= parts_) continue;
      }
      job_(p);  // job_ is not replaced before pending_ drops to 0
      {
        std::lock_guard<std::mutex> lock(m_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }

  std::vector<std::thread> threads_;
  std::mutex m_;
  std::condition_variable wake_, done_;
  std::function<void(int)> job_;
  int parts_ = 1, pending_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false;
};

std::mutex g_pool_mutex;
Pool* g_pool = nullptr;

Pool& pool_for(int threads) {  // g_pool_mutex held
  if (g_pool == nullptr || g_pool->workers() < threads) {
    delete g_pool;  // waits for a job in flight
    g_pool = new Pool(threads);
  }
  return *g_pool;
}

int clamp_threads(int32_t threads) { return std::max(1, std::min<int32_t>(threads, 64)); }

// One dense row: zeros and the row's non-zeros assembled in a cache-resident buffer, then streamed out once with
// non-temporal stores. (Zero-filling the array first and scattering into it afterwards touches every line twice
// and pays a read-for-ownership on the second visit: 3x the memory traffic.)
constexpr int kMaxWidth = 4096;

__attribute__((target("avx2"))) void stream_row_avx2(const double* buf, double* dst, int width) {
  int i = 0;
  while (i < width && (reinterpret_cast<uintptr_t>(dst + i) & 31u)) { dst[i] = buf[i]; ++i; }
  for (; i + 4 <= width; i += 4) _mm256_stream_pd(dst + i, _mm256_loadu_pd(buf + i));
  for (; i < width; ++i) dst[i] = buf[i];
}

// rows [lo, hi); returns false on a malformed row. A ring of row buffers: row r is streamed out two rows after it
// was assembled, when its scalar stores have left the store buffer (a 32-byte load over fresh 8-byte stores cannot
// be forwarded and stalls for each of them: measured 2x on the whole pass).
bool expand_rows(const int64_t* offsets, const uint16_t* cols, const float* vals, int width, double* dst, int64_t lo,
                 int64_t hi, bool avx2) {
  constexpr int kRing = 4, kLag = 2;
  static thread_local double* ring = nullptr;
  if (ring == nullptr) ring = static_cast<double*>(aligned_alloc(64, sizeof(double) * kRing * kMaxWidth));
  for (int k = 0; k < kRing; ++k) memset(ring + k * kMaxWidth, 0, sizeof(double) * width);
  auto flush = [&](int64_t r) {  // stream row r out and clear what it had set
    double* buf = ring + (r % kRing) * kMaxWidth;
    if (avx2) stream_row_avx2(buf, dst + r * width, width);
    else memcpy(dst + r * width, buf, sizeof(double) * width);
    for (int64_t i = offsets[r]; i < offsets[r + 1]; ++i) buf[cols[i]] = 0.0;
  };
  for (int64_t r = lo; r < hi; ++r) {
    const int64_t b = offsets[r], e = offsets[r + 1];
    if (b > e || e - b > width) return false;
    double* buf = ring + (r % kRing) * kMaxWidth;
    for (int64_t i = b; i < e; ++i) {
      if (cols[i] >= width) return false;
      buf[cols[i]] = double(vals[i]);
    }
    if (r - kLag >= lo) flush(r - kLag);
  }
  for (int64_t r = std::max(lo, hi - kLag); r < hi; ++r) flush(r);
  if (avx2) _mm_sfence();
  return true;
}

// dst[i] = double(src[i]) (exact), streamed out with non-temporal stores: the float64 array is written once and not
// read back by these threads.
__attribute__((target("avx2"))) void widen_avx2(const float* src, double* dst, int64_t n) {
  int64_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31u)) { dst[i] = double(src[i]); ++i; }
  for (; i + 8 <= n; i += 8) {
    const __m256 v = _mm256_loadu_ps(src + i);
    _mm256_stream_pd(dst + i, _mm256_cvtps_pd(_mm256_castps256_ps128(v)));
    _mm256_stream_pd(dst + i + 4, _mm256_cvtps_pd(_mm256_extractf128_ps(v, 1)));
  }
  for (; i < n; ++i) dst[i] = double(src[i]);
  _mm_sfence();
}

// part p of `parts` of [0, n), cut on multiples of `quantum`
void share(int64_t n, int p, int parts, int64_t quantum, int64_t& lo, int64_t& hi) {
  const int64_t per = ((n + parts - 1) / parts + quantum - 1) / quantum * quantum;
  lo = std::min<int64_t>(n, per * p);
  hi = std::min<int64_t>(n, lo + per);
}

}  // namespace

std::atomic<int> g_job_failed{0};

extern "C" int sf_host_expand_rows_begin(const int64_t* offsets, const uint16_t* cols, const float* vals, int64_t n_rows,
                                         int32_t width, double* dst, int32_t threads) {
  if (n_rows < 0 || width <= 0 || width > kMaxWidth ||
      (n_rows > 0 && (offsets == nullptr || dst == nullptr || cols == nullptr || vals == nullptr))) {
    sf::set_error("sf_host_expand_rows_begin: bad arguments (width <= %d)", kMaxWidth);
    return SF_ERR_ARG;
  }
  if (n_rows == 0) return SF_OK;
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  Pool& pool = pool_for(clamp_threads(threads));
  const int parts =
      int(std::max<int64_t>(1, std::min<int64_t>(std::min(clamp_threads(threads), pool.workers()), n_rows / 256)));
  static const bool avx2 = __builtin_cpu_supports("avx2");
  pool.start(parts, [=](int p) {  // waits for the previous job first: successive blocks are expanded in order
    int64_t lo, hi;
    share(n_rows, p, parts, 1, lo, hi);
    if (hi > lo && !expand_rows(offsets, cols, vals, width, dst, lo, hi, avx2)) g_job_failed.store(1);
  });
  return SF_OK;
}

extern "C" int sf_host_widen_begin(const float* src, int64_t n, double* dst, int32_t threads) {
  if (n < 0 || (n > 0 && (src == nullptr || dst == nullptr))) {
    sf::set_error("sf_host_widen_begin: bad arguments");
    return SF_ERR_ARG;
  }
  if (n == 0) return SF_OK;
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  Pool& pool = pool_for(clamp_threads(threads));
  const int parts = int(std::max<int64_t>(1, std::min<int64_t>(std::min(clamp_threads(threads), pool.workers()), n / 4096)));
  static const bool avx2 = __builtin_cpu_supports("avx2");
  pool.start(parts, [=](int p) {  // waits for the previous job first
    int64_t lo, hi;
    share(n, p, parts, 16, lo, hi);
    if (hi <= lo) return;
    if (avx2) widen_avx2(src + lo, dst + lo, hi - lo);
    else for (int64_t i = lo; i < hi; ++i) dst[i] = double(src[i]);
  });
  return SF_OK;
}

// Parallel copy of `bytes` from pageable caller memory into a page-locked staging buffer (the DMA engine then takes it
// from there): a single memcpy runs at ~10 GB/s, the pool's threads together at the host's memory bandwidth.
extern "C" int sf_host_copy_begin(const void* src, void* dst, int64_t bytes, int32_t threads) {
  if (bytes < 0 || (bytes > 0 && (src == nullptr || dst == nullptr))) {
    sf::set_error("sf_host_copy_begin: bad arguments");
    return SF_ERR_ARG;
  }
  if (bytes == 0) return SF_OK;
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  Pool& pool = pool_for(clamp_threads(threads));
  const int parts = int(std::max<int64_t>(1, std::min<int64_t>(std::min(clamp_threads(threads), pool.workers()), bytes / 65536)));
  pool.start(parts, [=](int p) {  // waits for the previous job first
    int64_t lo, hi;
    share(bytes, p, parts, 4096, lo, hi);
    if (hi > lo) memcpy(static_cast<char*>(dst) + lo, static_cast<const char*>(src) + lo, size_t(hi - lo));
  });
  return SF_OK;
}

// madvise(MADV_HUGEPAGE) on a fresh result buffer: its first touch then costs one fault per 2 MB instead of per 4 KB
// (288 MB of SHOT rows: 70 000 faults, ~20 ms, spread over the writers). A hint: ignored where THP is off.
extern "C" int sf_host_advise_huge(void* ptr, int64_t bytes) {
#if defined(__linux__)
  if (ptr != nullptr && bytes > 0) {
    const uintptr_t page = 2u << 20, lo = (reinterpret_cast<uintptr_t>(ptr) + page - 1) & ~(page - 1);
    const uintptr_t hi = (reinterpret_cast<uintptr_t>(ptr) + uintptr_t(bytes)) & ~(page - 1);
    if (hi > lo) madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
  }
#endif
  return SF_OK;
}

extern "C" int sf_host_wait(void) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  if (g_pool != nullptr) g_pool->wait();
  if (g_job_failed.exchange(0)) {
    sf::set_error("sf_host_wait: malformed compact rows");
    return SF_ERR_ARG;
  }
  return SF_OK;
}
